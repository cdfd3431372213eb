set -x
T=${TAG:-r1s}
timeout 900 python -m pytest tests/test_search_gpu.py -m gpu -q --timeout 300 -x 2>&1 | tail -25 > gpurun_out/${T}_pytest_gpu.log
tail -4 gpurun_out/${T}_pytest_gpu.log
grep -q " passed" gpurun_out/${T}_pytest_gpu.log || exit 1
grep -q " failed" gpurun_out/${T}_pytest_gpu.log && exit 1
python scripts/gpu_latency.py > gpurun_out/${T}_latency.json
python - <<'PY'
import json
j=json.load(open('gpurun_out/r1s_latency.json'))
for k,v in j.items(): print(k, v)
PY
