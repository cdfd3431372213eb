set -x
T=${TAG:-r1t}
timeout 900 python -m pytest tests/test_search_gpu.py -m gpu -q --timeout 300 -x 2>&1 | tail -25 > gpurun_out/${T}_pytest_gpu.log
tail -4 gpurun_out/${T}_pytest_gpu.log
grep -q " passed" gpurun_out/${T}_pytest_gpu.log || exit 1
grep -q " failed" gpurun_out/${T}_pytest_gpu.log && exit 1
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/${T}_bench_cfg2.json 2> gpurun_out/${T}_bench_cfg2.err
python - <<PY
import json
j=json.load(open('gpurun_out/${T}_bench_cfg2.json')); e=j['extra']; e.pop('encoder',None)
for k,v in e.items(): print(k, v)
print(j['value'], j['e2e'], j['roofline'])
PY
