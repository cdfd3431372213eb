set -x
T=${TAG:-r1ship}
timeout 1200 python -m pytest tests/test_search_gpu.py -m gpu -q --timeout 600 -x 2>&1 | tail -3
timeout 900 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/${T}_bench_cfg2.json 2> gpurun_out/${T}_bench.err
tail -3 gpurun_out/${T}_bench.err
python - <<PY
import json
j=json.load(open('gpurun_out/${T}_bench_cfg2.json')); e=j.pop('extra', {})
print(j['value'], j['e2e']['value'], j['roofline']['ms_per_launch']); print(json.dumps(e.get('shipped index shape: 400k x 1024 fp32')))
PY
