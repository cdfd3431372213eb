set -x
timeout 1500 python -m pytest tests/test_search_gpu.py -m gpu -q --timeout 900 -x 2>&1 | tail -3
timeout 900 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r1mq2_bench_cfg2.json 2> gpurun_out/r1mq2_bench.err
python - <<PY
import json
j=json.load(open('gpurun_out/r1mq2_bench_cfg2.json')); e=j.pop('extra', {})
print(j['value'], j['e2e']['value'], j['roofline']['ms_per_launch']); print(json.dumps(e.get('shipped index shape: 400k x 1024 fp32'))); print(json.dumps(e.get('Q=1 k=1000 (engine default faiss_k)')), json.dumps(e.get('Q=1')))
PY
