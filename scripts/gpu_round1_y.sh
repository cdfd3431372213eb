set -x
T=${TAG:-r1y}
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -30 > gpurun_out/${T}_pytest_gpu.log
tail -5 gpurun_out/${T}_pytest_gpu.log
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/${T}_bench_cfg2.json 2> gpurun_out/${T}_bench_cfg2.err
python - <<PY
import json
j=json.load(open('gpurun_out/${T}_bench_cfg2.json')); e=j['extra']
print(json.dumps(e.pop('encoder',None), indent=1))
for k,v in e.items(): print(k, v)
print(j['value'], j['e2e'], j['roofline'], j['cpu_baseline'])
PY
python -c "import __graft_entry__ as g; g.smoke()"
