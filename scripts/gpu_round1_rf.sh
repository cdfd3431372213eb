set -x
T=${TAG:-r1rf}
timeout 1200 python -m pytest tests/test_search_gpu.py -m gpu -q --timeout 600 -x 2>&1 | tail -3
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/${T}_bench_cfg2.json 2> gpurun_out/${T}_bench.err
python - <<PY
import json
j=json.load(open('gpurun_out/${T}_bench_cfg2.json')); e=j.pop('extra', {})
print(j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['ms_per_launch'], j['roofline']['frac']); print({k: (v['ms_per_step'], v['scan_ms'], v['hbm_frac']) for k, v in e.items() if k.startswith('Q=')})
PY
timeout 400 python bench.py --workload cfg3 --steps 30 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/${T}_bench_cfg3.json 2>> gpurun_out/${T}_bench.err
python -c "
import json; j=json.load(open('gpurun_out/${T}_bench_cfg3.json')); print('cfg3', j['value'], j['e2e']['value'], j['roofline']['ms_per_launch'], j['roofline']['frac'])"
