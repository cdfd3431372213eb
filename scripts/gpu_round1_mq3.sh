set -x
timeout 900 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r1mq3_bench_cfg2.json 2> gpurun_out/r1mq3_bench.err
python - <<PY
import json
j=json.load(open('gpurun_out/r1mq3_bench_cfg2.json')); e=j.pop('extra', {})
print(j['value'], j['e2e']['value'], j['roofline']['ms_per_launch'], j['roofline']['merge_ms_per_launch'])
s=e.get('shipped index shape: 400k x 1024 fp32'); print({k:(v['ms_per_step'], v['merge_ms']) for k,v in s.items()}); print({k:(v['ms_per_step'], v['merge_ms']) for k,v in e.items() if k.startswith('Q=')})
PY
