set -x
T=${TAG:-r1Z3}
time timeout 900 python bench.py > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err
tail -4 gpurun_out/${T}_bench_default.err
python - <<PY
import json
j=json.load(open('gpurun_out/${T}_bench_default.json')); e=j.pop('extra', {})
print(json.dumps(j)[:1600]); print(list(e)); print(json.dumps(e.get('qwen3'))[:900]); print(json.dumps(e.get('shipped index shape: 400k x 1024 fp32'))[:900])
PY
time timeout 900 python bench.py --impl reference > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
cut -c1-1200 gpurun_out/${T}_bench_reference.json
