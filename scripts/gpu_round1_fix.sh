set -x
T=${TAG:-r1G}
timeout 900 python -m pytest tests/test_search_gpu.py tests/test_encoder_gpu.py -m gpu -q --timeout 600 -x 2>&1 | tail -5
timeout 400 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-extra > gpurun_out/${T}_bench_cfg2.json 2> gpurun_out/${T}_bench_cfg2.err
python -c "
import json; j=json.load(open('gpurun_out/${T}_bench_cfg2.json')); print(j['value'], j['e2e']['value'], j['roofline'])"
for am in 1 0; do
LXG_SCAN_ASMEM=$am timeout 400 python bench.py --workload cfg3 --steps 30 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/${T}_bench_cfg3_asmem$am.json 2> gpurun_out/${T}_bench_cfg3.err
python -c "
import json; j=json.load(open('gpurun_out/${T}_bench_cfg3_asmem$am.json')); print('asmem$am', j['value'], j['e2e']['value'], j['roofline']['ms_per_launch'], j['roofline']['frac'])"
done
