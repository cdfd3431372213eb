set -x
timeout 1500 python -m pytest tests/test_search_gpu.py tests/test_hybrid_gpu.py -m gpu -q --timeout 900 2>&1 | tail -3
for zc in 1 0 1 0; do
LXG_ZERO_COPY=$zc timeout 400 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-extra > gpurun_out/r1zc_bench_cfg2_zc$zc.json 2> gpurun_out/r1zc_bench.err
python -c "
import json; j=json.load(open('gpurun_out/r1zc_bench_cfg2_zc$zc.json')); print('zc$zc', j['value'], j['e2e']['value'], j['ms_per_step'])"
done
LXG_ZERO_COPY=1 timeout 400 python bench.py --workload cfg3 --steps 30 --warmup 5 --no-cpu-baseline --no-extra | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('cfg3 zc1', j['value'], j['e2e']['value'])"
