set -x
T=${TAG:-r1o}
timeout 600 python -m pytest tests/test_search_gpu.py -m gpu -q --timeout 120 -x 2>&1 | tail -25 > gpurun_out/${T}_pytest_search.log
tail -5 gpurun_out/${T}_pytest_search.log
grep -q " passed" gpurun_out/${T}_pytest_search.log || exit 1
grep -q " failed" gpurun_out/${T}_pytest_search.log && exit 1
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/${T}_bench_cfg2.json 2> gpurun_out/${T}_bench_cfg2.err
python -c "
import json; j=json.load(open('gpurun_out/${T}_bench_cfg2.json')); e=j['extra']; e.pop('encoder',None); print(json.dumps(e)); print(j['value'], j['e2e'], j['roofline'])"
timeout 400 python bench.py --workload cfg3 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_cfg3.json 2> gpurun_out/${T}_bench_cfg3.err
python -c "
import json; j=json.load(open('gpurun_out/${T}_bench_cfg3.json')); print(j['value'], j['e2e'], j['roofline'])"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 3 -c 1 -o gpurun_out/${T}_prof_scan_cfg2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${T}_ncu_scan_cfg2.log 2>&1
