set -x
T=${TAG:-r1f}
timeout 600 python -m pytest tests/test_search_gpu.py -m gpu -q --timeout 120 -x 2>&1 | tail -25 > gpurun_out/${T}_pytest_search.log
tail -25 gpurun_out/${T}_pytest_search.log
grep -q " passed" gpurun_out/${T}_pytest_search.log || exit 1
grep -q " failed" gpurun_out/${T}_pytest_search.log && exit 1
timeout 400 python bench.py --steps 100 --warmup 10 > gpurun_out/${T}_bench_cfg2.json 2> gpurun_out/${T}_bench_cfg2.err
timeout 400 python bench.py --workload cfg3 --steps 30 --warmup 5 --cpu-budget 6 > gpurun_out/${T}_bench_cfg3.json 2> gpurun_out/${T}_bench_cfg3.err
LXG_SCAN_NOLEVEL=1 timeout 400 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/${T}_bench_cfg2_nolevel.json 2> gpurun_out/${T}_bench_cfg2_nolevel.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 3 -c 1 -o gpurun_out/${T}_prof_scan_cfg2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${T}_ncu_full_cfg2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 3 -c 1 -o gpurun_out/${T}_prof_scan_cfg3 python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${T}_ncu_full_cfg3.log 2>&1
cat gpurun_out/${T}_bench_cfg2.json gpurun_out/${T}_bench_cfg3.json gpurun_out/${T}_bench_cfg2_nolevel.json
tail -5 gpurun_out/${T}_bench_cfg2.err
