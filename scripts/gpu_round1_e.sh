set -x
T=${TAG:-r1e}
timeout 600 python -m pytest tests/test_search_gpu.py -m gpu -q --timeout 120 -x 2>&1 | tail -15 > gpurun_out/${T}_pytest_search.log
tail -3 gpurun_out/${T}_pytest_search.log
timeout 900 python -m pytest tests/test_encoder_gpu.py -m gpu -q --timeout 300 2>&1 | tail -40 > gpurun_out/${T}_pytest_encoder.log
tail -25 gpurun_out/${T}_pytest_encoder.log
grep -q " passed" gpurun_out/${T}_pytest_search.log || exit 1
timeout 400 python bench.py --steps 100 --warmup 10 > gpurun_out/${T}_bench_cfg2.json 2> gpurun_out/${T}_bench_cfg2.err
timeout 400 python bench.py --workload cfg3 --steps 30 --warmup 5 --cpu-budget 6 > gpurun_out/${T}_bench_cfg3.json 2> gpurun_out/${T}_bench_cfg3.err
LXG_SCAN_NOLEVEL=1 timeout 400 python bench.py --workload cfg3 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_cfg3_nolevel.json 2> gpurun_out/${T}_bench_cfg3_nolevel.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 3 -c 1 -o gpurun_out/${T}_prof_scan_cfg3 python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${T}_ncu_full_cfg3.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 3 -c 1 -o gpurun_out/${T}_prof_scan_cfg2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${T}_ncu_full_cfg2.log 2>&1
cat gpurun_out/${T}_bench_cfg2.json gpurun_out/${T}_bench_cfg3.json gpurun_out/${T}_bench_cfg3_nolevel.json
