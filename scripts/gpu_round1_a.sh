set -x
timeout 900 python -m pytest tests -m gpu -q --timeout 180 --durations=8 2>&1 | tail -30 > gpurun_out/r1_pytest.log
timeout 400 python bench.py --steps 100 --warmup 10 > gpurun_out/r1_bench_cfg2.json 2> gpurun_out/r1_bench_cfg2.err
timeout 400 python bench.py --workload cfg3 --steps 30 --warmup 5 --cpu-budget 10 > gpurun_out/r1_bench_cfg3.json 2> gpurun_out/r1_bench_cfg3.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r1_launches_cfg2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r1_ncu_cfg2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 3 -c 1 -o gpurun_out/r1_prof_scan_cfg2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r1_ncu_full_cfg2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 3 -c 1 -o gpurun_out/r1_prof_scan_cfg3 python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r1_ncu_full_cfg3.log 2>&1
tail -5 gpurun_out/r1_pytest.log; cat gpurun_out/r1_bench_cfg2.json gpurun_out/r1_bench_cfg3.json
