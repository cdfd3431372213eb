T=${TAG:-pm1}
LXG_SCAN_PERF_MODE=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 3 -c 1 -o gpurun_out/${T}_prof_pm1_cfg2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${T}_ncu_pm1_cfg2.log 2>&1
LXG_SCAN_PERF_MODE=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 3 -c 1 -o gpurun_out/${T}_prof_pm1_cfg3 python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${T}_ncu_pm1_cfg3.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 3 -c 1 -o gpurun_out/${T}_prof_scan_cfg2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${T}_ncu_cfg2.log 2>&1
ls -la gpurun_out
