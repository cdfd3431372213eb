set -x
T=${TAG:-r1Z2}
timeout 1500 python -m pytest tests/test_search_gpu.py -m gpu -q --timeout 900 2>&1 | tail -3
timeout 400 python bench.py --workload cfg3 --steps 30 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/${T}_bench_cfg3.json 2> gpurun_out/${T}_bench.err
cat gpurun_out/${T}_bench_cfg3.json | cut -c1-1500
timeout 400 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-extra > gpurun_out/${T}_bench_cfg2.json 2>> gpurun_out/${T}_bench.err
cat gpurun_out/${T}_bench_cfg2.json | cut -c1-1500
timeout 400 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 3 -c 1 -o gpurun_out/${T}_prof_scan_cfg3 python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${T}_ncu_scan_cfg3.log 2>&1
python scripts/ncu_summary.py gpurun_out/${T}_prof_scan_cfg3.ncu-rep > gpurun_out/${T}_ncu_scan_cfg3.txt 2>&1; grep -E "dram__bytes_read.sum \[|dram__bytes_write.sum \[|gpu__time_duration|tensor_cycles_active.avg.pct_of_peak_sustained_active|cycles_elapsed.avg.per_second" gpurun_out/${T}_ncu_scan_cfg3.txt
