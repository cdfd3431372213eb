set -x
T=${TAG:-r1Z}
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -30 > gpurun_out/${T}_pytest_gpu.log
tail -5 gpurun_out/${T}_pytest_gpu.log
timeout 900 python bench.py --steps 200 --warmup 10 > gpurun_out/${T}_bench_cfg2.json 2> gpurun_out/${T}_bench_cfg2.err
tail -c 1500 gpurun_out/${T}_bench_cfg2.err
python - <<PY
import json
j=json.load(open('gpurun_out/${T}_bench_cfg2.json')); e=j.pop('extra', {})
print(json.dumps(j)); print(json.dumps(e.get('qwen3'))); print({k: v for k, v in e.items() if k.startswith('Q=')})
PY
timeout 400 python bench.py --workload cfg3 --steps 30 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/${T}_bench_cfg3.json 2> gpurun_out/${T}_bench_cfg3.err
cat gpurun_out/${T}_bench_cfg3.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_cfg2.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${T}_launches_cfg2.log 2>&1
python scripts/ncu_launches.py gpurun_out/${T}_launches_cfg2.csv > gpurun_out/${T}_launches_cfg2.txt; cat gpurun_out/${T}_launches_cfg2.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 3 -c 1 -o gpurun_out/${T}_prof_scan_cfg3 python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${T}_ncu_scan_cfg3.log 2>&1
python scripts/ncu_summary.py gpurun_out/${T}_prof_scan_cfg3.ncu-rep > gpurun_out/${T}_ncu_scan_cfg3.txt 2>&1; grep -E "dram__bytes_read.sum \[|dram__bytes_write.sum \[|gpu__time_duration|tensor_cycles_active.avg.pct_of_peak_sustained_active|cycles_elapsed.avg.per_second" gpurun_out/${T}_ncu_scan_cfg3.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 3 -c 1 -o gpurun_out/${T}_prof_scan_cfg2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${T}_ncu_scan_cfg2.log 2>&1
python scripts/ncu_summary.py gpurun_out/${T}_prof_scan_cfg2.ncu-rep > gpurun_out/${T}_ncu_scan_cfg2.txt 2>&1; grep -E "dram__bytes_read.sum \[|dram__bytes_write.sum \[|gpu__time_duration|tensor_cycles_active.avg.pct_of_peak_sustained_active|cycles_elapsed.avg.per_second" gpurun_out/${T}_ncu_scan_cfg2.txt
python -c "import __graft_entry__ as g; g.smoke()"
