set -x
T=${TAG:-r1sync2}
for mb in 56 28 14; do
LXG_SCAN_SYNC_MB=$mb timeout 400 python bench.py --workload cfg3 --steps 30 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/${T}_bench_cfg3_mb$mb.json 2> gpurun_out/${T}_bench.err
python - <<PY
import json
j = json.load(open("gpurun_out/${T}_bench_cfg3_mb$mb.json")); print("mb$mb", j["value"], j["e2e"]["value"], j["roofline"]["ms_per_launch"], j["roofline"]["frac"], j["clocks"]["sm_mhz"])
PY
LXG_SCAN_SYNC_MB=$mb timeout 400 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum --clock-control none -k regex:scan_topk -s 3 -c 1 python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline --no-extra 2>&1 | grep -E "dram__bytes_read|gpu__time_duration"
done
