set -x
T=${TAG:-r1sync}
timeout 1500 python -m pytest tests/test_search_gpu.py -m gpu -q --timeout 900 -x 2>&1 | tail -3
for sy in 1 0; do
LXG_SCAN_SYNC=$sy timeout 400 python bench.py --workload cfg3 --steps 30 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/${T}_bench_cfg3_sync$sy.json 2> gpurun_out/${T}_bench.err
LXG_SCAN_SYNC=$sy timeout 400 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-extra > gpurun_out/${T}_bench_cfg2_sync$sy.json 2>> gpurun_out/${T}_bench.err
python - <<PY
import json
for wl in ("cfg3", "cfg2"):
    j = json.load(open("gpurun_out/${T}_bench_%s_sync$sy.json" % wl)); print("sync$sy", wl, j["value"], j["e2e"]["value"], j["roofline"]["ms_per_launch"], j["roofline"]["frac"], j["clocks"]["sm_mhz"])
PY
LXG_SCAN_SYNC=$sy timeout 400 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum --clock-control none -k regex:scan_topk -s 3 -c 2 python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline --no-extra 2>&1 | grep -E "dram__bytes_read|gpu__time_duration"
done
