T=${TAG:-pm}
for wl in cfg2 cfg3; do
  for pm in 0 1 2; do
    LXG_SCAN_PERF_MODE=$pm timeout 300 python bench.py --workload $wl --steps 40 --warmup 5 --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print('$wl perf_mode=$pm', 'scan_ms', j['roofline']['ms_per_launch'], 'frac', j['roofline']['frac'], 'clk', j['clocks']['sm_mhz'], j['clocks']['reasons'])"
  done
  LXG_SCAN_SINGLE=1 LXG_SCAN_PERF_MODE=1 timeout 300 python bench.py --workload $wl --steps 40 --warmup 5 --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print('$wl single perf_mode=1', 'scan_ms', j['roofline']['ms_per_launch'], 'frac', j['roofline']['frac'], 'clk', j['clocks']['sm_mhz'], j['clocks']['reasons'])"
done
