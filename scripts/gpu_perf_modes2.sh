for rep in 1 2; do
for wl in cfg2; do
  for pm in 1 3 2 0; do
    LXG_SCAN_PERF_MODE=$pm timeout 300 python bench.py --workload $wl --steps 60 --warmup 5 --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print('$wl perf_mode=$pm', 'scan_ms', j['roofline']['ms_per_launch'], 'frac', j['roofline']['frac'], 'clk', j['clocks']['sm_mhz'], j['clocks']['reasons'])"
  done
done
done
