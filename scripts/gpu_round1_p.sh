set -x
T=${TAG:-r1p}
timeout 900 python -m pytest tests/test_search_gpu.py -m gpu -q --timeout 180 -x 2>&1 | tail -25 > gpurun_out/${T}_pytest_search.log
tail -8 gpurun_out/${T}_pytest_search.log
grep -q " passed" gpurun_out/${T}_pytest_search.log || exit 1
grep -q " failed" gpurun_out/${T}_pytest_search.log && exit 1
for rep in 1 2; do
for lib in liblxg.so liblxg_g2.so; do
  for wl in cfg2 cfg3; do
    LXG_LIB_PATH=$PWD/lean_explore_b200/$lib timeout 300 python bench.py --workload $wl --steps 60 --warmup 5 --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.read()); r=j['roofline']; print('$lib $wl', 'qps', j['value'], 'scan_ms', r['ms_per_launch'], 'frac', r['frac'], 'merge', r['merge_ms_per_launch'], 'clk', j['clocks']['sm_mhz'], j['clocks']['reasons'])"
  done
done
done
