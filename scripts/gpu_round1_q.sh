set -x
T=${TAG:-r1q}
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -25 > gpurun_out/${T}_pytest_gpu.log
tail -8 gpurun_out/${T}_pytest_gpu.log
grep -q " passed" gpurun_out/${T}_pytest_gpu.log || exit 1
grep -q " failed" gpurun_out/${T}_pytest_gpu.log && exit 1
for wl in cfg2 cfg3; do
    timeout 300 python bench.py --workload $wl --steps 60 --warmup 5 --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.read()); r=j['roofline']; print('$wl', 'qps', j['value'], 'e2e', j['e2e']['value'], 'scan_ms', r['ms_per_launch'], 'frac', r['frac'], 'prep', r['prep_ms_per_launch'], 'merge', r['merge_ms_per_launch'], 'exact', r['exact_ms_per_launch'], 'clk', j['clocks']['sm_mhz'], j['clocks']['reasons'])"
done
python -c "import __graft_entry__ as g; g.smoke()"
