set -x
T=${TAG:-r1v}
timeout 900 python -m pytest tests/test_search_gpu.py -m gpu -q --timeout 300 -x 2>&1 | tail -25 > gpurun_out/${T}_pytest_gpu.log
tail -6 gpurun_out/${T}_pytest_gpu.log
for wl in cfg1 cfg4; do
    timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --cpu-budget 5 2>gpurun_out/${T}_bench_$wl.err > gpurun_out/${T}_bench_$wl.json
    python -c "
import sys,json
j=json.load(open('gpurun_out/${T}_bench_$wl.json')); r=j['roofline']; print('$wl', 'qps', j['value'], 'e2e', j['e2e']['value'], 'scan_ms', r['ms_per_launch'], r['bound'], 'frac', r['frac'], 'merge', r['merge_ms_per_launch'], 'cpu', j.get('cpu_baseline'))"
    tail -2 gpurun_out/${T}_bench_$wl.err
done
