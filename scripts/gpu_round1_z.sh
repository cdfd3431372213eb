set -x
T=${TAG:-r1z}
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -30 > gpurun_out/${T}_pytest_gpu.log
tail -5 gpurun_out/${T}_pytest_gpu.log
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/${T}_bench_cfg2.json 2> gpurun_out/${T}_bench_cfg2.err
tail -c 3000 gpurun_out/${T}_bench_cfg2.json
timeout 400 python bench.py --workload cfg3 --steps 30 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/${T}_bench_cfg3.json 2> gpurun_out/${T}_bench_cfg3.err
cat gpurun_out/${T}_bench_cfg3.json
python -c "import __graft_entry__ as g; g.smoke()"
