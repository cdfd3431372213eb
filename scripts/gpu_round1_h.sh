set -x
T=${TAG:-r1h}
timeout 600 python -m pytest tests/test_search_gpu.py -m gpu -q --timeout 120 -x 2>&1 | tail -25 > gpurun_out/${T}_pytest_search.log
tail -5 gpurun_out/${T}_pytest_search.log
grep -q " passed" gpurun_out/${T}_pytest_search.log || exit 1
grep -q " failed" gpurun_out/${T}_pytest_search.log && exit 1
bash scripts/gpu_perf_modes.sh 2>&1 | tee gpurun_out/${T}_perf_modes.log
