set -x
T=${TAG:-r1e2e}
timeout 900 python -m pytest tests/test_hybrid_gpu.py -m gpu -q --timeout 600 -x 2>&1 | tail -40 > gpurun_out/${T}_pytest.log
cat gpurun_out/${T}_pytest.log
