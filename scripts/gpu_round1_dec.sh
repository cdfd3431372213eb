set -x
T=${TAG:-r1dec}
timeout 900 python -m pytest tests/test_decoder_gpu.py -m gpu -q --timeout 600 -x 2>&1 | tail -40 > gpurun_out/${T}_pytest_decoder.log
cat gpurun_out/${T}_pytest_decoder.log
