set -x
T=${TAG:-r1x}
timeout 900 python -m pytest tests/test_encoder_gpu.py -m gpu -q --timeout 300 2>&1 | tail -30 > gpurun_out/${T}_pytest_encoder.log
tail -12 gpurun_out/${T}_pytest_encoder.log
TAG=$T bash scripts/gpu_encoder_prof.sh
