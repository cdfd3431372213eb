#!/bin/bash
# compute-sanitizer passes over the hand-rolled mbarrier / cluster / TMEM protocols (run on the GPU box):
#   gpurun --timeout 3300 -- 'bash scripts/gpu_sanitizers.sh TAG'
TAG=${1:-san}
export SAN_TIMEOUT=${SAN_TIMEOUT:-700}
S="tests/test_search_gpu.py -m gpu -k 'ragged or adversarial or kat or duplicate or k_larger or zero_norm or merge_topk'"
E="tests/test_encoder_gpu.py tests/test_decoder_gpu.py -m gpu -k golden"
bash scripts/gpu_run.sh $TAG "san:memcheck:$S" "san:memcheck:$E" "san:racecheck:$S" "san:synccheck:$S" "san:synccheck:$E" "san:racecheck:$E"
