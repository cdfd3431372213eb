#!/bin/bash
# compute-sanitizer passes over the hand-rolled mbarrier / cluster / TMEM / grid-barrier protocols (run on the GPU box):
#   gpurun --timeout 3300 -- 'bash scripts/gpu_sanitizers.sh TAG [new]'
# "new" restricts the run to the round-2 kernels (single-kernel encoder, three-kernel merge);
# "r3" = the scan with level warps + the tcgen05 decoder attention (second session of round 2).
TAG=${1:-san}
export SAN_TIMEOUT=${SAN_TIMEOUT:-700}
S="tests/test_search_gpu.py -m gpu -k 'ragged or adversarial or kat or duplicate or k_larger or zero_norm or merge_topk'"
E="tests/test_encoder_gpu.py tests/test_decoder_gpu.py -m gpu -k golden"
N="tests/test_search_gpu.py -m gpu -k 'three_stage'"
F="tests/test_encoder_gpu.py -m gpu -k 'query_path and (tiny or (minilm and 1-16) or (bge and 1-9))'"
A="tests/test_decoder_gpu.py -m gpu -k 'embedding_matches and tiny and (70 or 300 or 520 or 260)'"
if [ "${2:-}" = "r3" ]; then
  bash scripts/gpu_run.sh $TAG "san:memcheck:$S" "san:memcheck:$A" "san:racecheck:$S" "san:racecheck:$A" "san:synccheck:$S" "san:synccheck:$A"
elif [ "${2:-}" = "new" ]; then
  bash scripts/gpu_run.sh $TAG "san:memcheck:$N" "san:memcheck:$F" "san:racecheck:$N" "san:racecheck:$F" "san:synccheck:$N" "san:synccheck:$F"
else
  bash scripts/gpu_run.sh $TAG "san:memcheck:$S" "san:memcheck:$E" "san:racecheck:$S" "san:synccheck:$S" "san:synccheck:$E" "san:racecheck:$E"
fi
