set -x
T=${TAG:-r1pair}
timeout 1200 python -m pytest tests/test_encoder_gpu.py tests/test_decoder_gpu.py -m gpu -q --timeout 600 -x 2>&1 | tail -30 > gpurun_out/${T}_pytest.log
cat gpurun_out/${T}_pytest.log
grep -q " failed\|rror" gpurun_out/${T}_pytest.log && exit 1
for mode in 0 1; do
LXG_GEMM_SINGLE=$mode timeout 900 python - <<'PY' > gpurun_out/${T}_numbers_single${mode}.json 2> gpurun_out/${T}_numbers.err
import json, torch, bench
out = {"qwen3": bench.decoder_numbers(torch.device("cuda", 0), cpu=False), "encoder": bench.encoder_numbers(torch.device("cuda", 0), cpu=False)}
print(json.dumps(out, indent=1))
PY
cat gpurun_out/${T}_numbers_single${mode}.json; tail -3 gpurun_out/${T}_numbers.err
done
