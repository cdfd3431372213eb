set -x
T=${TAG:-r1pdl}
timeout 1200 python -m pytest tests/test_decoder_gpu.py tests/test_encoder_gpu.py tests/test_hybrid_gpu.py -m gpu -q --timeout 600 -x 2>&1 | tail -30 > gpurun_out/${T}_pytest.log
cat gpurun_out/${T}_pytest.log
grep -q " failed\|rror" gpurun_out/${T}_pytest.log && exit 1
for pd in 1 0; do
LXG_PDL=$pd timeout 900 python - <<'PY'
import json, torch, bench
print(json.dumps(bench.decoder_numbers(torch.device("cuda", 0), cpu=False)))
PY
done
