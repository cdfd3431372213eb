set -x
timeout 900 python -m pytest tests/test_encoder_gpu.py -m gpu -q --timeout 600 2>&1 | tail -3
for pd in 1 0; do
LXG_PDL=$pd timeout 600 python - <<'PY'
import json, torch, bench
r = bench.encoder_numbers(torch.device("cuda", 0), cpu=False)
print({k: {kk: (vv.get("ms_per_call"), vv.get("e2e_ms_per_call")) for kk, vv in v.items()} for k, v in r.items()})
PY
done
