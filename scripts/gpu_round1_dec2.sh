set -x
T=${TAG:-r1dec2}
timeout 900 python - <<'PY' > gpurun_out/${T}_decoder_numbers.json 2> gpurun_out/${T}_decoder_numbers.err
import json, torch, bench
print(json.dumps(bench.decoder_numbers(torch.device("cuda", 0), cpu=True), indent=1))
PY
cat gpurun_out/${T}_decoder_numbers.json; tail -5 gpurun_out/${T}_decoder_numbers.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_rerank.csv python - <<'PY' > /dev/null 2>&1
import numpy as np, torch
from oracle import qwen3_decoder as qd
from lean_explore_b200.decoder import Qwen3Decoder
model, cfg = qd.make_model("qwen3-0.6b", seed=0)
dec = Qwen3Decoder(model.state_dict(), hidden=cfg.hidden_size, layers=cfg.num_hidden_layers, heads=cfg.num_attention_heads, kv_heads=cfg.num_key_value_heads, ffn=cfg.intermediate_size, head_dim=cfg.head_dim, rms_eps=cfg.rms_norm_eps, rope_theta=1e6)
ids, mask = qd.make_inputs(16, 256, seed=5)
dec.rerank_ids(ids, mask, 1837, 3082)
PY
python scripts/ncu_launches.py gpurun_out/${T}_launches_rerank.csv > gpurun_out/${T}_launches_rerank.txt 2>&1; tail -30 gpurun_out/${T}_launches_rerank.txt
