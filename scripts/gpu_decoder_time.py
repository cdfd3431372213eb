"""Reranker / embedding forwards of the Qwen3-0.6B geometry (random weights): ms per call, host ids in /
host scores out, for A/B runs (LXG_ATTN_TC=0|1, LXG_DECODER_PACK=0|1)."""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bench import qwen3_random_model, ragged_left_padded_ids  # noqa: E402
from lean_explore_b200.decoder import Qwen3Decoder  # noqa: E402

model, cfg = qwen3_random_model()
dec = Qwen3Decoder(model.state_dict(), hidden=cfg.hidden_size, layers=cfg.num_hidden_layers,
                   heads=cfg.num_attention_heads, kv_heads=cfg.num_key_value_heads, ffn=cfg.intermediate_size,
                   head_dim=cfg.head_dim, rms_eps=cfg.rms_norm_eps, rope_theta=1e6)
out = {"LXG_ATTN_TC": os.environ.get("LXG_ATTN_TC"), "LXG_DECODER_PACK": os.environ.get("LXG_DECODER_PACK")}
ref = None
for b, s in ((16, 256), (50, 256), (16, 512), (64, 128)):
    ids, mask = ragged_left_padded_ids(b, s, seed=5)
    for _ in range(3):
        sc = dec.rerank_ids(ids, mask, 1837, 3082)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 20
    for _ in range(n):
        sc = dec.rerank_ids(ids, mask, 1837, 3082)
    torch.cuda.synchronize()
    out[f"rerank {b}x{s} ms"] = round((time.perf_counter() - t0) / n * 1e3, 3)
    out[f"rerank {b}x{s} scores"] = [round(float(x), 5) for x in sc[:3]]
print(json.dumps(out))
