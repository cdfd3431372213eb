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
out = {k: os.environ.get(k) for k in ("LXG_ATTN_TC", "LXG_DECODER_PACK", "LXG_GEMM_NARROW", "LXG_QUERY_GEMM") if os.environ.get(k)}
for b, s in ((1, 24), (1, 12), (1, 32)):  # the query path: one short text
    ids, mask = ragged_left_padded_ids(b, s, seed=5)
    mask[:] = 1
    for _ in range(5):
        v = dec.embed_ids(ids, mask)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 200
    for _ in range(n):
        v = dec.embed_ids(ids, mask)
    torch.cuda.synchronize()
    out[f"embed {b}x{s} ms"] = round((time.perf_counter() - t0) / n * 1e3, 4)
    out[f"embed {b}x{s} v"] = [round(float(x), 5) for x in v[0, :3]]
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
