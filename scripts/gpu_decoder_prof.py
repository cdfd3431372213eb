"""One Qwen3-0.6B-geometry reranker forward (16 pairs x 256 tokens) for ncu launch lists / captures."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import qwen3_decoder as qd  # noqa: E402
from lean_explore_b200.decoder import Qwen3Decoder  # noqa: E402

b, s = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (16, 256)
model, cfg = qd.make_model("qwen3-0.6b", seed=0)
dec = Qwen3Decoder(model.state_dict(), hidden=cfg.hidden_size, layers=cfg.num_hidden_layers,
                   heads=cfg.num_attention_heads, kv_heads=cfg.num_key_value_heads, ffn=cfg.intermediate_size,
                   head_dim=cfg.head_dim, rms_eps=cfg.rms_norm_eps, rope_theta=1e6)
ids, mask = qd.make_inputs(b, s, seed=5)
print(dec.rerank_ids(ids, mask, 1837, 3082)[:4])
