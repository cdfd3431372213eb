"""One Qwen3-0.6B-geometry reranker forward (default 16 pairs x 256 tokens) for ncu launch lists / captures."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bench import qwen3_random_model, ragged_left_padded_ids  # noqa: E402
from lean_explore_b200.decoder import Qwen3Decoder  # noqa: E402

b, s = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (16, 256)
model, cfg = qwen3_random_model()
dec = Qwen3Decoder(model.state_dict(), hidden=cfg.hidden_size, layers=cfg.num_hidden_layers,
                   heads=cfg.num_attention_heads, kv_heads=cfg.num_key_value_heads, ffn=cfg.intermediate_size,
                   head_dim=cfg.head_dim, rms_eps=cfg.rms_norm_eps, rope_theta=1e6)
ids, mask = ragged_left_padded_ids(b, s, seed=5)
print(dec.rerank_ids(ids, mask, 1837, 3082)[:4])
