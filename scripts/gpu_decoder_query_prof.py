"""One Qwen3-0.6B-geometry query embedding (B = 1, S = 24) for ncu launch lists."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bench import qwen3_random_model, ragged_left_padded_ids  # noqa: E402
from lean_explore_b200.decoder import Qwen3Decoder  # noqa: E402

model, cfg = qwen3_random_model()
dec = Qwen3Decoder(model.state_dict(), hidden=cfg.hidden_size, layers=cfg.num_hidden_layers,
                   heads=cfg.num_attention_heads, kv_heads=cfg.num_key_value_heads, ffn=cfg.intermediate_size,
                   head_dim=cfg.head_dim, rms_eps=cfg.rms_norm_eps, rope_theta=1e6)
ids, mask = ragged_left_padded_ids(1, 24, seed=5)
mask[:] = 1
print(dec.embed_ids(ids, mask)[0, :4], dec.last_launches())
