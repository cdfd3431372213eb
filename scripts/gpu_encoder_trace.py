"""Per-phase timeline of the single-kernel encoder forward (lxg_encoder_read_trace):
python scripts/gpu_encoder_trace.py [minilm-l6|bge-base] [B] [S]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from transformers import BertConfig, BertModel

from lean_explore_b200.encoder import POOL_MEAN, BertSentenceEncoder

G = {"minilm-l6": dict(hidden_size=384, num_hidden_layers=6, num_attention_heads=12, intermediate_size=1536),
     "bge-base": dict(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072)}
name = sys.argv[1] if len(sys.argv) > 1 else "minilm-l6"
b = int(sys.argv[2]) if len(sys.argv) > 2 else 1
s = int(sys.argv[3]) if len(sys.argv) > 3 else 16
g = G[name]
torch.manual_seed(0)
model = BertModel(BertConfig(vocab_size=30522, max_position_embeddings=512, **g), add_pooling_layer=False).eval()
enc = BertSentenceEncoder(model.state_dict(), hidden=g["hidden_size"], layers=g["num_hidden_layers"],
                          heads=g["num_attention_heads"], ffn=g["intermediate_size"], pool=POOL_MEAN)
dev = torch.device("cuda", 0)
ids = torch.randint(1000, 30000, (b, s), device=dev, dtype=torch.int32)
mask = torch.ones((b, s), dtype=torch.int32, device=dev)
enc.set_fused(2)
for _ in range(5):
    enc.encode_ids_torch(ids, mask)
torch.cuda.synchronize()
tr = enc.read_trace().astype(np.int64)  # [grid, phases, 6]
t0 = tr[:, 0, 0][tr[:, 0, 0] > 0].min()
nph = tr.shape[1]
print(f"{name} B={b} S={s}: grid {tr.shape[0]}, {nph} phases; total {(tr.max() - t0) / 1e3:.1f} us")
print("phase  jobs  barrier_release  wait->staged  staged->acc  acc->epi_done  epi->arrive   phase_span (us, medians over CTAs with a job)")
names = "ABCD"
prev_done = t0
for ph in range(nph - 1):
    has = tr[:, ph, 1] > 0
    if not has.any():
        continue
    x = tr[has, ph]
    start = x[:, 1].min()              # first CTA through the barrier
    done = tr[:, ph, 5].max()          # last arrival of the phase
    last_arrive_prev = tr[:, ph - 1, 5].max() if ph else t0
    med = lambda a: float(np.median(a)) / 1e3
    print(f"L{ph // 4}{names[ph % 4]}  {int(has.sum()):4d}  {(np.median(x[:, 1]) - last_arrive_prev) / 1e3:15.2f}  {med(x[:, 2] - x[:, 1]):12.2f}  "
          f"{med(x[:, 3] - x[:, 2]):11.2f}  {med(x[:, 4] - x[:, 3]):13.2f}  {med(x[:, 5] - x[:, 4]):11.2f}  {(done - last_arrive_prev) / 1e3:10.2f}")
print(f"pool: {(tr[0, nph - 1, 0] - tr[:, nph - 2, 5].max()) / 1e3:.2f} us after the last arrival")
