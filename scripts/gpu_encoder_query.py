"""Query-path encoder latency, single persistent kernel vs layered kernels:
python scripts/gpu_encoder_query.py   (prints one JSON object)"""
import json
import sys

import torch

sys.path.insert(0, ".")
from transformers import BertConfig, BertModel

from lean_explore_b200.encoder import POOL_MEAN, BertSentenceEncoder

G = {"minilm-l6": dict(hidden_size=384, num_hidden_layers=6, num_attention_heads=12, intermediate_size=1536),
     "bge-base": dict(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072)}
dev = torch.device("cuda", 0)
out = {}
for name, g in G.items():
    torch.manual_seed(0)
    model = BertModel(BertConfig(vocab_size=30522, max_position_embeddings=512, **g), add_pooling_layer=False).eval()
    enc = BertSentenceEncoder(model.state_dict(), hidden=g["hidden_size"], layers=g["num_hidden_layers"],
                              heads=g["num_attention_heads"], ffn=g["intermediate_size"], pool=POOL_MEAN)
    H, F, L = g["hidden_size"], g["intermediate_size"], g["num_hidden_layers"]
    wbytes = 2.0 * L * (4 * H * H + 2 * H * F)
    for b, s in ((1, 8), (1, 16), (1, 32), (1, 64), (4, 16)):
        ids = torch.randint(1000, 30000, (b, s), device=dev, dtype=torch.int32)
        mask = torch.ones((b, s), dtype=torch.int32, device=dev)
        row = {}
        for label, fused in (("fused", True), ("layered", False)):
            enc.set_fused(fused)
            for _ in range(5):
                enc.encode_ids_torch(ids, mask)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(200):
                enc.encode_ids_torch(ids, mask)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 200
            row[label] = {"us": round(ms * 1e3, 2), "launches": enc.last_launches(),
                          "weight_read_hbm_frac": round(wbytes / (ms / 1e3) / 6551e9, 4)}
        out[f"{name} B={b} S={s}"] = row
print(json.dumps(out, indent=1))
