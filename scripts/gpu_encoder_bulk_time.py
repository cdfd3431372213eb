"""Bulk BERT-class encoder forwards (B = 256, S = 64, random weights): ms per call, device ids in / device
vectors out (CUDA events), for A/B runs of two builds (LXG_LIB_PATH)."""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from transformers import BertConfig, BertModel  # noqa: E402

from lean_explore_b200.encoder import POOL_MEAN, BertSentenceEncoder  # noqa: E402

out = {"lib": os.path.basename(os.environ.get("LXG_LIB_PATH", "liblxg.so"))}
dev = torch.device("cuda", 0)
for geom, G in (("minilm", dict(hidden_size=384, num_hidden_layers=6, num_attention_heads=12, intermediate_size=1536)),
                ("bge", dict(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072))):
    torch.manual_seed(0)
    model = BertModel(BertConfig(vocab_size=30522, max_position_embeddings=512, **G), add_pooling_layer=False).eval()
    enc = BertSentenceEncoder(model.state_dict(), hidden=G["hidden_size"], layers=G["num_hidden_layers"],
                              heads=G["num_attention_heads"], ffn=G["intermediate_size"], pool=POOL_MEAN)
    ids = torch.randint(1000, 30000, (256, 64), device=dev, dtype=torch.int32)
    mask = torch.ones((256, 64), dtype=torch.int32, device=dev)
    for _ in range(3):
        v = enc.encode_ids_torch(ids, mask)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        v = enc.encode_ids_torch(ids, mask)
    e1.record()
    torch.cuda.synchronize()
    out[geom + " bulk 256x64 ms"] = round(e0.elapsed_time(e1) / 30, 4)
    out[geom + " v"] = [round(float(x), 5) for x in v[0, :3]]
print(json.dumps(out))
