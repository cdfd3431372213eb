"""Encoder forward for ncu launch lists: python scripts/gpu_encoder_prof.py GEOM B S"""
import sys
import torch
sys.path.insert(0, ".")
from transformers import BertConfig, BertModel
from lean_explore_b200.encoder import POOL_MEAN, BertSentenceEncoder
geom, b, s = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
G = {"minilm": dict(hidden_size=384, num_hidden_layers=6, num_attention_heads=12, intermediate_size=1536),
     "bge": dict(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072)}[geom]
torch.manual_seed(0)
model = BertModel(BertConfig(vocab_size=30522, max_position_embeddings=512, **G), add_pooling_layer=False).eval()
enc = BertSentenceEncoder(model.state_dict(), hidden=G["hidden_size"], layers=G["num_hidden_layers"],
                          heads=G["num_attention_heads"], ffn=G["intermediate_size"], pool=POOL_MEAN)
dev = torch.device("cuda", 0)
ids = torch.randint(1000, 30000, (b, s), device=dev, dtype=torch.int32)
mask = torch.ones((b, s), dtype=torch.int32, device=dev)
for _ in range(2):
    enc.encode_ids_torch(ids, mask)
torch.cuda.synchronize()
