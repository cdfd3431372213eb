"""CPU ORACLE of the sentence encoder - test infrastructure, not product code.

Restates what ``SentenceTransformer.encode`` computes for a BERT-class model inside the
reference's ``EmbeddingClient.embed`` (``src/lean_explore/util/embedding_client.py:88-101``):
HF ``transformers.BertModel`` forward (fp32) -> sentence-transformers ``Pooling`` (mean over
unmasked tokens with ``clamp(sum_mask, 1e-9)``, or the CLS token) -> ``Normalize``
(``F.normalize(p=2, dim=1)``).  sentence-transformers itself is not installable here; BertModel's
source is the installed ``transformers`` package.  No checkpoints exist offline, so the weights
are seeded random initialisations of the real geometries (MiniLM-L6: L6/H384/12 heads/FFN1536;
bge-base: L12/H768/12 heads/FFN3072) - **parity is unpinned by the reference**, which holds no
golden embeddings (``tests/util/embedding_client_test.py:124-140`` only checks length 384).
"""

from __future__ import annotations

import numpy as np
import torch

GEOMETRIES = {
    "tiny": dict(hidden_size=128, num_hidden_layers=2, num_attention_heads=4, intermediate_size=256),
    "minilm-l6": dict(hidden_size=384, num_hidden_layers=6, num_attention_heads=12, intermediate_size=1536),
    "bge-base": dict(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072),
}


def make_model(geometry: str, seed: int = 0, vocab_size: int = 30522, max_pos: int = 512, init_std: float = 0.05):
    """Seeded random BertModel (fp32, eval).  init_std is larger than BERT's 0.02 so that
    attention and LayerNorm are exercised away from the near-linear regime."""
    from transformers import BertConfig, BertModel

    cfg = BertConfig(vocab_size=vocab_size, max_position_embeddings=max_pos, initializer_range=init_std,
                     **GEOMETRIES[geometry])
    torch.manual_seed(seed)
    model = BertModel(cfg, add_pooling_layer=False).eval()
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():  # non-trivial biases and LayerNorm affine parameters
        for name, p in model.named_parameters():
            if name.endswith("bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
            elif "LayerNorm.weight" in name:
                p.copy_(1.0 + torch.randn(p.shape, generator=g) * 0.1)
    return model, cfg


def make_inputs(batch: int, seq: int, vocab_size: int = 30522, seed: int = 0):
    """Synthetic token ids with ragged right padding (no tokenizer vocab exists offline)."""
    rng = np.random.default_rng(seed)
    ids = rng.integers(1000, vocab_size, size=(batch, seq)).astype(np.int32)
    lens = rng.integers(min(seq, max(2, seq // 4)), seq + 1, size=batch)
    lens[0] = seq
    mask = (np.arange(seq)[None, :] < lens[:, None]).astype(np.int32)
    ids[:, 0] = 101
    ids = np.where(mask == 1, ids, 0).astype(np.int32)
    return ids, mask


@torch.no_grad()
def encode(model, ids: np.ndarray, mask: np.ndarray, pool: str = "mean") -> np.ndarray:
    out = model(input_ids=torch.from_numpy(ids).long(), attention_mask=torch.from_numpy(mask).long()).last_hidden_state
    if pool == "cls":
        emb = out[:, 0]
    else:
        m = torch.from_numpy(mask).to(out.dtype).unsqueeze(-1)
        emb = (out * m).sum(1) / torch.clamp(m.sum(1), min=1e-9)
    emb = torch.nn.functional.normalize(emb, p=2, dim=1)
    return emb.numpy().astype(np.float32)
