"""CPU ORACLE of the Qwen3 embedding model and reranker - test infrastructure, not product code.

Restates, with the installed ``transformers`` package (``models/qwen3/modeling_qwen3.py``), what
the reference computes through its two third-party model wrappers:

* ``EmbeddingClient.embed`` (``src/lean_explore/util/embedding_client.py:88-101``) ->
  ``SentenceTransformer("Qwen/Qwen3-Embedding-0.6B").encode``: ``Qwen3Model`` forward (fp32) ->
  sentence-transformers ``Pooling(lasttoken)`` -> ``Normalize``.  Pooling(lasttoken) takes column
  -1 when the batch is left padded and otherwise the last position whose mask is 1.
* ``RerankerClient._compute_scores_sync`` (``src/lean_explore/util/reranker_client.py:110-141``):
  ``Qwen3ForCausalLM`` logits of the LAST position, ``log_softmax([false, true])[1].exp()``.

No checkpoints exist offline, so weights are seeded random initialisations of the real geometry
(Qwen3-0.6B: L28 / H1024 / 16 q heads, 8 kv heads x 128 / FFN 3072; vocabulary cut to 4096 rows -
it only feeds an embedding gather) - **parity is unpinned by the reference**, which holds no golden
embeddings or reranker scores (``tests/util/reranker_client_test.py`` mocks the model).
"""

from __future__ import annotations

import numpy as np
import torch

GEOMETRIES = {
    "tiny": dict(hidden_size=256, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                 intermediate_size=512),
    "small": dict(hidden_size=512, num_hidden_layers=4, num_attention_heads=8, num_key_value_heads=2,
                  intermediate_size=1088),
    "qwen3-0.6b": dict(hidden_size=1024, num_hidden_layers=28, num_attention_heads=16, num_key_value_heads=8,
                       intermediate_size=3072),
}


def make_model(geometry: str, seed: int = 0, vocab_size: int = 4096, init_std: float = 0.05, causal_lm: bool = True):
    """Seeded random Qwen3ForCausalLM (fp32, eval, tied lm_head as Qwen3-*-0.6B ship)."""
    from transformers import Qwen3Config, Qwen3ForCausalLM

    cfg = Qwen3Config(vocab_size=vocab_size, head_dim=128, max_position_embeddings=32768, rope_theta=1e6,
                      rms_norm_eps=1e-6, tie_word_embeddings=True, initializer_range=init_std,
                      attention_bias=False, **GEOMETRIES[geometry])
    torch.manual_seed(seed)
    model = Qwen3ForCausalLM(cfg).eval()
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():  # non-trivial RMSNorm gains
        for name, p in model.named_parameters():
            if name.endswith("norm.weight") or "layernorm.weight" in name:
                p.copy_(1.0 + torch.randn(p.shape, generator=g) * 0.1)
    return model, cfg


def make_inputs(batch: int, seq: int, vocab_size: int = 4096, seed: int = 0, side: str = "left"):
    """Synthetic token ids with ragged padding on `side` (both reference clients pad on the left)."""
    rng = np.random.default_rng(seed)
    ids = rng.integers(10, vocab_size, size=(batch, seq)).astype(np.int32)
    lens = rng.integers(min(seq, max(1, seq // 4)), seq + 1, size=batch)
    lens[0] = seq
    pos = np.arange(seq)[None, :]
    mask = (pos >= seq - lens[:, None]) if side == "left" else (pos < lens[:, None])
    mask = mask.astype(np.int32)
    ids = np.where(mask == 1, ids, 0).astype(np.int32)
    return ids, mask


def _last_index(mask: torch.Tensor) -> torch.Tensor:
    # sentence-transformers Pooling(lasttoken): left padded -> -1, else index of the last 1
    s = mask.shape[1]
    return s - 1 - torch.flip(mask, dims=[1]).argmax(dim=1)


@torch.no_grad()
def embed(model, ids: np.ndarray, mask: np.ndarray) -> np.ndarray:
    m = torch.from_numpy(mask).long()
    out = model.model(input_ids=torch.from_numpy(ids).long(), attention_mask=m).last_hidden_state
    emb = out[torch.arange(out.shape[0]), _last_index(m)]
    emb = torch.nn.functional.normalize(emb, p=2, dim=1)
    return emb.numpy().astype(np.float32)


@torch.no_grad()
def rerank(model, ids: np.ndarray, mask: np.ndarray, token_true: int, token_false: int) -> np.ndarray:
    """reranker_client.py:124-139 verbatim in meaning: last position (left padded inputs)."""
    m = torch.from_numpy(mask).long()
    logits = model(input_ids=torch.from_numpy(ids).long(), attention_mask=m).logits[:, -1, :]
    stacked = torch.stack([logits[:, token_false], logits[:, token_true]], dim=1)
    return torch.nn.functional.log_softmax(stacked, dim=1)[:, 1].exp().numpy().astype(np.float32)
