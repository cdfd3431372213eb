"""Embedding client on the B200 encoder kernels - same duck type as the reference's
``EmbeddingClient`` (``src/lean_explore/util/embedding_client.py:16-106``): constructor
``(model_name, device=None, max_length=None, batch_size=None)``, attribute ``model_name``,
``async embed(texts, is_query=False) -> EmbeddingResponse{texts, embeddings, model}`` with the
blocking work run in the event loop's default executor (``:88-101``).
"""

from __future__ import annotations

import asyncio
import logging
import os

from pydantic import BaseModel

logger = logging.getLogger(__name__)

DEFAULT_BATCH_SIZE = 8  # embedding_client.py:13


class EmbeddingResponse(BaseModel):
    """Response from embedding generation (field-for-field the reference's model, :16-26)."""

    texts: list[str]
    embeddings: list[list[float]]
    model: str


class GpuEmbeddingClient:
    def __init__(self, model_name: str, device: str | None = None, max_length: int | None = None,
                 batch_size: int | None = None):
        self.model_name = model_name
        self.device = device or "cuda"
        if not str(self.device).startswith("cuda"):
            raise RuntimeError("GpuEmbeddingClient runs on a B200 only (there is no CPU fallback)")
        self.max_length = max_length
        self.batch_size = batch_size or int(os.getenv("LEAN_EXPLORE_EMBEDDING_BATCH_SIZE", DEFAULT_BATCH_SIZE))
        from .decoder import load_qwen3, model_type_of
        from .encoder import load_sentence_encoder

        logger.info("Loading embedding model %s on %s", model_name, self.device)
        if model_type_of(model_name) == "qwen3":  # the shipped Qwen/Qwen3-Embedding-0.6B (d = 1024)
            self.model = load_qwen3(model_name, device=self.device, max_length=max_length, with_lm_head=False)
        else:  # BERT-class sentence-transformers models (all-MiniLM-L6-v2, bge-base-en-v1.5)
            self.model = load_sentence_encoder(model_name, device=self.device, max_length=max_length)
        if max_length is not None:
            logger.info("Set max sequence length to %d", max_length)

    def embed_array(self, texts: list[str], is_query: bool = False):
        """Batched twin without the per-float Python lists of ``EmbeddingResponse``: float32
        ``[len(texts), d]`` straight from the encoder (what ``retrieve_semantic_candidates_batch``
        feeds the k-NN search with; a 1024 x 1024 batch spends ~50 ms in ``tolist()`` otherwise)."""
        return self.model.encode(texts, batch_size=self.batch_size, is_query=is_query)

    async def embed(self, texts: list[str], is_query: bool = False) -> EmbeddingResponse:
        loop = asyncio.get_event_loop()

        def _encode():
            # BERT-class sentence-transformers models define no "query" prompt: is_query is a
            # no-op for them; Qwen3-Embedding defines one and gets it prepended.
            return self.model.encode(texts, batch_size=self.batch_size, is_query=is_query)

        embeddings = await loop.run_in_executor(None, _encode)
        return EmbeddingResponse(texts=texts, embeddings=[emb.tolist() for emb in embeddings], model=self.model_name)
