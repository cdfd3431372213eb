"""Reranker client on the B200 decoder kernels - same duck type as the reference's
``RerankerClient`` (``src/lean_explore/util/reranker_client.py:18-205``): constructor
``(model_name, device=None, max_length=512, instruction=DEFAULT_INSTRUCTION, batch_size=None)``,
``rerank_sync(query, documents) -> RerankerResponse`` and ``async rerank(query, documents,
batch_size=None)`` with the same batching rule (small inputs run inline, larger ones in the
event loop's default executor), scores in input order.
"""

from __future__ import annotations

import asyncio
import logging
import os

from pydantic import BaseModel

logger = logging.getLogger(__name__)

DEFAULT_INSTRUCTION = "Find relevant Lean 4 math declarations"  # reranker_client.py:13
DEFAULT_CUDA_BATCH_SIZE = 16  # reranker_client.py:14


class RerankerResponse(BaseModel):
    """Response from reranking operation (field-for-field the reference's model, :18-28)."""

    query: str
    scores: list[float]
    model: str


class GpuRerankerClient:
    def __init__(self, model_name: str = "Qwen/Qwen3-Reranker-0.6B", device: str | None = None, max_length: int = 512,
                 instruction: str = DEFAULT_INSTRUCTION, batch_size: int | None = None, model=None):
        self.model_name = model_name
        self.device = device or "cuda"
        if not str(self.device).startswith("cuda"):
            raise RuntimeError("GpuRerankerClient runs on a B200 only (there is no CPU fallback)")
        self.max_length = max_length
        self.instruction = instruction
        env_batch_size = os.getenv("LEAN_EXPLORE_RERANKER_BATCH_SIZE")
        if batch_size is not None:
            self.batch_size = batch_size
        elif env_batch_size:
            self.batch_size = int(env_batch_size)
        else:
            self.batch_size = DEFAULT_CUDA_BATCH_SIZE
        if model is None:
            from .decoder import load_qwen3

            logger.info("Loading reranker model %s on %s", model_name, self.device)
            model = load_qwen3(model_name, device=self.device, max_length=max_length, with_lm_head=True)
        self.model = model
        self.model.max_length = max_length
        self.tokenizer = model.tokenizer
        # token ids for true/false classification (reranker_client.py:85-86)
        self._token_true_id = self.tokenizer.convert_tokens_to_ids("true")
        self._token_false_id = self.tokenizer.convert_tokens_to_ids("false")

    def _format_pair(self, query: str, document: str) -> str:
        return f"<Instruct>: {self.instruction}\n<Query>: {query}\n<Document>: {document}"

    def _compute_scores_sync(self, pairs: list[str]) -> list[float]:
        return self.model.score_pairs(pairs, self._token_true_id, self._token_false_id)

    def rerank_sync(self, query: str, documents: list[str]) -> RerankerResponse:
        if not documents:
            return RerankerResponse(query=query, scores=[], model=self.model_name)
        pairs = [self._format_pair(query, doc) for doc in documents]
        return RerankerResponse(query=query, scores=self._compute_scores_sync(pairs), model=self.model_name)

    async def rerank(self, query: str, documents: list[str], batch_size: int | None = None) -> RerankerResponse:
        if not documents:
            return RerankerResponse(query=query, scores=[], model=self.model_name)
        if batch_size is None:
            batch_size = self.batch_size
        if len(documents) <= batch_size:
            return self.rerank_sync(query, documents)
        pairs = [self._format_pair(query, doc) for doc in documents]
        loop = asyncio.get_event_loop()
        all_scores: list[float] = []
        for i in range(0, len(pairs), batch_size):
            batch = pairs[i : i + batch_size]
            all_scores.extend(await loop.run_in_executor(None, self._compute_scores_sync, batch))
        return RerankerResponse(query=query, scores=all_scores, model=self.model_name)
