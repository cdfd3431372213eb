"""Reranker client on the B200 decoder kernels.

Drop-in for the reference's ``RerankerClient`` (``src/lean_explore/util/reranker_client.py:31-205``):
the engine only needs ``await client.rerank(query, documents)`` -> an object with ``.scores`` in
input order (``search/engine.py:386-387``); the constructor keywords, ``rerank_sync``, the
``_format_pair`` prompt layout (``:97-108``), the batching rule (``:164-205``: up to ``batch_size``
documents are scored inline, more go through the event loop's default executor batch by batch) and
the ``LEAN_EXPLORE_RERANKER_BATCH_SIZE`` override (``:59-67``) are kept so callers and tests written
against the reference class work unchanged.  The model forward + true/false softmax
(``:110-141``) is ``lxg_decoder_rerank``; there is no CPU fallback.
"""

from __future__ import annotations

import asyncio
import logging
import os

from pydantic import BaseModel

logger = logging.getLogger(__name__)

DEFAULT_INSTRUCTION = "Find relevant Lean 4 math declarations"  # reranker_client.py:13
DEFAULT_CUDA_BATCH_SIZE = 16                                    # reranker_client.py:14
_PROMPT_FIELDS = ("<Instruct>: ", "<Query>: ", "<Document>: ")  # one per line, reranker_client.py:106-108


class RerankerResponse(BaseModel):
    """Same three fields as the reference's response model (``reranker_client.py:18-28``)."""

    query: str
    scores: list[float]
    model: str


def _default_batch_size(explicit: int | None) -> int:
    if explicit is not None:
        return explicit
    from_env = os.getenv("LEAN_EXPLORE_RERANKER_BATCH_SIZE")
    return int(from_env) if from_env else DEFAULT_CUDA_BATCH_SIZE


class GpuRerankerClient:
    def __init__(self, model_name: str = "Qwen/Qwen3-Reranker-0.6B", device: str | None = None, max_length: int = 512,
                 instruction: str = DEFAULT_INSTRUCTION, batch_size: int | None = None, model=None):
        self.device = device or "cuda"
        if not str(self.device).startswith("cuda"):
            raise RuntimeError("GpuRerankerClient runs on a B200 only (there is no CPU fallback)")
        self.model_name, self.max_length, self.instruction = model_name, max_length, instruction
        self.batch_size = _default_batch_size(batch_size)
        if model is None:  # `model`: an already loaded Qwen3Decoder (tests, shared weights)
            from .decoder import load_qwen3

            logger.info("Loading reranker model %s on %s", model_name, self.device)
            model = load_qwen3(model_name, device=self.device, max_length=max_length, with_lm_head=True)
        model.max_length = max_length
        self.model, self.tokenizer = model, model.tokenizer
        # the two classification tokens whose last-position logits are compared (reranker_client.py:85-86)
        self._token_true_id, self._token_false_id = (self.tokenizer.convert_tokens_to_ids(t) for t in ("true", "false"))

    # ------------------------------------------------------------------ prompt + scoring
    def _format_pair(self, query: str, document: str) -> str:
        return "\n".join(field + value for field, value in zip(_PROMPT_FIELDS, (self.instruction, query, document)))

    def _compute_scores_sync(self, pairs: list[str]) -> list[float]:
        """P("true") over {"false", "true"} at the last position of every formatted pair, in [0, 1]."""
        return self.model.score_pairs(pairs, self._token_true_id, self._token_false_id)

    def _answer(self, query: str, scores: list[float]) -> RerankerResponse:
        return RerankerResponse(query=query, scores=scores, model=self.model_name)

    # ------------------------------------------------------------------ public calls
    def rerank_sync(self, query: str, documents: list[str]) -> RerankerResponse:
        scores = self._compute_scores_sync([self._format_pair(query, d) for d in documents]) if documents else []
        return self._answer(query, scores)

    async def rerank(self, query: str, documents: list[str], batch_size: int | None = None) -> RerankerResponse:
        step = self.batch_size if batch_size is None else batch_size
        if len(documents) <= step:  # also the empty list: no executor round trip for small inputs
            return self.rerank_sync(query, documents)
        prompts = [self._format_pair(query, d) for d in documents]
        loop = asyncio.get_event_loop()
        scores: list[float] = []
        for start in range(0, len(prompts), step):
            scores += await loop.run_in_executor(None, self._compute_scores_sync, prompts[start : start + step])
        return self._answer(query, scores)
