"""Semantic retrieval of lean-explore's local backend on the B200 path.

Mirror of the semantic half of ``SearchEngine`` (reference ``src/lean_explore/search/engine.py``):
same constructor keywords for the pieces it owns (``:53-63``), same lazy loading
(``_ensure_faiss_loaded`` ``:151-161``, ``embedding_client`` ``:127-137``), same
``FileNotFoundError`` text (``:120-125``), same ``_retrieve_semantic_candidates`` contract
(``:225-261``: embed -> fp32 [1, d] -> normalize_L2 -> search(faiss_k) -> skip -1 /
out-of-range labels -> max similarity per declaration id, info log line).  Added beside it:
``retrieve_semantic_candidates_batch`` - the nq > 1 entry point the reference lacks (its
call is hard-wired to one query, ``:237-238``) that BASELINE.json's QPS configs need.

BM25, RRF, dependency boost and the rerank blend around this path live in
``lean_explore_b200.hybrid`` (``HybridSearchEngine``); an unmodified reference engine gets this
path through ``lean_explore_b200.faiss_compat.install()`` + ``SearchEngine(embedding_client=...)``.
"""

from __future__ import annotations

import logging
from pathlib import Path

import numpy as np

from . import corpus as _corpus

logger = logging.getLogger(__name__)

FAISS_INDEX_FILENAME = "informalization_faiss.index"        # engine.py:94
FAISS_IDS_MAP_FILENAME = "informalization_faiss_ids_map.json"  # engine.py:97


class SemanticRetriever:
    def __init__(
        self,
        embedding_client=None,
        embedding_model_name: str = "Qwen/Qwen3-Embedding-0.6B",
        faiss_index_path: Path | None = None,
        faiss_ids_map_path: Path | None = None,
        base_path: Path | None = None,
    ):
        self._embedding_client = embedding_client
        self._embedding_model_name = embedding_model_name
        if base_path is None and (faiss_index_path is None or faiss_ids_map_path is None):
            raise ValueError("give base_path or both artefact paths")
        self._faiss_informal_path = Path(faiss_index_path or (Path(base_path) / FAISS_INDEX_FILENAME))
        self._faiss_informal_ids_path = Path(faiss_ids_map_path or (Path(base_path) / FAISS_IDS_MAP_FILENAME))
        self._faiss_informal_index = None
        self._faiss_informal_id_map: list[int] | None = None
        self._validate_paths()

    def _validate_paths(self) -> None:
        for path in (self._faiss_informal_path, self._faiss_informal_ids_path):
            if not path.exists():
                raise FileNotFoundError(
                    f"Required file not found at {path}. "
                    "Please run 'lean-explore data fetch' to download the data."
                )

    @property
    def embedding_client(self):
        """Lazily created, ``max_length=512`` as the reference does (engine.py:127-137)."""
        if self._embedding_client is None:
            from .embedding_client import GpuEmbeddingClient

            self._embedding_client = GpuEmbeddingClient(model_name=self._embedding_model_name, max_length=512)
        return self._embedding_client

    def _ensure_faiss_loaded(self) -> None:
        if self._faiss_informal_index is not None:
            return
        from . import faiss_compat

        logger.info("Loading FAISS index from %s", self._faiss_informal_path)
        self._faiss_informal_index = faiss_compat.read_index(str(self._faiss_informal_path))
        self._faiss_informal_id_map = _corpus.load_ids_map(self._faiss_informal_ids_path)

    @property
    def faiss_informal_index(self):
        self._ensure_faiss_loaded()
        return self._faiss_informal_index

    @property
    def faiss_informal_id_map(self) -> list[int]:
        self._ensure_faiss_loaded()
        return self._faiss_informal_id_map

    @staticmethod
    def _to_map(indices_row, distances_row, id_map) -> dict[int, float]:
        semantic_map: dict[int, float] = {}
        for idx, dist in zip(indices_row, distances_row):
            if idx == -1 or idx >= len(id_map):
                continue
            decl_id = id_map[idx]
            semantic_map[decl_id] = max(semantic_map.get(decl_id, 0.0), float(dist))
        return semantic_map

    async def _retrieve_semantic_candidates(self, query: str, faiss_k: int) -> dict[int, float]:
        embedding_response = await self.embedding_client.embed([query], is_query=True)
        query_embedding = np.array([embedding_response.embeddings[0]], dtype=np.float32)
        informal_index = self.faiss_informal_index
        informal_id_map = self.faiss_informal_id_map
        if hasattr(informal_index, "nprobe"):
            informal_index.nprobe = 64
        # faiss.normalize_L2 (engine.py:242) is fused into the search kernel's prologue
        distances, indices = informal_index.search(query_embedding, faiss_k, normalize=True)
        semantic_map = self._to_map(indices[0], distances[0], informal_id_map)
        logger.info("FAISS informal: %d candidates", len(semantic_map))
        return semantic_map

    async def retrieve_semantic_candidates_batch(self, queries: list[str], faiss_k: int) -> list[dict[int, float]]:
        """nq > 1 twin: one encoder call and one search for the whole batch."""
        if not queries:
            return []
        client = self.embedding_client
        if hasattr(client, "embed_array"):  # GpuEmbeddingClient: numpy straight from the encoder
            import asyncio

            x = await asyncio.get_event_loop().run_in_executor(None, lambda: client.embed_array(list(queries), is_query=True))
            x = np.ascontiguousarray(x, dtype=np.float32)
        else:  # any other client with the reference's duck type
            embedding_response = await client.embed(list(queries), is_query=True)
            x = np.array(embedding_response.embeddings, dtype=np.float32)
        distances, indices = self.faiss_informal_index.search(x, faiss_k, normalize=True)
        id_map = self.faiss_informal_id_map
        return [self._to_map(indices[i], distances[i], id_map) for i in range(len(queries))]

    def search_embeddings(self, x: np.ndarray, faiss_k: int):
        """Raw batched twin for callers that already hold query embeddings:
        float32 [nq, d] -> (D float32 [nq, k], I int64 [nq, k]) labels, FAISS contract."""
        return self.faiss_informal_index.search(np.ascontiguousarray(x, dtype=np.float32), faiss_k, normalize=True)
