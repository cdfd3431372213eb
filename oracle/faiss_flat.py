"""CPU ORACLE - test infrastructure, not product code.

Restatement in numpy of the k-NN half of lean-explore's semantic-retrieval hot path:
``faiss.normalize_L2`` + ``faiss.IndexFlatIP.search`` as called from
``SearchEngine._retrieve_semantic_candidates`` (reference
``src/lean_explore/search/engine.py:238-258``) on the matrix built by
``src/lean_explore/extract/index.py:59-71``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference arm may
import this module.  The product path (``lean_explore_b200``) never does.

PARITY PINNING.  The arithmetic lives in the un-vendored wheel ``faiss-cpu`` (declared only as
``faiss-cpu>=1.7`` in the reference's ``pyproject.toml:36``; no lock file; not installable in
the build container, no network).  The reference's own tests hold exactly one known-answer
vector for this boundary (``tests/extract/index_test.py:185-205``: a one-hot row is its own
nearest neighbour) and no golden scores or ids.  This oracle is pinned against that KAT and
against FAISS' published semantics restated below; beyond that **parity is unpinned by the
reference** (see DESIGN.md).  What makes the comparison meaningful anyway: the ranking is
defined by exact arithmetic (``flat_ip_search_f64``), not by an implementation.

FAISS semantics restated (faiss/utils/distances.cpp, faiss/utils/Heap.h, faiss 1.7-1.12):

* ``fvec_renorm_L2``: per row ``nr = ||x||^2``; ``if nr > 0: inv_nr = 1.0 / sqrtf(nr)``
  (a double division rounded to float) and ``x[j] *= inv_nr`` in fp32.  All-zero rows are left
  untouched.  FAISS accumulates ``nr`` in fp32 in SIMD order; here it is the exactly rounded
  sum (fp64 accumulation, one rounding), which is what the CUDA path computes as well.
* ``IndexFlatIP.search`` -> ``knn_inner_product``: metric = raw inner product, larger is
  better; per query a size-k min-heap (``CMin<float,int64>``, k < 100) or a reservoir
  (k >= 100); a candidate enters only if ``ip > threshold`` (strict); results are emitted best
  first; unfilled slots keep ``D = -FLT_MAX (lowest float)``, ``I = -1``; NaN never enters.
* Order among EXACTLY equal scores is an artefact of FAISS internals (heap eviction order,
  reservoir partitioning, sgemm block boundaries; ``oracle/flat_ip.c`` restates the heap
  variant) and differs between FAISS' own code paths.  Oracle and CUDA path therefore fix
  it: ascending row id.  This is the one place where the restatement is a definition rather
  than a copy of behaviour; tests on tie-free data are unaffected.
* ``nq < 20`` uses per-pair SIMD dot products, ``nq >= 20`` blocked ``sgemm``: low-order bits of
  ``D`` depend on the BLAS, which is why the contract is "ids exact, scores to 1e-3".
"""

from __future__ import annotations

import numpy as np

NEG_FLT_MAX = np.float32(np.finfo(np.float32).min)


def normalize_L2(x: np.ndarray) -> None:
    """In-place ``faiss.normalize_L2`` (reference call site engine.py:242)."""
    if x.dtype != np.float32 or x.ndim != 2:
        raise TypeError("normalize_L2 expects a 2-D float32 array")
    nr = np.einsum("ij,ij->i", x.astype(np.float64), x.astype(np.float64)).astype(np.float32)
    nz = nr > 0
    inv = np.ones_like(nr)
    inv[nz] = (1.0 / np.sqrt(nr[nz]).astype(np.float64)).astype(np.float32)
    x *= inv[:, None]


def _topk_rows(scores: np.ndarray, k: int, base: int):
    """Per-row top-k of a score block, (score desc, column asc); returns (D, I) of width <= k."""
    n = scores.shape[1]
    kk = min(k, n)
    if kk < n:
        part = np.argpartition(-scores, kk - 1, axis=1)[:, :kk]
        # argpartition is not tie-stable: pull in every column equal to the k-th value
        kth = np.take_along_axis(scores, part, axis=1).min(axis=1)
        out_d = np.empty((scores.shape[0], kk), dtype=scores.dtype)
        out_i = np.empty((scores.shape[0], kk), dtype=np.int64)
        for r in range(scores.shape[0]):
            cols = np.flatnonzero(scores[r] >= kth[r])
            order = np.lexsort((cols, -scores[r, cols]))[:kk]
            out_i[r] = cols[order]
            out_d[r] = scores[r, cols[order]]
        return out_d, out_i + base
    order = np.lexsort((np.broadcast_to(np.arange(n), scores.shape), -scores), axis=1)
    return np.take_along_axis(scores, order, axis=1), order.astype(np.int64) + base


def _search(corpus: np.ndarray, x: np.ndarray, k: int, dtype, block: int):
    nq = x.shape[0]
    n = corpus.shape[0]
    best_d = np.full((nq, 0), 0, dtype=dtype)
    best_i = np.zeros((nq, 0), dtype=np.int64)
    xq = np.ascontiguousarray(x, dtype=dtype)
    for j0 in range(0, n, block):
        cb = np.ascontiguousarray(corpus[j0 : j0 + block]).astype(dtype, copy=False)
        s = xq @ cb.T
        s = np.where(np.isnan(s), -np.inf, s)  # NaN never enters a FAISS heap
        d_blk, i_blk = _topk_rows(s, k, j0)
        cat_d = np.concatenate([best_d, d_blk], axis=1)
        cat_i = np.concatenate([best_i, i_blk], axis=1)
        order = np.lexsort((cat_i, -cat_d), axis=1)[:, :k]
        best_d = np.take_along_axis(cat_d, order, axis=1)
        best_i = np.take_along_axis(cat_i, order, axis=1)
    out_d = np.full((nq, k), NEG_FLT_MAX, dtype=dtype)
    out_i = np.full((nq, k), -1, dtype=np.int64)
    w = best_d.shape[1]
    out_d[:, :w] = best_d
    out_i[:, :w] = best_i
    dead = ~(out_d > NEG_FLT_MAX)  # -inf / -FLT_MAX never beat the heap's neutral element
    out_i[dead] = -1
    out_d[dead] = NEG_FLT_MAX
    return out_d, out_i


def flat_ip_search(corpus: np.ndarray, x: np.ndarray, k: int, block: int = 65536):
    """``IndexFlatIP.search`` in fp32 (BLAS sgemm, like FAISS for nq >= 20).

    corpus: [N, d] float32 or float16 (up-cast to fp32, as a CPU FAISS index stores fp32);
    x: [nq, d] float32.  Returns (D float32[nq,k], I int64[nq,k]).
    """
    d, i = _search(corpus, x, k, np.float32, block)
    return d.astype(np.float32), i


def flat_ip_search_f64(corpus: np.ndarray, x: np.ndarray, k: int, block: int = 65536):
    """Same search with the inner products evaluated in fp64: the exact-arithmetic ranking
    the fp32 result approximates (products of fp32/fp16 inputs are exact in fp64)."""
    return _search(corpus, x, k, np.float64, block)


def ambiguous_positions(d64: np.ndarray, tol: float = 4e-6) -> np.ndarray:
    """Boolean [nq,k] mask of ranks whose fp64 score is within `tol` of a neighbouring rank -
    the places where two correct fp32 implementations (different BLAS / summation order) may
    legitimately disagree on the order."""
    gap_next = np.abs(np.diff(d64, axis=1))
    amb = np.zeros(d64.shape, dtype=bool)
    amb[:, :-1] |= gap_next < tol
    amb[:, 1:] |= gap_next < tol
    return amb


class IndexFlatIP:
    """Duck-type of ``faiss.IndexFlatIP`` (``add``, ``search``, ``ntotal``, ``d``) - the
    attributes the reference touches (engine.py:247-250, tests/extract/index_test.py:171-173)."""

    def __init__(self, d: int):
        self.d = int(d)
        self._rows = np.zeros((0, self.d), dtype=np.float32)

    @property
    def ntotal(self) -> int:
        return self._rows.shape[0]

    def add(self, x: np.ndarray) -> None:
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 2 or x.shape[1] != self.d:
            raise ValueError("add expects [n, d]")
        self._rows = np.concatenate([self._rows, x], axis=0)

    def search(self, x: np.ndarray, k: int):
        x = np.ascontiguousarray(x, dtype=np.float32)
        return flat_ip_search(self._rows, x, k)


def retrieve_semantic_candidates(index, id_map, query_embedding, faiss_k):
    """The glue of ``SearchEngine._retrieve_semantic_candidates`` after the embedding call
    (engine.py:238-258): fp32 cast, normalize_L2, search, skip -1 / out-of-range labels,
    keep the max similarity per declaration id."""
    q = np.array([query_embedding], dtype=np.float32)
    normalize_L2(q)
    distances, indices = index.search(q, faiss_k)
    semantic_map: dict[int, float] = {}
    for idx, dist in zip(indices[0], distances[0]):
        if idx == -1 or idx >= len(id_map):
            continue
        decl_id = id_map[idx]
        semantic_map[decl_id] = max(semantic_map.get(decl_id, 0.0), float(dist))
    return semantic_map
