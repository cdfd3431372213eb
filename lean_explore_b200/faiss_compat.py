"""The slice of the ``faiss`` module API that lean-explore's local backend touches, served by
the B200 index.

The reference imports faiss lazily inside two methods (``search/engine.py:156`` and ``:240``)
and uses exactly: ``faiss.read_index(path)``, ``faiss.normalize_L2(x)``,
``index.search(x, k)``, an optional settable ``index.nprobe`` and, in its tests,
``index.ntotal`` / ``index.d`` (``tests/extract/index_test.py:171-173``).  ``install()``
registers this module as ``sys.modules["faiss"]`` so that an unmodified
``lean_explore.search.engine.SearchEngine`` runs its semantic retrieval on the GPU
(INTEGRATION.md).  Everything numeric goes through ``liblxg.so``; there is no CPU fallback.
"""

from __future__ import annotations

import os
import sys

import numpy as np

from . import corpus as _corpus
from .index import GpuIndexFlatIP, normalize_L2  # noqa: F401  (re-exported: faiss.normalize_L2)

METRIC_INNER_PRODUCT = _corpus.METRIC_INNER_PRODUCT
METRIC_L2 = _corpus.METRIC_L2


def _corpus_dtype() -> str:
    """LEAN_EXPLORE_CORPUS_DTYPE = float32 (default: scores are exactly those of the fp32 rows
    FAISS stores) or float16 (half the HBM, rows rounded once at load)."""
    v = os.getenv("LEAN_EXPLORE_CORPUS_DTYPE", "float32").lower()
    if v not in ("float32", "float16"):
        raise ValueError("LEAN_EXPLORE_CORPUS_DTYPE must be float32 or float16")
    return v


def _device() -> int:
    return int(os.getenv("LEAN_EXPLORE_GPU", "0"))


class IndexFlatIP(GpuIndexFlatIP):
    """``faiss.IndexFlatIP(d)``."""

    def __init__(self, d: int):
        super().__init__(d, dtype=_corpus_dtype(), device=_device())

    is_trained = True
    metric_type = METRIC_INNER_PRODUCT


def read_index(path: str) -> IndexFlatIP:
    """``faiss.read_index`` (engine.py:159).  Flat and IVFFlat files are both served as an exact
    flat index: the IVF search of the reference (nprobe=64, engine.py:247-248) returns a subset
    of what exhaustive search returns; the north star fixes IndexFlatIP as the semantics."""
    matrix, info = _corpus.read_index_matrix(path)
    if info["metric_type"] != METRIC_INNER_PRODUCT:
        raise ValueError("only inner-product indexes are supported")
    ix = IndexFlatIP(info["d"])
    if matrix.shape[0]:
        ix.add(matrix)
    return ix


def write_index(index, path: str) -> None:
    """``faiss.write_index`` for a flat index (extract/index.py:173)."""
    _corpus.write_flat_index(path, index.corpus.float().cpu().numpy())


def get_num_gpus() -> int:
    import torch

    return torch.cuda.device_count()


def install() -> None:
    """Make ``import faiss`` resolve to this module (the reference imports it lazily)."""
    sys.modules["faiss"] = sys.modules[__name__]
