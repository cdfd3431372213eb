"""B200-native semantic-search hot path of lean-explore (drop-in for the FAISS +
sentence-transformers arithmetic behind ``SearchEngine._retrieve_semantic_candidates``).

Importing the package is cheap; the CUDA library is loaded on first use and there is no CPU
fallback (see ``_lib.py``).
"""

__all__ = ["GpuIndexFlatIP", "normalize_L2"]


def __getattr__(name):
    if name in ("GpuIndexFlatIP", "normalize_L2"):
        from . import index

        return getattr(index, name)
    raise AttributeError(name)
