"""ctypes binding of ``oracle/liblxoracle.so`` (the C restatement in ``flat_ip.c``).

CPU ORACLE - test infrastructure, not product code (see ``faiss_flat.py`` for who may import it).
"""

from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
SO = HERE / "liblxoracle.so"


def load() -> ctypes.CDLL:
    if not SO.exists() or SO.stat().st_mtime < (HERE / "flat_ip.c").stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE)], check=True, capture_output=True)
    lib = ctypes.CDLL(str(SO))
    vp, sz, i64 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int64
    lib.lxo_renorm_l2.argtypes = [sz, sz, vp]
    lib.lxo_heap_init.argtypes = [sz, sz, vp, vp]
    lib.lxo_heap_add_block.argtypes = [sz, sz, vp, vp, vp, sz, i64]
    lib.lxo_heap_finish.argtypes = [sz, sz, vp, vp]
    lib.lxo_knn_inner_product_seq.argtypes = [vp, vp, sz, sz, sz, sz, vp, vp]
    lib.lxo_num_threads.restype = ctypes.c_int
    lib.lxo_f64_topk_init.argtypes = [sz, sz, vp, vp]
    lib.lxo_f64_topk_add_rows.argtypes = [sz, sz, sz, vp, vp, ctypes.c_int, sz, i64, vp, vp]
    return lib


def renorm_l2(x: np.ndarray) -> None:
    """faiss fvec_renorm_L2 with fp32 SIMD-order accumulation (in place)."""
    assert x.dtype == np.float32 and x.flags.c_contiguous
    load().lxo_renorm_l2(x.shape[1], x.shape[0], x.ctypes.data)


def knn_inner_product_seq(x: np.ndarray, y: np.ndarray, k: int):
    """FAISS' nq < 20 path: per-pair fp32 dot products + CMin heap."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.ascontiguousarray(y, dtype=np.float32)
    D = np.empty((x.shape[0], k), dtype=np.float32)
    I = np.empty((x.shape[0], k), dtype=np.int64)
    load().lxo_knn_inner_product_seq(x.ctypes.data, y.ctypes.data, x.shape[1], x.shape[0], y.shape[0], k,
                                     D.ctypes.data, I.ctypes.data)
    return D, I


def knn_inner_product_blas(x: np.ndarray, y: np.ndarray, k: int, block: int = 1024):
    """FAISS' nq >= 20 path: blocked sgemm (numpy/OpenBLAS) + HeapBlockResultHandler."""
    lib = load()
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.ascontiguousarray(y, dtype=np.float32)
    nq = x.shape[0]
    D = np.empty((nq, k), dtype=np.float32)
    I = np.empty((nq, k), dtype=np.int64)
    lib.lxo_heap_init(nq, k, D.ctypes.data, I.ctypes.data)
    for j0 in range(0, y.shape[0], block):
        s = np.ascontiguousarray(x @ y[j0 : j0 + block].T)
        lib.lxo_heap_add_block(nq, k, D.ctypes.data, I.ctypes.data, s.ctypes.data, s.shape[1], j0)
    lib.lxo_heap_finish(nq, k, D.ctypes.data, I.ctypes.data)
    return D, I


class ExactTopK:
    """Streaming exact-arithmetic ranking (``lxo_f64_topk_add_rows``): feed the corpus block by block
    (fp16 or fp32 rows, ascending global row numbers), then ``result()`` gives what
    ``faiss_flat.flat_ip_search_f64`` gives on the whole matrix - (D float64 [nq, k], I int64 [nq, k]),
    best first, ``-FLT_MAX`` / ``-1`` padded."""

    def __init__(self, x: np.ndarray, k: int):
        self.lib = load()
        self.x = np.ascontiguousarray(x, dtype=np.float32)
        self.k = int(k)
        nq = self.x.shape[0]
        self.D = np.empty((nq, k), dtype=np.float64)
        self.I = np.empty((nq, k), dtype=np.int64)
        self.lib.lxo_f64_topk_init(nq, k, self.D.ctypes.data, self.I.ctypes.data)

    def add(self, rows: np.ndarray, row0: int) -> None:
        assert rows.dtype in (np.float16, np.float32) and rows.flags.c_contiguous and rows.shape[1] == self.x.shape[1]
        self.lib.lxo_f64_topk_add_rows(self.x.shape[0], self.x.shape[1], self.k, self.x.ctypes.data, rows.ctypes.data,
                                       int(rows.dtype == np.float16), rows.shape[0], int(row0), self.D.ctypes.data,
                                       self.I.ctypes.data)

    def merge(self, D: np.ndarray, I: np.ndarray) -> None:
        """Fold another partial result (e.g. another row shard's) into this one."""
        d = np.concatenate([self.D, D], axis=1)
        i = np.concatenate([self.I, I], axis=1)
        d = np.where(i < 0, -np.inf, d)
        order = np.lexsort((i, -d), axis=1)[:, : self.k]
        self.D = np.ascontiguousarray(np.take_along_axis(d, order, axis=1))
        self.I = np.ascontiguousarray(np.take_along_axis(i, order, axis=1))

    def result(self):
        D = self.D.copy()
        D[self.I < 0] = float(np.finfo(np.float32).min)
        return D, self.I.copy()
