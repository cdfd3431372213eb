"""FAISS-shaped flat inner-product index whose arithmetic runs in ``liblxg.so`` on a B200.

Mirrors what the reference touches on a ``faiss.Index`` (``src/lean_explore/search/engine.py``
lines 151-161 and 242-250, ``tests/extract/index_test.py:171-173``): ``.d``, ``.ntotal``,
``.add(x)``, ``.search(x, k) -> (D float32[nq,k], I int64[nq,k])`` and the free function
``normalize_L2(x)``.  PyTorch is used only to own device memory (the corpus tensor, device
outputs) and streams; every computation is a call through the C ABI of ``include/lxg.h``.
"""

from __future__ import annotations

import ctypes
from ctypes import c_void_p

import numpy as np
import torch

from . import _lib

MAX_K = 2048  # lxg_search returns LXG_EUNSUPPORTED above this (include/lxg.h)


def _current_stream_ptr(device: int) -> int:
    return int(torch.cuda.current_stream(device).cuda_stream)


def normalize_L2(x: np.ndarray, device: int = 0) -> None:
    """In-place ``faiss.normalize_L2`` (reference call site engine.py:242) for a host array."""
    if not isinstance(x, np.ndarray) or x.dtype != np.float32 or x.ndim != 2 or not x.flags.c_contiguous:
        raise TypeError("normalize_L2 expects a C-contiguous 2-D float32 numpy array")
    lib = _lib.init(device)
    _lib.check(lib.lxg_normalize_l2(x.ctypes.data, x.shape[0], x.shape[1], None))


class GpuIndexFlatIP:
    """Exact inner-product index on one GPU (``IndexFlatIP`` semantics, no training).

    The corpus is kept as ONE torch tensor in HBM, fp16 (default, what the B200 path streams)
    or fp32 (config 1 of BASELINE.json; an fp16 scan copy is derived, exact scores still come
    from the fp32 rows).  Row ``i`` is label ``i + row_offset``.
    """

    def __init__(self, d: int, dtype: str = "float16", device: int = 0, row_offset: int = 0):
        if dtype not in ("float16", "float32"):
            raise ValueError("dtype must be 'float16' or 'float32'")
        self.d = int(d)
        self.device = int(device)
        self.row_offset = int(row_offset)
        self._torch_dtype = torch.float16 if dtype == "float16" else torch.float32
        self._lib = _lib.init(self.device)
        self._corpus: torch.Tensor | None = None
        self._handle = c_void_p()
        self._pending: list[torch.Tensor] = []

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_tensor(cls, corpus: torch.Tensor, row_offset: int = 0) -> "GpuIndexFlatIP":
        """Wrap an existing [N, d] CUDA tensor (fp16 or fp32) without copying it."""
        if corpus.dim() != 2 or not corpus.is_cuda or not corpus.is_contiguous():
            raise ValueError("corpus must be a contiguous 2-D CUDA tensor")
        if corpus.dtype not in (torch.float16, torch.float32):
            raise ValueError("corpus dtype must be float16 or float32")
        dev = corpus.device.index if corpus.device.index is not None else torch.cuda.current_device()
        ix = cls(corpus.shape[1], "float16" if corpus.dtype == torch.float16 else "float32", dev, row_offset)
        ix._corpus = corpus
        ix._create()
        return ix

    def add(self, x) -> None:
        """``faiss.Index.add``: append rows (numpy or torch, any float dtype; stored in the
        index dtype).  The matrix is what extract/index.py:59-71 builds - it is NOT normalised."""
        t = torch.as_tensor(x)
        if t.dim() != 2 or t.shape[1] != self.d:
            raise ValueError(f"add expects [n, {self.d}]")
        self._pending.append(t.to(device=f"cuda:{self.device}", dtype=self._torch_dtype))
        self._destroy()

    def _materialise(self) -> None:
        if self._pending:
            parts = ([self._corpus] if self._corpus is not None else []) + self._pending
            self._corpus = torch.cat(parts, dim=0).contiguous()
            self._pending = []
        if self._corpus is None:
            self._corpus = torch.empty((0, self.d), dtype=self._torch_dtype, device=f"cuda:{self.device}")
        if not self._handle:
            self._create()

    def _create(self) -> None:
        assert self._corpus is not None
        torch.cuda.synchronize(self.device)
        h = c_void_p()
        dtype = _lib.LXG_F16 if self._corpus.dtype == torch.float16 else _lib.LXG_F32
        ptr = self._corpus.data_ptr() if self._corpus.shape[0] > 0 else None
        _lib.check(self._lib.lxg_index_create(ctypes.byref(h), ptr, self._corpus.shape[0], self.d, dtype,
                                              self.row_offset))
        self._handle = h

    def _destroy(self) -> None:
        if self._handle:
            self._lib.lxg_index_destroy(self._handle)
            self._handle = c_void_p()

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    # --------------------------------------------------------------------- properties
    @property
    def ntotal(self) -> int:
        n = 0 if self._corpus is None else self._corpus.shape[0]
        return n + sum(p.shape[0] for p in self._pending)

    @property
    def corpus(self) -> torch.Tensor:
        self._materialise()
        return self._corpus

    # ------------------------------------------------------------------------- search
    def search(self, x: np.ndarray, k: int, normalize: bool = False):
        """``faiss.Index.search`` on host arrays (engine.py:250): x float32 [nq, d] ->
        (D float32 [nq, k], I int64 [nq, k]).  ``normalize=True`` fuses the preceding
        ``faiss.normalize_L2`` (engine.py:242) into the kernel prologue; x is not modified."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 2 or x.shape[1] != self.d:
            raise ValueError(f"search expects [nq, {self.d}] float32")
        if k <= 0:
            raise ValueError("k must be positive")
        if k > MAX_K:
            raise ValueError(f"k = {k}: this index returns at most {MAX_K} neighbours per query (FAISS has no such limit; "
                             "the engine's faiss_k defaults to 1000, engine.py:538)")
        self._materialise()
        nq = x.shape[0]
        # results land in page-locked memory (torch's caching host allocator): the library then
        # copies device -> host straight into them; a pinned `x` is likewise read without a bounce
        D = torch.empty((nq, k), dtype=torch.float32, pin_memory=True).numpy()
        I = torch.empty((nq, k), dtype=torch.int64, pin_memory=True).numpy()
        _lib.check(self._lib.lxg_search(self._handle, x.ctypes.data, nq, k, int(normalize),
                                        D.ctypes.data, I.ctypes.data, _current_stream_ptr(self.device)))
        return D, I

    def search_torch(self, x: torch.Tensor, k: int, normalize: bool = False, out=None,
                     want_f64: bool = False):
        """Device-resident variant: x float32 CUDA [nq, d]; returns CUDA tensors (D, I[, D64]),
        asynchronous on the current stream (``sync()`` reports what a synchronous call would)."""
        if not x.is_cuda or x.dtype != torch.float32 or x.dim() != 2 or x.shape[1] != self.d:
            raise ValueError(f"search_torch expects a float32 CUDA tensor [nq, {self.d}]")
        if k <= 0:
            raise ValueError("k must be positive")
        if k > MAX_K:
            raise ValueError(f"k = {k}: this index returns at most {MAX_K} neighbours per query")
        x = x.contiguous()
        self._materialise()
        nq = x.shape[0]
        if out is None:
            D = torch.empty((nq, k), dtype=torch.float32, device=x.device)
            I = torch.empty((nq, k), dtype=torch.int64, device=x.device)
        else:
            D, I = out
        D64 = torch.empty((nq, k), dtype=torch.float64, device=x.device) if want_f64 else None
        _lib.check(self._lib.lxg_search_ex(self._handle, x.data_ptr(), nq, k, int(normalize), D.data_ptr(),
                                           I.data_ptr(), D64.data_ptr() if want_f64 else None,
                                           _current_stream_ptr(self.device)))
        return (D, I, D64) if want_f64 else (D, I)

    def search_packed(self, x: torch.Tensor, k: int, normalize: bool = False, packed: torch.Tensor | None = None,
                      scratch_d: torch.Tensor | None = None) -> torch.Tensor:
        """Row-shard form of ``search_torch``: the result lands in ONE contiguous int64 tensor
        ``packed[2, nq, k]`` - plane 0 the bits of the exact fp64 scores, plane 1 the global ids -
        which is exactly this rank's contribution to the all-gather (``lxg_merge_topk_packed``
        reads the gathered buffer in place).  Asynchronous on the current stream."""
        if not x.is_cuda or x.dtype != torch.float32 or x.dim() != 2 or x.shape[1] != self.d:
            raise ValueError(f"search_packed expects a float32 CUDA tensor [nq, {self.d}]")
        x = x.contiguous()
        self._materialise()
        nq = x.shape[0]
        if packed is None:
            packed = torch.empty((2, nq, k), dtype=torch.int64, device=x.device)
        if scratch_d is None:
            scratch_d = torch.empty((nq, k), dtype=torch.float32, device=x.device)
        assert packed.is_contiguous() and packed.shape == (2, nq, k) and packed.dtype == torch.int64
        _lib.check(self._lib.lxg_search_ex(self._handle, x.data_ptr(), nq, k, int(normalize), scratch_d.data_ptr(),
                                           packed[1].data_ptr(), packed[0].data_ptr(),
                                           _current_stream_ptr(self.device)))
        return packed

    def sync(self) -> int:
        """Waits for the last (asynchronous) search on this index; raises ``LxgError`` (LXG_ETIES) if a
        query had more exact ties with its k-th score than can be represented, and returns the
        number of queries the exact path re-did."""
        self._materialise()
        n = ctypes.c_int32(0)
        _lib.check(self._lib.lxg_index_sync(self._handle, ctypes.byref(n)))
        return int(n.value)

    def last_stats(self) -> dict:
        st = _lib.SearchStats()
        _lib.check(self._lib.lxg_index_last_stats(self._handle, ctypes.byref(st)))
        return {name: getattr(st, name) for name, _ in st._fields_}

    def set_timing(self, enable: bool) -> None:
        """Record CUDA events around the kernels of every search (bench.py's roofline)."""
        self._materialise()
        _lib.check(self._lib.lxg_index_set_timing(self._handle, int(enable)))

    def get_timing(self) -> dict:
        """Sum of per-kernel device times (ms) since the last call; synchronises."""
        t = _lib.Timing()
        _lib.check(self._lib.lxg_index_get_timing(self._handle, ctypes.byref(t)))
        return {name: getattr(t, name) for name, _ in t._fields_}

    def debug_scores(self, x: torch.Tensor, normalize: bool = False):
        """Test hook: raw tensor-core scores of pass 1, un-scaled back to true units."""
        self._materialise()
        x = x.contiguous()
        nq = x.shape[0]
        scores = torch.zeros((nq, self._corpus.shape[0]), dtype=torch.float32, device=x.device)
        qscale = torch.zeros((nq,), dtype=torch.float32, device=x.device)
        sscale = ctypes.c_float(1.0)
        _lib.check(self._lib.lxg_debug_scores(self._handle, x.data_ptr(), nq, int(normalize), scores.data_ptr(),
                                              qscale.data_ptr(), ctypes.byref(sscale),
                                              _current_stream_ptr(self.device)))
        torch.cuda.synchronize(self.device)
        return scores / (qscale[:, None] * sscale.value)
