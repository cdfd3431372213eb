"""Row-sharded flat inner-product search across the GPUs of one node.

No reference analogue (the reference is a single process): used only
when the corpus exceeds one GPU (and for the strong-scaling sweep of BASELINE.json config 4).
One process per GPU (torchrun env), rank g holds the contiguous rows ``shard_rows(n, world, g)``
as its own ``GpuIndexFlatIP`` with ``row_offset = lo`` so per-shard ids are already global.
Per batch of queries (replicated on every rank) there is exactly ONE collective:

    local search   lxg_search_ex writes this rank's packed [2, nq, k] block (exact fp64 score
                   bits | int64 global ids = 16 bytes per candidate) - no repacking kernels
    exchange       dist.all_gather_into_tensor(gathered[world, 2, nq, k], packed)
    merge          lxg_merge_topk_packed reads the gathered buffer in place on every rank

Top-k is decomposable, so the result equals the single-index search bit for bit (scores are
merged in fp64, ties by ascending id).  With ``timing=True`` CUDA events bracket the three phases
on the launching stream (bench.py prints them as local_ms / collective_ms / merge_ms).
"""

from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_rows(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous row block [lo, hi) of rank `rank`: ceil(n / world) rows each, last may be short."""
    per = -(-n // world)
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def merge_topk_host(dg: np.ndarray, ig: np.ndarray, k: int):
    """Host restatement of lxg_merge_topk for plumbing tests of the collective (gloo, no GPU):
    dg float64 [shards, nq, k], ig int64 [shards, nq, k] (-1 padded)."""
    shards, nq, _ = dg.shape
    d = np.transpose(dg, (1, 0, 2)).reshape(nq, shards * k)
    i = np.transpose(ig, (1, 0, 2)).reshape(nq, shards * k)
    d = np.where(i < 0, -np.inf, d)
    order = np.lexsort((i, -d), axis=1)[:, :k]
    out_d = np.take_along_axis(d, order, axis=1)
    out_i = np.take_along_axis(i, order, axis=1)
    dead = out_i < 0
    out_d = out_d.astype(np.float32)
    out_d[dead] = np.finfo(np.float32).min
    return out_d, out_i


class ShardedFlatIP:
    """Search over a row-sharded corpus; `local_index` is this rank's ``GpuIndexFlatIP``.

    `local_search(x, k, normalize) -> packed int64 [2, nq, k]` and
    `merge(gathered int64 [world, 2, nq, k], k) -> (D, I)` can be injected (CPU plumbing tests
    under gloo); by default they are the CUDA path: ``lxg_search_ex`` and ``lxg_merge_topk_packed``.
    """

    def __init__(self, local_index, world: int, rank: int, group=None, local_search=None, merge=None,
                 timing: bool = False):
        self.index = local_index
        self.world = int(world)
        self.rank = int(rank)
        self.group = group
        self._local_search = local_search or self._gpu_local_search
        self._merge = merge or self._gpu_merge
        self._buffers: dict = {}
        self.timing = bool(timing) and local_search is None
        self._events: list = []

    # ------------------------------------------------------------------ CUDA defaults
    def _bufs(self, nq: int, k: int, device):
        """Per (nq, k) buffers, reused from call to call: this rank's packed block, the gathered
        buffer and the local fp32 scores (unused by the merge).  Reuse is safe: calls on one index
        are ordered on the device."""
        key = (nq, k)
        b = self._buffers.get(key)
        if b is None:
            if len(self._buffers) >= 8:
                self._buffers.clear()
            b = (torch.empty((2, nq, k), dtype=torch.int64, device=device),
                 torch.empty((self.world, 2, nq, k), dtype=torch.int64, device=device),
                 torch.empty((nq, k), dtype=torch.float32, device=device))
            self._buffers[key] = b
        return b

    def _gpu_local_search(self, x: torch.Tensor, k: int, normalize: bool):
        packed, _, scratch = self._bufs(x.shape[0], k, x.device)
        return self.index.search_packed(x, k, normalize=normalize, packed=packed, scratch_d=scratch)

    def _gpu_merge(self, gathered: torch.Tensor, k: int):
        from . import _lib
        from .index import _current_stream_ptr

        shards, _, nq, _ = gathered.shape
        out_d = torch.empty((nq, k), dtype=torch.float32, device=gathered.device)
        out_i = torch.empty((nq, k), dtype=torch.int64, device=gathered.device)
        lib = _lib.init(self.index.device)
        _lib.check(lib.lxg_merge_topk_packed(gathered.data_ptr(), nq, k, shards, out_d.data_ptr(), out_i.data_ptr(),
                                             _current_stream_ptr(self.index.device)))
        return out_d, out_i

    # ------------------------------------------------------------------------ search
    def search_torch(self, x: torch.Tensor, k: int, normalize: bool = False):
        """x: [nq, d] float32, identical on every rank.  Returns (D, I) on every rank."""
        ev = None
        if self.timing:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
        packed = self._local_search(x, k, normalize)  # [2, nq, k] int64
        nq = packed.shape[1]
        if ev:
            ev[1].record()
        if self.world == 1:
            gathered = packed.unsqueeze(0)
        else:
            if self._local_search == self._gpu_local_search:
                gathered = self._bufs(nq, k, packed.device)[1]
            else:
                gathered = torch.empty((self.world, 2, nq, k), dtype=torch.int64, device=packed.device)
            # (gloo wants the concatenated shape [world * 2, nq, k]; the bytes are the same)
            nvtx = packed.is_cuda
            if nvtx:
                torch.cuda.nvtx.range_push("ShardedFlatIP: all_gather of the per-shard top-k")
            dist.all_gather_into_tensor(gathered.view(self.world * 2, nq, k), packed, group=self.group)  # the single exchange step
            if nvtx:
                torch.cuda.nvtx.range_pop()
        if ev:
            ev[2].record()
        out = self._merge(gathered, k)
        if ev:
            ev[3].record()
            self._events.append(ev)
        return out

    def pop_timing(self) -> dict:
        """Sums of the per-phase device times (ms) since the last call; synchronises."""
        t = {"calls": len(self._events), "local_ms": 0.0, "collective_ms": 0.0, "merge_ms": 0.0}
        for ev in self._events:
            ev[3].synchronize()
            t["local_ms"] += ev[0].elapsed_time(ev[1])
            t["collective_ms"] += ev[1].elapsed_time(ev[2])
            t["merge_ms"] += ev[2].elapsed_time(ev[3])
        self._events = []
        return t

    def search(self, x: np.ndarray, k: int, normalize: bool = False):
        """Host-array API (numpy in, numpy out), same contract as ``GpuIndexFlatIP.search``: the
        queries go up with one async copy (straight from the caller's array when it is page-locked),
        the results come back into page-locked arrays, one synchronise at the end."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        if self.index is None:  # injected host stand-ins (CPU plumbing tests)
            d, i = self.search_torch(torch.from_numpy(x), k, normalize=normalize)
            return d.numpy(), i.numpy()
        dev = self.index.corpus.device
        d, i = self.search_torch(torch.from_numpy(x).to(dev, non_blocking=True), k, normalize=normalize)
        D = torch.empty(d.shape, dtype=torch.float32, pin_memory=True)
        I = torch.empty(i.shape, dtype=torch.int64, pin_memory=True)
        D.copy_(d, non_blocking=True)
        I.copy_(i, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        self.index.sync()  # surfaces LXG_ETIES of the asynchronous local search
        return D.numpy(), I.numpy()
