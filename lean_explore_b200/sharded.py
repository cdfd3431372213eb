"""Row-sharded flat inner-product search across the GPUs of one node.

No reference analogue (the reference is a single process): used only
when the corpus exceeds one GPU.  One process per GPU (torchrun env), rank g holds the
contiguous rows ``shard_rows(n, world, g)`` as its own ``GpuIndexFlatIP`` with
``row_offset = lo`` so per-shard ids are already global.  Per batch of queries (replicated on
every rank) there is exactly ONE collective: an all-gather of the packed per-shard top-k
(exact fp64 score bits + int64 id = 16 bytes per candidate), followed by the merge kernel
``lxg_merge_topk`` on every rank.  Top-k is decomposable, so the result equals the
single-index search bit for bit (scores are merged in fp64, ties by ascending id).
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

    `local_search` / `merge` can be injected (CPU plumbing tests under gloo); by default they
    are the CUDA path: ``lxg_search_ex`` with fp64 scores and ``lxg_merge_topk``.
    """

    def __init__(self, local_index, world: int, rank: int, group=None, local_search=None, merge=None):
        self.index = local_index
        self.world = int(world)
        self.rank = int(rank)
        self.group = group
        self._local_search = local_search or self._gpu_local_search
        self._merge = merge or self._gpu_merge

    # ------------------------------------------------------------------ CUDA defaults
    def _gpu_local_search(self, x: torch.Tensor, k: int, normalize: bool):
        _, ids, d64 = self.index.search_torch(x, k, normalize=normalize, want_f64=True)
        return d64, ids

    def _gpu_merge(self, dg: torch.Tensor, ig: torch.Tensor, k: int):
        from . import _lib
        from .index import _current_stream_ptr

        shards, nq, _ = dg.shape
        out_d = torch.empty((nq, k), dtype=torch.float32, device=dg.device)
        out_i = torch.empty((nq, k), dtype=torch.int64, device=dg.device)
        lib = _lib.init(self.index.device)
        _lib.check(lib.lxg_merge_topk(dg.data_ptr(), ig.data_ptr(), nq, k, shards, out_d.data_ptr(),
                                      out_i.data_ptr(), _current_stream_ptr(self.index.device)))
        return out_d, out_i

    # ------------------------------------------------------------------------ search
    def search_torch(self, x: torch.Tensor, k: int, normalize: bool = False):
        """x: [nq, d] float32, identical on every rank.  Returns (D, I) on every rank."""
        d64, ids = self._local_search(x, k, normalize)
        nq = d64.shape[0]
        if self.world == 1:
            return self._merge(d64.unsqueeze(0).contiguous(), ids.unsqueeze(0).contiguous(), k)
        packed = torch.stack([d64.contiguous().view(torch.int64), ids], dim=0).contiguous()  # [2, nq, k]
        gathered = torch.empty((self.world * 2, nq, k), dtype=torch.int64, device=packed.device)
        dist.all_gather_into_tensor(gathered, packed, group=self.group)  # the single exchange step
        gathered = gathered.view(self.world, 2, nq, k)
        dg = gathered[:, 0].contiguous().view(torch.float64)
        ig = gathered[:, 1].contiguous()
        return self._merge(dg, ig, k)

    def search(self, x: np.ndarray, k: int, normalize: bool = False):
        """Host-array API (numpy in, numpy out), same contract as ``GpuIndexFlatIP.search``."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        dev = self.index.corpus.device if self.index is not None else torch.device("cpu")
        d, i = self.search_torch(torch.from_numpy(x).to(dev), k, normalize=normalize)
        return d.cpu().numpy(), i.cpu().numpy()
