"""CPU suite: the row-sharded search's exchange step over a real process group (gloo,
world_size 2).  The per-shard search and the merge are injected with host restatements, so
what is exercised is the product's own plumbing in ``lean_explore_b200/sharded.py``: shard
bounds, global-id offsets, packing, the single all-gather, and the merge order."""

import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, d, nq, k, out_dir):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from conftest import make_corpus, make_queries
        from lean_explore_b200.sharded import ShardedFlatIP, merge_topk_host, shard_rows
        from oracle import faiss_flat as ff

        corpus = make_corpus(n, d, dtype=np.float32)
        corpus[n - 3] = corpus[1]  # an exact tie that straddles the two shards
        x = make_queries(nq, d)
        lo, hi = shard_rows(n, world, rank)

        def local_search(xt, kk, normalize):
            xn = xt.numpy().copy()
            if normalize:
                ff.normalize_L2(xn)
            dd, ii = ff.flat_ip_search_f64(corpus[lo:hi], xn, kk)
            ii = np.where(ii >= 0, ii + lo, -1)
            # this rank's contribution to the all-gather: [2, nq, k] 8-byte words (fp64 bits | ids)
            return torch.from_numpy(np.stack([np.ascontiguousarray(dd).view(np.int64), ii]))

        def merge(gathered, kk):
            g = gathered.numpy()  # [world, 2, nq, k]
            dd, ii = merge_topk_host(np.ascontiguousarray(g[:, 0]).view(np.float64), np.ascontiguousarray(g[:, 1]), kk)
            return torch.from_numpy(dd), torch.from_numpy(ii)

        eng = ShardedFlatIP(None, world, rank, local_search=local_search, merge=merge)
        D, I = eng.search_torch(torch.from_numpy(x), k, normalize=True)
        np.savez(Path(out_dir) / f"rank{rank}.npz", D=D.numpy(), I=I.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,k", [(1001, 10), (7, 12)])
def test_sharded_search_over_gloo_matches_single_index(tmp_path, n, k):
    sys.path.insert(0, str(ROOT / "tests"))
    from conftest import make_corpus, make_queries
    from oracle import faiss_flat as ff

    world, d, nq = 2, 48, 9
    mp.spawn(_worker, args=(world, _free_port(), n, d, nq, k, str(tmp_path)), nprocs=world, join=True)
    corpus = make_corpus(n, d, dtype=np.float32)
    corpus[n - 3] = corpus[1]
    x = make_queries(nq, d)
    ff.normalize_L2(x)
    D64, I64 = ff.flat_ip_search_f64(corpus, x, k)
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        assert np.array_equal(z["I"], I64), f"rank {r}"
        live = I64 >= 0
        assert np.allclose(z["D"][live], D64[live].astype(np.float32))
        assert (z["D"][~live] == ff.NEG_FLT_MAX).all()
