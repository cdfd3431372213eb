"""GPU parity tests of the row-sharded search: lxg_search_ex writing the packed
per-shard block, the exchange, lxg_merge_topk_packed - against the CPU oracle over the WHOLE corpus.

* two shards on ONE GPU (no collective): the packed blocks are concatenated the way the all-gather
  lays them out - exercises the kernels and layouts on every box;
* ShardedFlatIP with world = 1;
* two ranks over NCCL (spawned; skipped when fewer than two GPUs are visible) - the product's own
  exchange step, ids compared with the oracle on every rank, including an exact tie that straddles
  the shards."""

import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from conftest import make_corpus, make_queries
from oracle import faiss_flat as ff

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _case(n, d, nq, dtype=np.float16):
    corpus = make_corpus(n, d, dtype=dtype)
    corpus[n - 3] = corpus[1]          # an exact tie that straddles the shards
    if n > 20:
        corpus[n // 2 + 5] = corpus[n // 2 - 5]
    x = make_queries(nq, d)
    x[0] = corpus[1].astype(np.float32)  # a query whose best two rows tie exactly across the shards
    return corpus, x


def _oracle(corpus, x, k):
    xn = x.copy()
    ff.normalize_L2(xn)
    return ff.flat_ip_search_f64(corpus, xn, k)


@pytest.mark.parametrize("n,d,nq,k,shards", [(6000, 256, 40, 25, 2), (9001, 384, 130, 50, 3), (5, 64, 3, 12, 2),
                                              (4000, 128, 7, 1000, 4)])
def test_packed_shards_on_one_gpu_match_the_oracle(n, d, nq, k, shards):
    from lean_explore_b200 import GpuIndexFlatIP, _lib
    from lean_explore_b200.sharded import shard_rows

    corpus, x = _case(n, d, nq)
    ct = torch.from_numpy(corpus).cuda()
    xt = torch.from_numpy(x).cuda()
    gathered = torch.empty((shards, 2, nq, k), dtype=torch.int64, device="cuda")
    keep = []
    for s in range(shards):
        lo, hi = shard_rows(n, shards, s)
        ix = GpuIndexFlatIP.from_tensor(ct[lo:hi].contiguous(), row_offset=lo)
        ix.search_packed(xt, k, normalize=True, packed=gathered[s])
        keep.append(ix)
    D = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    I = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    lib = _lib.init(0)
    _lib.check(lib.lxg_merge_topk_packed(gathered.data_ptr(), nq, k, shards, D.data_ptr(), I.data_ptr(), None))
    torch.cuda.synchronize()
    D64, I64 = _oracle(corpus, x, k)
    assert np.array_equal(I.cpu().numpy(), I64)
    live = I64 >= 0
    assert np.abs(D.cpu().numpy()[live] - D64[live]).max() < 1e-3
    assert (D.cpu().numpy()[~live] == ff.NEG_FLT_MAX).all()
    # the legacy two-array entry point reads the same data through strides
    Dg = gathered[:, 0].contiguous().view(torch.float64)
    Ig = gathered[:, 1].contiguous()
    D2 = torch.empty_like(D)
    I2 = torch.empty_like(I)
    _lib.check(lib.lxg_merge_topk(Dg.data_ptr(), Ig.data_ptr(), nq, k, shards, D2.data_ptr(), I2.data_ptr(), None))
    torch.cuda.synchronize()
    assert torch.equal(I, I2) and torch.equal(D, D2)


def test_sharded_engine_world_one():
    from lean_explore_b200 import GpuIndexFlatIP
    from lean_explore_b200.sharded import ShardedFlatIP

    corpus, x = _case(30000, 384, 200)
    eng = ShardedFlatIP(GpuIndexFlatIP.from_tensor(torch.from_numpy(corpus).cuda()), 1, 0, timing=True)
    D, I = eng.search(x, 50, normalize=True)
    D64, I64 = _oracle(corpus, x, 50)
    assert np.array_equal(I, I64) and np.abs(D - D64).max() < 1e-3
    t = eng.pop_timing()
    assert t["calls"] == 1 and t["local_ms"] > 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _nccl_worker(rank, world, port, n, d, nq, k, out_dir):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from lean_explore_b200 import GpuIndexFlatIP
        from lean_explore_b200.sharded import ShardedFlatIP, shard_rows

        corpus, x = _case(n, d, nq)
        lo, hi = shard_rows(n, world, rank)
        ix = GpuIndexFlatIP.from_tensor(torch.from_numpy(corpus[lo:hi]).cuda(), row_offset=lo)
        eng = ShardedFlatIP(ix, world, rank, timing=True)
        for _ in range(3):  # buffers are reused from call to call
            D, I = eng.search(x, k, normalize=True)
        Dt, It = eng.search_torch(torch.from_numpy(x).cuda(), k, normalize=True)
        torch.cuda.synchronize()
        assert np.array_equal(It.cpu().numpy(), I) and np.array_equal(Dt.cpu().numpy(), D)
        t = eng.pop_timing()
        np.savez(Path(out_dir) / f"rank{rank}.npz", D=D, I=I, collective_ms=t["collective_ms"])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,d,nq,k", [(40001, 384, 130, 50), (9, 64, 3, 12)])
def test_two_ranks_over_nccl_match_the_oracle(tmp_path, n, d, nq, k):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp

    world = 2
    mp.spawn(_nccl_worker, args=(world, _free_port(), n, d, nq, k, str(tmp_path)), nprocs=world, join=True)
    corpus, x = _case(n, d, nq)
    D64, I64 = _oracle(corpus, x, k)
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        assert np.array_equal(z["I"], I64), f"rank {r}: {(z['I'] != I64).sum()} ids differ"
        live = I64 >= 0
        assert np.abs(z["D"][live] - D64[live]).max() < 1e-3
        assert (z["D"][~live] == ff.NEG_FLT_MAX).all()
    assert I64[0][0] == 1 and I64[0][1] == n - 3  # the tie straddles the shards, lower id first
