#!/usr/bin/env python
"""Benchmark of the semantic-search hot path: queries/sec for exact top-k over an N x d corpus.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg3|cfg1|cfg4]
                  [--queries Q] [--impl b200|reference]

A "step" is one batch of Q synthetic queries searched against the whole corpus (fused
normalise + fp16 tensor-core scan + exact re-score + top-k).  Prints ONE JSON line (rank 0).

  value     QPS with the query batch already resident in HBM (CUDA events, max over ranks)
  e2e       QPS through the public host API (numpy in / numpy out, copies inside the timing)
  roofline  the scan kernel's achieved TFLOP/s (or GB/s) against MEASURED_PEAKS.json
  cpu_baseline / --impl reference: the FAISS restatement (numpy sgemm + FAISS-style heaps,
            oracle/) on this box's host cores.
  extra     (default workload, one GPU) the query-batch sweep Q in {1, 8, 64, 256, 4096} and the
            engine's own request shape (one query, faiss_k = 1000); the same on the shape of the
            shipped index (400k x 1024 fp32); the BERT-class encoders and the Qwen3 embedding model /
            reranker through their host APIs, each with its CPU leg (HF fp32 on the host cores)

Workloads (BASELINE.json configs): cfg1 50k x 384 fp32 top-10 (Q=1000); cfg2 500k x 384 fp16
top-50 (default; Q=1024); cfg3 2M x 768 fp16 top-50; cfg4 16M x 768 fp16 row-sharded over the
ranks with an NCCL all-gather of per-shard candidates (strong scaling).
With --gpus N > 1 the default workload runs as N query-parallel replicas (each rank holds the
whole corpus and searches its own batches; no collective; weak scaling) - the corpus fits one
GPU, and the north star shards rows only when it does not.
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

# The reference arm times the CPU path "with all the host threads it can use": torchrun exports
# OMP_NUM_THREADS=1 to every rank, which would silently make it a one-core baseline.  Must happen
# before numpy (OpenBLAS) and the OpenMP oracle library are loaded.
if "reference" in sys.argv[1:] or "--impl=reference" in sys.argv[1:]:
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    "cfg1": dict(n=50_000, d=384, dtype="float32", k=10, q=1000,
                 name="cfg1: 1k-query batch, 50k x 384 fp32 corpus, top-10"),
    "cfg2": dict(n=500_000, d=384, dtype="float16", k=50, q=1024,
                 name="cfg2: Mathlib-scale 500k x 384 fp16 corpus, top-50"),
    "cfg3": dict(n=2_000_000, d=768, dtype="float16", k=50, q=1024,
                 name="cfg3: 2M x 768 fp16 corpus, top-50"),
    "cfg4": dict(n=16_000_000, d=768, dtype="float16", k=50, q=1024,
                 name="cfg4: 16M x 768 fp16 corpus row-sharded, top-50"),
}


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        j = json.loads(p.read_text())
        return dict(hbm_gbs=j["hbm_gbs"], tflops=j["bf16_tflops"],
                    tflops_sustained=j.get("bf16_tflops_sustained", j["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, tflops=1590.0, tflops_sustained=1400.0, source="fallback")


# --------------------------------------------------------------------------- synthetic data
def make_corpus_gpu(n, d, dtype, device, row0=0, seed=0):
    """BASELINE.md synthetic corpus: N(0,1) rows, L2-normalised in fp32, cast to `dtype`.
    Generated on the GPU in 64Ki-row chunks seeded by (seed, chunk) so any row range is
    reproducible on any rank."""
    import torch

    tdt = torch.float16 if dtype == "float16" else torch.float32
    out = torch.empty((n, d), dtype=tdt, device=device)
    chunk = 65536
    first = row0 // chunk
    last = (row0 + n + chunk - 1) // chunk
    for c in range(first, last):
        g = torch.Generator(device=device)
        g.manual_seed(seed * 1_000_003 + c)
        blk = torch.randn((chunk, d), generator=g, device=device, dtype=torch.float32)
        blk /= blk.norm(dim=1, keepdim=True)
        lo = max(row0, c * chunk)
        hi = min(row0 + n, (c + 1) * chunk)
        out[lo - row0 : hi - row0] = blk[lo - c * chunk : hi - c * chunk].to(tdt)
    return out


def make_queries_gpu(q, d, device, seed):
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(1_000_000_007 + seed)
    return torch.randn((q, d), generator=g, device=device, dtype=torch.float32)


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------- CPU reference
class CpuFlatIP:
    """The reference's CPU path for this hot path, restated (oracle/): faiss.normalize_L2 +
    IndexFlatIP.search = blocked sgemm (numpy/OpenBLAS, all cores) + FAISS-style per-query
    heaps (oracle/flat_ip.c, OpenMP, all cores).  bench.py is one of the three places allowed
    to execute oracle/ code, and only as the baseline being reported."""

    def __init__(self, corpus32: np.ndarray):
        import ctypes

        self.c = corpus32
        so = ROOT / "oracle" / "liblxoracle.so"
        if not so.exists():
            import subprocess

            subprocess.run(["make", "-C", str(ROOT / "oracle")], check=True)
        self.lib = ctypes.CDLL(str(so))
        vp, sz, i64 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int64
        self.lib.lxo_renorm_l2.argtypes = [sz, sz, vp]
        self.lib.lxo_heap_init.argtypes = [sz, sz, vp, vp]
        self.lib.lxo_heap_add_block.argtypes = [sz, sz, vp, vp, vp, sz, i64]
        self.lib.lxo_heap_finish.argtypes = [sz, sz, vp, vp]
        self.threads = int(self.lib.lxo_num_threads())

    def search(self, x: np.ndarray, k: int, block: int = 16384):
        x = np.ascontiguousarray(x, dtype=np.float32).copy()
        nq, d = x.shape
        self.lib.lxo_renorm_l2(d, nq, x.ctypes.data)
        D = np.empty((nq, k), dtype=np.float32)
        I = np.empty((nq, k), dtype=np.int64)
        self.lib.lxo_heap_init(nq, k, D.ctypes.data, I.ctypes.data)
        for j0 in range(0, self.c.shape[0], block):
            s = x @ self.c[j0 : j0 + block].T
            s = np.ascontiguousarray(s)
            self.lib.lxo_heap_add_block(nq, k, D.ctypes.data, I.ctypes.data, s.ctypes.data, s.shape[1], j0)
        self.lib.lxo_heap_finish(nq, k, D.ctypes.data, I.ctypes.data)
        return D, I


def time_cpu(cpu: CpuFlatIP, x: np.ndarray, k: int, budget_s: float):
    """QPS of the CPU path on a bounded sample: grow the query sample until ~budget_s."""
    nq = min(16, x.shape[0])
    cpu.search(x[:nq], k)  # warm-up (BLAS thread pool, page-in)
    while True:
        t0 = time.perf_counter()
        cpu.search(x[:nq], k)
        dt = time.perf_counter() - t0
        if dt > budget_s / 3 or nq >= x.shape[0]:
            return nq / dt, nq, dt
        nq = min(x.shape[0], max(nq * 2, int(nq * budget_s / 2 / max(dt, 1e-3))))


# --------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--queries", type=int, default=None)
    ap.add_argument("--k", type=int, default=None)
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU baseline work")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary cfg3 / sweep numbers")
    args = ap.parse_args()

    wl = dict(WORKLOADS[args.workload])
    if args.queries:
        wl["q"] = args.queries
    if args.k:
        wl["k"] = args.k
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world

    if args.impl == "reference":
        return run_reference(args, wl, rank)

    import torch
    import torch.distributed as dist

    from lean_explore_b200 import GpuIndexFlatIP

    assert torch.cuda.is_available(), "bench.py needs a B200; there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    sharded = args.workload == "cfg4"
    n, d, k, q = wl["n"], wl["d"], wl["k"], wl["q"]
    peaks = load_peaks()

    if sharded:
        from lean_explore_b200.sharded import ShardedFlatIP, shard_rows

        lo, hi = shard_rows(n, world, rank)
        corpus = make_corpus_gpu(hi - lo, d, wl["dtype"], dev, row0=lo)
        index = GpuIndexFlatIP.from_tensor(corpus, row_offset=lo)
        engine = ShardedFlatIP(index, world, rank)
        batches = [make_queries_gpu(q, d, dev, seed=s) for s in range(4)]  # same queries on every rank
        rows_local = hi - lo
    else:
        corpus = make_corpus_gpu(n, d, wl["dtype"], dev)
        index = GpuIndexFlatIP.from_tensor(corpus)
        engine = None
        batches = [make_queries_gpu(q, d, dev, seed=rank * 16 + s) for s in range(4)]
        rows_local = n
    D = torch.empty((q, k), dtype=torch.float32, device=dev)
    I = torch.empty((q, k), dtype=torch.int64, device=dev)

    def step(i):
        x = batches[i % len(batches)]
        if sharded:
            return engine.search_torch(x, k, normalize=True)
        return index.search_torch(x, k, normalize=True, out=(D, I))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(3, args.warmup)):
        step(i)
    barrier()

    sampler = ClockSampler(local_rank)
    index.set_timing(True)
    index.get_timing()
    with sampler:
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(args.steps):
            step(i)
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        tm = index.get_timing()
        index.set_timing(False)
        launches = index.last_stats()["kernel_launches"] * args.steps + (args.steps if sharded else 0)

        # end to end through the public host API: numpy in, numpy out, copies inside the timing
        e2e = None
        if not sharded:
            xh = [b.cpu().pin_memory().numpy() for b in batches]  # pinned host inputs (bench contract)
            index.search(xh[0], k, normalize=True)
            barrier()
            t0 = time.perf_counter()
            for i in range(args.steps):
                index.search(xh[i % len(xh)], k, normalize=True)
            torch.cuda.synchronize()
            e2e_s = time.perf_counter() - t0
        else:
            xh = [b.cpu().numpy() for b in batches]
            engine.search(xh[0], k, normalize=True)
            barrier()
            t0 = time.perf_counter()
            for i in range(args.steps):
                engine.search(xh[i % len(xh)], k, normalize=True)
            torch.cuda.synchronize()
            e2e_s = time.perf_counter() - t0

    if world > 1:
        t = torch.tensor([ms, e2e_s * 1e3], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(t[0]), float(t[1]) / 1e3
    units = q * args.steps * (1 if sharded else world)
    value = units / (ms / 1e3)
    e2e_value = units / e2e_s

    # roofline of the dominant kernel (pass-1 scan), per launch, from CUDA events on its stream
    scan_ms = tm["scan_ms"] / max(1, tm["calls"])
    flops = 2.0 * q * rows_local * d
    bytes_alg = rows_local * d * 2 + q * d * 4 + q * k * 12
    tf = flops / (scan_ms / 1e3) / 1e12
    gbs = bytes_alg / (scan_ms / 1e3) / 1e9
    tensor_frac, hbm_frac = tf / peaks["tflops"], gbs / peaks["hbm_gbs"]
    ridge_q = peaks["tflops"] * 1e12 / (peaks["hbm_gbs"] * 1e9)  # fp16: flop/byte == Q
    if q >= ridge_q:
        roof = dict(bound="tensor", achieved=round(tf, 2), peak=peaks["tflops"], unit="TFLOP/s",
                    frac=round(tensor_frac, 4), traffic=None)
    else:
        roof = dict(bound="hbm", achieved=round(gbs, 1), peak=peaks["hbm_gbs"], unit="GB/s",
                    frac=round(hbm_frac, 4), traffic=None)
    roof.update(kernel="scan_topk_kernel", ms_per_launch=round(scan_ms, 4), peak_source=peaks["source"],
                other_bound_frac=round(hbm_frac if roof["bound"] == "tensor" else tensor_frac, 4),
                prep_ms_per_launch=round(tm.get("prep_ms", 0.0) / max(1, tm["calls"]), 4),
                merge_ms_per_launch=round(tm["merge_ms"] / max(1, tm["calls"]), 4),
                exact_ms_per_launch=round(tm["exact_ms"] / max(1, tm["calls"]), 4))
    traffic_file = ROOT / "profiles" / "traffic.json"
    if traffic_file.exists():
        roof["traffic"] = json.loads(traffic_file.read_text()).get(args.workload)

    out = {
        "metric": "queries/sec top-50 over Nxd corpus" if k == 50 else f"queries/sec top-{k} over Nxd corpus",
        "value": round(value, 1), "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True,
        "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "f16xf16->f32 scan, f64 re-score",
        "data": "synthetic",
        "config": {"workload": wl["name"], "corpus_rows": n, "d": d, "corpus_dtype": wl["dtype"], "k": k,
                   "queries_per_step": q, "normalize": True,
                   "parallelism": ("row-sharded x%d + NCCL all-gather of per-shard top-k" % world) if sharded
                   else ("query-parallel replicas x%d" % world if world > 1 else "single GPU"),
                   "l2_policy": "corpus (%.0f MB) larger than L2; 4 rotating query batches" % (n * d * 2 / 1e6)
                   if n * d * 2 > 126e6 else "inputs fit L2 (config as specified by BASELINE.json)"},
        "e2e": {"value": round(e2e_value, 1), "unit": "queries/s", "h2d_bytes_per_step": q * d * 4,
                "d2h_bytes_per_step": q * k * 12, "api": "GpuIndexFlatIP.search(numpy) -> lxg_search",
                "host_buffers": "pinned; read / written in place by the kernels over PCIe (no staging copies)"},
        "gpu_launches": int(launches),
        "roofline": roof,
        "clocks": sampler.summary(),
    }

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        c32 = corpus.float().cpu().numpy()
        cpu = CpuFlatIP(c32)
        xs = torch.cat(batches).cpu().numpy()
        qps, nq_s, dt = time_cpu(cpu, xs, k, args.cpu_budget)
        out["cpu_baseline"] = {"value": round(qps, 1), "unit": "queries/s", "cores": cpu.threads, "kind": "port",
                               "sample": "%d queries x full %d x %d corpus (fp32), %.1f s; numpy sgemm + FAISS-style heaps"
                               % (nq_s, n, d, dt), "host_cpus": os.cpu_count()}
        # parity spot check of what was just timed
        Dg, Ig = index.search(xs[:nq_s], k, normalize=True)
        Dc, Ic = cpu.search(xs[:nq_s], k)
        out["cpu_baseline"]["ids_equal_frac"] = round(float((Ig == Ic).mean()), 6)
        del c32, cpu

    if rank == 0 and world == 1 and not args.no_extra and args.workload == "cfg2":
        out["extra"] = extra_numbers(index, d, k, dev, peaks)
        try:
            # the shape of the index the reference ships: Qwen3-Embedding vectors (d = 1024) stored as
            # fp32 (extract/index.py:59-71), ~400 k declarations; one query with faiss_k = 1000 is the
            # engine's own request, the 1024-query batch shows the d = 1024 scan variant's throughput
            index = None  # release the cfg2 corpus
            big = GpuIndexFlatIP.from_tensor(make_corpus_gpu(400_000, 1024, "float32", dev, seed=3))
            out["extra"]["shipped index shape: 400k x 1024 fp32"] = extra_numbers(big, 1024, k, dev, peaks,
                                                                                  sweep=((1, 1000), (64, 1000), (1024, 50)))
            del big
        except Exception as exc:  # noqa: BLE001
            out["extra"]["shipped index shape: 400k x 1024 fp32"] = {"error": repr(exc)}
        try:
            out["extra"]["encoder"] = encoder_numbers(dev, cpu=not args.no_cpu_baseline)
        except Exception as exc:  # noqa: BLE001 - secondary numbers must not lose the headline line
            out["extra"]["encoder"] = {"error": repr(exc)}
        try:
            out["extra"]["qwen3"] = decoder_numbers(dev, cpu=not args.no_cpu_baseline)
        except Exception as exc:  # noqa: BLE001
            out["extra"]["qwen3"] = {"error": repr(exc)}

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out), flush=True)


def extra_numbers(index, d, k, dev, peaks, sweep=None):
    """Query-batch sweep on the same corpus (Q in {1,8,64,256,1024,4096})."""
    import torch

    res = {}
    n = index.ntotal
    # (Q, k): the batch sweep at the workload's k, plus the reference's own request shape - one
    # query, faiss_k = 1000 candidates (SearchEngine.search default, engine.py:538)
    for q, k in (sweep or ((1, k), (8, k), (64, k), (256, k), (4096, k), (1, 1000))):
        xs = [make_queries_gpu(q, d, dev, seed=100 + s) for s in range(4)]
        for i in range(3):
            index.search_torch(xs[i], k, normalize=True)
        torch.cuda.synchronize()
        steps = 50
        index.set_timing(True)
        index.get_timing()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            index.search_torch(xs[i % 4], k, normalize=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        tm = index.get_timing()
        index.set_timing(False)
        scan = tm["scan_ms"] / steps
        label = f"Q={q}" if sweep is None else f"Q={q} k={k}"
        res[label if (q, k) != (1, 1000) or sweep is not None else "Q=1 k=1000 (engine default faiss_k)"] = {"qps": round(q / (ms / 1e3), 1), "ms_per_step": round(ms, 4), "scan_ms": round(scan, 4),
                         "prep_ms": round(tm.get("prep_ms", 0.0) / steps, 4), "merge_ms": round(tm["merge_ms"] / steps, 4),
                         "exact_ms": round(tm["exact_ms"] / steps, 4),
                         "hbm_frac": round(n * d * 2 / (scan / 1e3) / 1e9 / peaks["hbm_gbs"], 4),
                         "tensor_frac": round(2.0 * q * n * d / (scan / 1e3) / 1e12 / peaks["tflops"], 4)}
    return res


def encoder_numbers(dev, cpu: bool):
    """Row a2 of the hot path (SentenceTransformer.encode inside EmbeddingClient.embed): the
    sentence-encoder forward on random-init weights of the real geometries (no checkpoints exist
    offline), synthetic token ids.  Query path = one text per call (B=1); bulk path = the
    reference's corpus-embedding batches.  CPU leg: HF BertModel fp32 on the host cores through
    oracle/bert_encoder.py (what sentence-transformers runs), reference batch size 8."""
    import torch
    from transformers import BertConfig, BertModel

    from lean_explore_b200.encoder import POOL_CLS, POOL_MEAN, BertSentenceEncoder

    geoms = {"minilm-l6 (d=384)": (dict(hidden_size=384, num_hidden_layers=6, num_attention_heads=12,
                                        intermediate_size=1536), POOL_MEAN, "mean"),
             "bge-base (d=768)": (dict(hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                                       intermediate_size=3072), POOL_CLS, "cls")}
    res = {}
    for name, (g, pool, pool_name) in geoms.items():
        torch.manual_seed(0)
        model = BertModel(BertConfig(vocab_size=30522, max_position_embeddings=512, **g), add_pooling_layer=False).eval()
        enc = BertSentenceEncoder(model.state_dict(), hidden=g["hidden_size"], layers=g["num_hidden_layers"],
                                  heads=g["num_attention_heads"], ffn=g["intermediate_size"], pool=pool,
                                  device=dev.index or 0)
        entry = {}
        for label, b, sl, iters in (("query B=1 S=16", 1, 16, 200), ("bulk B=256 S=64", 256, 64, 20)):
            gen = torch.Generator(device=dev).manual_seed(5)
            ids = torch.randint(1000, 30000, (b, sl), generator=gen, device=dev, dtype=torch.int32)
            mask = torch.ones((b, sl), dtype=torch.int32, device=dev)
            for _ in range(3):
                enc.encode_ids_torch(ids, mask)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                enc.encode_ids_torch(ids, mask)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            flops = 2.0 * b * sl * g["num_hidden_layers"] * (4 * g["hidden_size"] ** 2 + 2 * g["hidden_size"] * g["intermediate_size"])
            entry[label] = {"ms_per_call": round(ms, 4), "sentences_per_s": round(b / (ms / 1e3), 1),
                            "gemm_tflops": round(flops / (ms / 1e3) / 1e12, 2), "launches": enc.last_launches()}
        # host API, one query text per call (numpy ids in, numpy vector out; copies inside the timing)
        ids_h = np.random.default_rng(0).integers(1000, 30000, (1, 16)).astype(np.int32)
        mask_h = np.ones((1, 16), dtype=np.int32)
        enc.encode_ids(ids_h, mask_h)
        t0 = time.perf_counter()
        for _ in range(200):
            enc.encode_ids(ids_h, mask_h)
        entry["query B=1 S=16"]["e2e_ms_per_call"] = round((time.perf_counter() - t0) / 200 * 1e3, 4)
        if cpu:
            from oracle import bert_encoder as be

            entry["cpu_baseline"] = {}
            for label, b, sl, reps in (("query B=1 S=16", 1, 16, 10), ("bulk B=8 S=64", 8, 64, 3)):
                ids_c = np.random.default_rng(1).integers(1000, 30000, (b, sl)).astype(np.int32)
                mask_c = np.ones((b, sl), dtype=np.int32)
                be.encode(model, ids_c, mask_c, pool_name)
                t0 = time.perf_counter()
                for _ in range(reps):
                    be.encode(model, ids_c, mask_c, pool_name)
                dt = (time.perf_counter() - t0) / reps
                entry["cpu_baseline"][label] = {"ms_per_call": round(dt * 1e3, 3), "sentences_per_s": round(b / dt, 1)}
            entry["cpu_baseline"]["kind"] = "HF BertModel fp32 (what sentence-transformers runs), torch %d threads" % torch.get_num_threads()
        res[name] = entry
        del enc, model
    return res


def qwen3_random_model(seed: int = 0, vocab_size: int = 4096):
    """Random-init HF Qwen3ForCausalLM of the 0.6B geometry (no checkpoints exist offline); only its
    state_dict is used on the GPU leg."""
    import torch
    from transformers import Qwen3Config, Qwen3ForCausalLM

    cfg = Qwen3Config(vocab_size=vocab_size, head_dim=128, max_position_embeddings=32768, rope_theta=1e6, rms_norm_eps=1e-6,
                      tie_word_embeddings=True, initializer_range=0.05, attention_bias=False, hidden_size=1024,
                      num_hidden_layers=28, num_attention_heads=16, num_key_value_heads=8, intermediate_size=3072)
    torch.manual_seed(seed)
    return Qwen3ForCausalLM(cfg).eval(), cfg


def ragged_left_padded_ids(batch: int, seq: int, vocab_size: int = 4096, seed: int = 0):
    """Synthetic token ids, lengths uniform in [seq/4, seq] (first row full), padded on the left."""
    rng = np.random.default_rng(seed)
    ids = rng.integers(10, vocab_size, size=(batch, seq)).astype(np.int32)
    lens = rng.integers(min(seq, max(1, seq // 4)), seq + 1, size=batch)
    lens[0] = seq
    mask = (np.arange(seq)[None, :] >= seq - lens[:, None]).astype(np.int32)
    return np.where(mask == 1, ids, 0).astype(np.int32), mask


def decoder_numbers(dev, cpu: bool):
    """The shipped models: Qwen3-Embedding-0.6B forward behind
    EmbeddingClient.embed and Qwen3-Reranker-0.6B behind RerankerClient._compute_scores_sync, on
    random-init weights of the real geometry (vocabulary cut to 4096 rows: it only feeds a gather),
    synthetic left-padded token ids with ragged lengths (uniform in [S/4, S]).  CPU leg: HF Qwen3ForCausalLM fp32 on the host cores through
    oracle/qwen3_decoder.py."""
    import torch

    from lean_explore_b200.decoder import Qwen3Decoder

    model, cfg = qwen3_random_model()
    dec = Qwen3Decoder(model.state_dict(), hidden=cfg.hidden_size, layers=cfg.num_hidden_layers,
                       heads=cfg.num_attention_heads, kv_heads=cfg.num_key_value_heads, ffn=cfg.intermediate_size,
                       head_dim=cfg.head_dim, rms_eps=cfg.rms_norm_eps, rope_theta=1e6, device=dev.index or 0)
    H, F, L = cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers
    qkv, c = (cfg.num_attention_heads + 2 * cfg.num_key_value_heads) * 128, cfg.num_attention_heads * 128
    flop_tok = 2.0 * L * (H * qkv + c * H + H * 2 * F + F * H)
    res = {"geometry": "Qwen3-0.6B: L28 H1024 16q/8kv x128 FFN3072, vocab rows cut to 4096", "launches": None}
    tt, tf = 1837, 3082
    for label, mode, b, sl, iters in (("embed query B=1 S=24", 0, 1, 24, 100), ("embed bulk B=64 S=128", 0, 64, 128, 10),
                                      ("rerank B=16 S=256 (reference CUDA batch)", 1, 16, 256, 10),
                                      ("rerank B=50 S=256 (rerank_top=50 in one call)", 1, 50, 256, 10)):
        ids_h, mask_h = ragged_left_padded_ids(b, sl, seed=5)
        call = (lambda: dec.embed_ids(ids_h, mask_h)) if mode == 0 else (lambda: dec.rerank_ids(ids_h, mask_h, tt, tf))
        for _ in range(3):
            call()
        t0 = time.perf_counter()
        for _ in range(iters):
            call()  # host ids in, host result out: H2D + forward + D2H + sync inside the timing
        ms = (time.perf_counter() - t0) / iters * 1e3
        computed = dec.last_tokens()  # host batches are packed: padding tokens never enter the layers
        res[label] = {"e2e_ms_per_call": round(ms, 4), "items_per_s": round(b / (ms / 1e3), 1),
                      "tokens_padded": b * sl, "tokens_computed": computed,
                      "gemm_tflops": round(flop_tok * computed / (ms / 1e3) / 1e12, 2)}
    res["launches"] = dec.last_launches()
    if cpu:
        from oracle import qwen3_decoder as qd  # the CPU leg is the only place the oracle runs

        res["cpu_baseline"] = {"kind": "HF Qwen3ForCausalLM fp32, torch %d threads" % torch.get_num_threads()}
        for label, mode, b, sl, reps in (("embed query B=1 S=24", 0, 1, 24, 3), ("rerank B=4 S=256", 1, 4, 256, 1)):
            ids_c, mask_c = ragged_left_padded_ids(b, sl, seed=6)
            fn = (lambda: qd.embed(model, ids_c, mask_c)) if mode == 0 else (lambda: qd.rerank(model, ids_c, mask_c, tt, tf))
            fn()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            dt = (time.perf_counter() - t0) / reps
            res["cpu_baseline"][label] = {"ms_per_call": round(dt * 1e3, 2), "items_per_s": round(b / dt, 2)}
    del dec, model
    return res


def run_reference(args, wl, rank):
    """--impl reference: the reference's CPU path for this hot path (FAISS restatement from
    oracle/) on the host cores, same config/metric; rank 0 only."""
    if rank != 0:
        return
    n, d, k, q = wl["n"], wl["d"], wl["k"], wl["q"]
    if args.workload == "cfg4":
        n = 2_000_000  # 16M x 768 fp32 does not fit host RAM: time a 2M slice, scale x1/8 below
    rng = np.random.default_rng(0)
    c32 = np.empty((n, d), dtype=np.float32)
    for j0 in range(0, n, 65536):
        blk = rng.standard_normal((min(65536, n - j0), d), dtype=np.float32)
        blk /= np.linalg.norm(blk, axis=1, keepdims=True)
        c32[j0 : j0 + blk.shape[0]] = blk.astype(np.float16).astype(np.float32) if wl["dtype"] == "float16" else blk
    cpu = CpuFlatIP(c32)
    xs = np.random.default_rng(1).standard_normal((q, d), dtype=np.float32)
    # bounded sample per step: as many queries as keep one step near 1 s
    qps0, nq_s, dt0 = time_cpu(cpu, xs, k, 3.0)
    nq_step = int(min(q, max(1, qps0 * 1.0)))
    for _ in range(min(args.warmup, 3)):
        cpu.search(xs[:nq_step], k)
    steps = max(1, min(args.steps, int(120.0 / max(nq_step / qps0, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu.search(xs[:nq_step], k)
    dt = time.perf_counter() - t0
    value = nq_step * steps / dt
    if args.workload == "cfg4":
        value /= 8.0
    out = {
        "impl": "reference", "metric": "queries/sec top-50 over Nxd corpus" if k == 50 else f"queries/sec top-{k} over Nxd corpus",
        "value": round(value, 1), "unit": "queries/s", "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 3),
        "ms_per_step": round(dt / steps * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "corpus_rows": wl["n"], "d": d, "corpus_dtype": wl["dtype"], "k": k,
                   "queries_per_step": nq_step, "normalize": True, "parallelism": "host CPU, all cores"},
        "cpu_baseline": {"value": round(value, 1), "unit": "queries/s", "cores": cpu.threads, "kind": "port",
                         "sample": "%d queries/step x %d steps over %d x %d fp32 rows%s; faiss-cpu is not installable "
                                   "offline, so this is its restatement (numpy/OpenBLAS sgemm + FAISS-style heaps)"
                                   % (nq_step, steps, n, d, " (2M-row slice, QPS scaled 1/8)" if args.workload == "cfg4" else ""),
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": round(value, 1), "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
