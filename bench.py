#!/usr/bin/env python
"""Benchmark of the semantic-search hot path: queries/sec for exact top-k over an N x d corpus.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg4|cfg3|cfg2|cfg1]
                  [--queries Q] [--impl b200|reference]

A "step" is one batch of Q synthetic queries searched against the whole corpus (fused
normalise + fp16 tensor-core scan + exact re-score + top-k [+ all-gather + shard merge]).
Prints ONE JSON line (rank 0).

Default workload for EVERY --gpus N: BASELINE.json config 4 - a 16M x 768 fp16 corpus, top-50,
1024 queries per step, row-sharded over the N ranks through ShardedFlatIP (lxg_search_ex ->
ONE NCCL all-gather of the packed per-shard candidates -> lxg_merge_topk_packed).  The corpus is
24.6 GB, so N = 1 holds all of it and 1/2/4/8 is a STRONG-scaling sweep ("scaling": "strong").

  value     whole-job QPS with the query batch already resident in HBM (CUDA events, max over ranks)
  e2e       QPS through the host API (numpy in / numpy out, copies inside the timing), pinned and pageable
  phases    per-step device time of local search / all-gather / merge (CUDA events on the launching stream)
  roofline  the scan kernel's achieved TFLOP/s (or GB/s) against MEASURED_PEAKS.json
  sustained a >= 2 s loop of the same step against the SUSTAINED tensor peak, with the SM clock under load
  parity    ids of the first queries of the batch against the exact-arithmetic ranking over the FULL corpus
            (oracle/flat_ip.c lxo_f64_topk_add_rows, streamed from HBM; every rank checks its shard, rank 0 merges)
  cpu_baseline / --impl reference: the FAISS restatement (numpy sgemm + FAISS-style heaps, oracle/)
            on this box's host cores.
  extra     (N = 1 only) the same block for cfg3 (2M x 768, the north star's >= 10k QPS / >= 60 % target)
            and cfg2 (500k x 384), cfg1 (50k x 384 fp32, top-10 - the reference's own CPU-sized case), the query-batch
            sweep, the shape of the shipped index
            (400k x 1024 fp32, one query with faiss_k = 1000), the BERT-class encoders and the Qwen3
            embedding model / reranker through their host APIs, each with its CPU leg.

--workload cfg1|cfg2|cfg3 runs that config on one GPU (with --gpus N > 1: N query-parallel replicas,
no collective, weak scaling).
"""

from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys
import threading
import time
from pathlib import Path

_WORLD = int(os.environ.get("WORLD_SIZE", "1"))
if "reference" in sys.argv[1:] or "--impl=reference" in sys.argv[1:]:
    # The reference arm times the CPU path "with all the host threads it can use": torchrun exports
    # OMP_NUM_THREADS=1 to every rank, which would silently make it a one-core baseline.  Must happen
    # before numpy (OpenBLAS) and the OpenMP oracle library are loaded.
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)
elif _WORLD > 1:
    # the parity check runs the C oracle on every rank's shard: share the host cores between the ranks
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[_v] = str(max(1, (os.cpu_count() or 1) // _WORLD))

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    "cfg1": dict(n=50_000, d=384, dtype="float32", k=10, q=1000, check_q=64, cpu_q=4000,
                 name="cfg1: 1k-query batch, 50k x 384 fp32 corpus, top-10"),
    "cfg2": dict(n=500_000, d=384, dtype="float16", k=50, q=1024, check_q=64, cpu_q=4096,
                 name="cfg2: Mathlib-scale 500k x 384 fp16 corpus, top-50"),
    "cfg3": dict(n=2_000_000, d=768, dtype="float16", k=50, q=1024, check_q=64, cpu_q=1024,
                 name="cfg3: 2M x 768 fp16 corpus, top-50"),
    "cfg4": dict(n=16_000_000, d=768, dtype="float16", k=50, q=1024, check_q=16, cpu_q=256,
                 name="cfg4: 16M x 768 fp16 corpus row-sharded + NCCL all-gather of per-shard top-k, top-50"),
}


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        j = json.loads(p.read_text())
        return dict(hbm_gbs=j["hbm_gbs"], tflops=j["bf16_tflops"],
                    tflops_sustained=j.get("bf16_tflops_sustained", j["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, tflops=1590.0, tflops_sustained=1400.0, source="fallback")


def scan_source_sha():
    """Digest of the scan kernel's sources: a static ncu traffic figure is only quoted for the build it
    was captured on."""
    h = hashlib.sha256()
    for f in ("scan_topk.cuh", "ptx.cuh"):
        h.update((ROOT / "lean_explore_b200" / "csrc" / f).read_bytes())
    return h.hexdigest()[:12]


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
            time.sleep(0.005)

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
    to execute oracle/ code, and only as the baseline being reported / the checker.

    Streaming form: begin(x, k); add_block(rows fp32, j0) for ascending row blocks; finish().
    `seconds` accumulates the time spent in the reference's own arithmetic (renorm, sgemm, heaps)."""

    def __init__(self):
        import ctypes

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
        self.seconds = 0.0

    def begin(self, x: np.ndarray, k: int):
        t0 = time.perf_counter()
        self.x = np.ascontiguousarray(x, dtype=np.float32).copy()
        nq, d = self.x.shape
        self.k = k
        self.lib.lxo_renorm_l2(d, nq, self.x.ctypes.data)
        self.D = np.empty((nq, k), dtype=np.float32)
        self.I = np.empty((nq, k), dtype=np.int64)
        self.lib.lxo_heap_init(nq, k, self.D.ctypes.data, self.I.ctypes.data)
        self.seconds += time.perf_counter() - t0

    def add_block(self, rows32: np.ndarray, j0: int, block: int = 16384):
        t0 = time.perf_counter()
        nq = self.x.shape[0]
        for b0 in range(0, rows32.shape[0], block):
            s = np.ascontiguousarray(self.x @ rows32[b0 : b0 + block].T)
            self.lib.lxo_heap_add_block(nq, self.k, self.D.ctypes.data, self.I.ctypes.data, s.ctypes.data, s.shape[1], j0 + b0)
        self.seconds += time.perf_counter() - t0

    def finish(self):
        t0 = time.perf_counter()
        self.lib.lxo_heap_finish(self.x.shape[0], self.k, self.D.ctypes.data, self.I.ctypes.data)
        self.seconds += time.perf_counter() - t0
        return self.D, self.I

    def search(self, corpus32: np.ndarray, x: np.ndarray, k: int):
        self.begin(x, k)
        self.add_block(corpus32, 0)
        return self.finish()


def stream_rows_fp32(corpus, chunk_rows=262144):
    """Yields (first row, fp32 numpy rows) over a CUDA corpus tensor: up-cast on the GPU (exact for
    fp16), copied through one pinned buffer.  The array is only valid until the next iteration."""
    import torch

    n, d = corpus.shape
    chunk_rows = min(chunk_rows, max(1, n))
    pin = torch.empty((chunk_rows, d), dtype=torch.float32, pin_memory=True)
    for r0 in range(0, n, chunk_rows):
        m = min(chunk_rows, n - r0)
        pin[:m].copy_(corpus[r0 : r0 + m], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        yield r0, pin[:m].numpy()


def host_pass(corpus, row_offset, x_check, x_cpu, k):
    """ONE streaming pass over this rank's corpus rows on the host: the exact-arithmetic top-k of the
    check queries (oracle, the checker) and - when x_cpu is given - the timed fp32 FAISS restatement
    on its query sample (the cpu_baseline).  Returns (exact D64, exact I, port or None)."""
    from oracle import c_oracle, faiss_flat as ff

    xn = np.ascontiguousarray(x_check, dtype=np.float32).copy()
    ff.normalize_L2(xn)
    exact = c_oracle.ExactTopK(xn, k)
    port = None
    if x_cpu is not None:
        port = CpuFlatIP()
        port.begin(x_cpu, k)
    for r0, rows in stream_rows_fp32(corpus):
        exact.add(rows, row_offset + r0)
        if port is not None:
            port.add_block(rows, row_offset + r0)
    if port is not None:
        port.finish()
    return exact, port


# --------------------------------------------------------------------------- one workload
def run_workload(key, wl, args, dev, rank, world, peaks, sharded, cpu_baseline, parity):
    """Builds the corpus of one BASELINE.json config on this rank and measures it.  sharded: rows are
    split over the ranks behind ShardedFlatIP (the same code at world == 1, minus the NCCL call);
    otherwise every rank holds the whole corpus and searches its own batches (replicas)."""
    import torch
    import torch.distributed as dist

    from lean_explore_b200 import GpuIndexFlatIP
    from lean_explore_b200.sharded import ShardedFlatIP, shard_rows

    n, d, k, q = wl["n"], wl["d"], wl["k"], wl["q"]
    steps, warmup = args.steps, max(3, args.warmup)
    lo, hi = shard_rows(n, world, rank) if sharded else (0, n)
    corpus = make_corpus_gpu(hi - lo, d, wl["dtype"], dev, row0=lo)
    index = GpuIndexFlatIP.from_tensor(corpus, row_offset=lo)
    engine = ShardedFlatIP(index, world if sharded else 1, rank if sharded else 0, timing=True) if sharded else None
    # sharded: the same queries on every rank; replicas: every rank its own batches
    batches = [make_queries_gpu(q, d, dev, seed=(0 if sharded else rank * 16) + s) for s in range(4)]
    rows_local = hi - lo
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

    def max_over_ranks(*vals):
        if world == 1:
            return vals
        t = torch.tensor(vals, device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return tuple(float(v) for v in t)

    for i in range(warmup):
        step(i)
    barrier()
    if sharded:
        engine.pop_timing()

    # ---- the timed region: K steps, CUDA events on the launching stream, max over ranks
    sampler = ClockSampler(dev.index or 0)
    index.set_timing(True)
    index.get_timing()
    with sampler:
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(steps):
            step(i)
        ev1.record()
        barrier()
    ms = ev0.elapsed_time(ev1)
    tm = index.get_timing()
    index.set_timing(False)
    phases = engine.pop_timing() if sharded else None
    if sharded:
        engine.timing = False
    my_launches = index.last_stats()["kernel_launches"] + (1 if sharded else 0)  # + merge_shards_kernel
    (ms,) = max_over_ranks(ms)
    units = q * steps * (1 if sharded else world)
    res = {"value": round(units / (ms / 1e3), 1), "ms_per_step": round(ms / steps, 4), "gpu_launches": my_launches * steps,
           "clocks": sampler.summary()}

    # ---- end to end through the public host API: numpy in, numpy out, copies inside the timing
    api = engine if sharded else index
    e2e = {}
    for kind in ("pinned", "pageable"):
        if kind == "pinned":
            xh = [b.cpu().pin_memory().numpy() for b in batches]
        else:
            xh = [np.array(b.cpu().numpy(), copy=True) for b in batches]
        api.search(xh[0], k, normalize=True)
        barrier()
        t0 = time.perf_counter()
        for i in range(steps):
            api.search(xh[i % len(xh)], k, normalize=True)
        torch.cuda.synchronize()
        (dt,) = max_over_ranks(time.perf_counter() - t0)
        e2e[kind] = units / dt
    res["e2e"] = {"value": round(e2e["pinned"], 1), "unit": "queries/s", "h2d_bytes_per_step": q * d * 4,
                  "d2h_bytes_per_step": q * k * 12,
                  "api": ("ShardedFlatIP.search(numpy) -> lxg_search_ex + all_gather + lxg_merge_topk_packed" if sharded
                          else "GpuIndexFlatIP.search(numpy) -> lxg_search"),
                  "host_buffers": "pinned (page-locked) query / result arrays",
                  "pageable_value": round(e2e["pageable"], 1),
                  "pageable_note": "what SearchEngine passes today (np.array, engine.py:238): bounced through a pinned stage"}

    # ---- roofline of the dominant kernel (pass-1 scan), per launch, from CUDA events on its stream
    calls = max(1, tm["calls"])
    scan_ms = tm["scan_ms"] / calls
    flops = 2.0 * q * rows_local * d
    esize = 2 if wl["dtype"] == "float16" else 4  # bytes per stored corpus element
    bytes_alg = rows_local * d * esize + q * d * 4 + q * k * 12
    tf = flops / (scan_ms / 1e3) / 1e12
    gbs = bytes_alg / (scan_ms / 1e3) / 1e9
    tensor_frac, hbm_frac = tf / peaks["tflops"], gbs / peaks["hbm_gbs"]
    ridge_q = peaks["tflops"] * 1e12 / (peaks["hbm_gbs"] * 1e9) * esize / 2  # flop/byte == 2 Q / esize
    # Which measured tensor peak applies (task contract: "the burst figure for a kernel timed alone, the
    # sustained one for a kernel timed inside a long step"): the kernel is timed per launch with CUDA events
    # inside K back-to-back steps; when that region lasts a quarter of a second or more the chip is at its
    # power limit for most of it (cfg4: 20 x 20 ms; `clocks` shows the SM clock and sw_power_cap) and the
    # sustained cuBLAS figure is the ceiling; shorter regions (cfg2: 8 ms, cfg3: 45 ms) run at burst clocks.
    long_region = ms >= 250.0
    tpeak = peaks["tflops_sustained"] if long_region else peaks["tflops"]
    if q >= ridge_q:
        roof = dict(bound="tensor", achieved=round(tf, 2), peak=tpeak, unit="TFLOP/s", frac=round(tf / tpeak, 4),
                    frac_of_burst_peak=round(tf / peaks["tflops"], 4), frac_of_sustained_peak=round(tf / peaks["tflops_sustained"], 4))
        psrc = ("sustained (kernel timed per launch inside a %.2f s back-to-back region)" % (ms / 1e3) if long_region
                else "burst (kernel timed per launch, %.0f ms region)" % ms)
    else:
        roof = dict(bound="hbm", achieved=round(gbs, 1), peak=peaks["hbm_gbs"], unit="GB/s", frac=round(hbm_frac, 4))
        psrc = "copy bandwidth (kernel timed per launch)"
    roof.update(traffic=None, traffic_source=None, kernel="scan_topk_kernel", ms_per_launch=round(scan_ms, 4),
                peak_source=peaks["source"] + " " + psrc,
                other_bound_frac=round(hbm_frac if roof["bound"] == "tensor" else tensor_frac, 4),
                rows_per_launch=rows_local, algorithmic_bytes=int(bytes_alg), algorithmic_flops=flops,
                prep_ms_per_launch=round(tm.get("prep_ms", 0.0) / calls, 4),
                merge_ms_per_launch=round(tm["merge_ms"] / calls, 4),
                exact_ms_per_launch=round(tm["exact_ms"] / calls, 4),
                step_frac=round(flops / (ms / steps / 1e3) / 1e12 / peaks["tflops"], 4))
    traffic_file = ROOT / "profiles" / "traffic.json"
    if traffic_file.exists():
        tj = json.loads(traffic_file.read_text())
        ent = tj.get("%dx%d" % (rows_local, d))
        if ent:
            fresh = ent.get("scan_sha") == scan_source_sha()
            roof["traffic"] = ent["bytes"] if fresh else None
            if fresh and "tensor_pipe_active_pct_of_elapsed" in ent:
                # same capture: sm__pipe_tensor_cycles_active (% of elapsed cycles, at the clock the capture ran at)
                roof["ncu_tensor_pipe_active_pct"] = ent["tensor_pipe_active_pct_of_elapsed"]
                roof["ncu_sm_clock_ghz"] = ent.get("sm_clock_ghz_under_ncu")
            roof["traffic_source"] = ("static ncu --set full capture %s (%s), same kernel sources" % (ent["file"], ent["date"])
                                      if fresh else "stale: %s was captured on other kernel sources" % ent["file"])
    res["roofline"] = roof
    if phases:
        pc = max(1, phases["calls"])
        res["phases"] = {"local_search_ms": round(phases["local_ms"] / pc, 4),
                         "collective_ms": round(phases["collective_ms"] / pc, 4),
                         "shard_merge_ms": round(phases["merge_ms"] / pc, 4),
                         "collective": ("all_gather_into_tensor of %d B per rank" % (q * k * 16)) if world > 1 else "none (one shard)"}

    # ---- sustained: the same step for >= 2 s (clocks settle at the power limit), against the sustained peak
    if not args.no_sustained:
        n_sus = max(steps, int(2000.0 / max(ms / steps, 1e-3)) + 1)
        sus_sampler = ClockSampler(dev.index or 0)
        with sus_sampler:
            barrier()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for i in range(n_sus):
                step(i)
            s1.record()
            barrier()
        (sms,) = max_over_ranks(s0.elapsed_time(s1))
        step_tf = 2.0 * q * rows_local * d / (sms / n_sus / 1e3) / 1e12
        res["sustained"] = {"seconds": round(sms / 1e3, 2), "steps": n_sus,
                            "value": round(q * n_sus * (1 if sharded else world) / (sms / 1e3), 1),
                            "ms_per_step": round(sms / n_sus, 4), "step_tflops_per_gpu": round(step_tf, 1),
                            "peak": peaks["tflops_sustained"], "frac_of_sustained_peak": round(step_tf / peaks["tflops_sustained"], 4),
                            "clocks": sus_sampler.summary()}

    # ---- parity at full size + the CPU baseline, in ONE host pass over the corpus
    if parity or cpu_baseline:
        cq = min(wl["check_q"], q)
        xs_all = torch.cat(batches).cpu().numpy()
        x_cpu = xs_all[: min(wl["cpu_q"], xs_all.shape[0])] if cpu_baseline else None
        Dg, Ig = api.search(xs_all if cpu_baseline else xs_all[:q], k, normalize=True)
        exact, port = host_pass(corpus, lo, xs_all[:cq], x_cpu, k)
        if sharded and world > 1:
            parts = [None] * world
            dist.all_gather_object(parts, (exact.D, exact.I))
            if rank == 0:
                for r in range(1, world):
                    exact.merge(*parts[r])
        D64, I64 = exact.result()
        live = I64 >= 0
        res["parity"] = {"queries_checked": cq, "corpus_rows_checked": n,
                         "oracle": "exact arithmetic (fp64 sums of exact products) over the full corpus, oracle/flat_ip.c",
                         "ids_equal_exact": round(float((Ig[:cq] == I64).mean()), 6),
                         "max_abs_score_err": float(np.abs(Dg[:cq][live] - D64[live]).max()) if live.any() else 0.0}
        if port is not None:
            Ic, nq_s = port.I, port.x.shape[0]
            differ = Ig[:nq_s] != Ic
            # a disagreement with the fp32 restatement is legitimate where two ranks are closer than fp32
            # sgemm noise: the GPU's own scores are the correctly rounded exact ones (checked above)
            scale = 1.0  # queries are normalised by the search, corpus rows are unit vectors: |score| <= 1
            gap = np.abs(np.diff(Dg[:nq_s].astype(np.float64), axis=1))
            near = np.zeros(differ.shape, dtype=bool)
            tol = 4e-6 * scale
            near[:, :-1] |= gap < tol
            near[:, 1:] |= gap < tol
            res["cpu_baseline"] = {
                "value": round(nq_s / port.seconds, 1), "unit": "queries/s", "cores": port.threads, "kind": "port",
                "sample": "%d queries x full %d x %d corpus (fp32 rows streamed from HBM), %.1f s of renorm + sgemm + heaps; "
                          "numpy/OpenBLAS sgemm + FAISS-style heaps (faiss-cpu is not installable offline)" % (nq_s, n, d, port.seconds),
                "host_cpus": os.cpu_count(),
                "ids_vs_fp32_port": {"frac_equal": round(float(1.0 - differ.mean()), 6), "mismatches": int(differ.sum()),
                                     "mismatches_at_near_ties": int((differ & near).sum()), "near_tie_tol": tol,
                                     "positions": int(differ.size)}}
    res["_rows_local"] = rows_local
    res["_index"] = index
    res["_corpus"] = corpus
    return res


# --------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--queries", type=int, default=None)
    ap.add_argument("--k", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-sustained", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary cfg3 / cfg2 / sweep / encoder numbers")
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

    assert torch.cuda.is_available(), "bench.py needs a B200; there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    sharded = args.workload == "cfg4"
    n, d, k, q = wl["n"], wl["d"], wl["k"], wl["q"]
    peaks = load_peaks()
    res = run_workload(args.workload, wl, args, dev, rank, world, peaks, sharded,
                       cpu_baseline=(world == 1 and not args.no_cpu_baseline), parity=not args.no_parity)
    index, corpus = res.pop("_index"), res.pop("_corpus")
    res.pop("_rows_local")

    out = {
        "metric": "queries/sec top-50 over Nxd corpus" if k == 50 else f"queries/sec top-{k} over Nxd corpus",
        "value": res["value"], "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "f16xf16->f32 scan, f64 re-score",
        "data": "synthetic",
        "config": {"workload": wl["name"], "corpus_rows": n, "d": d, "corpus_dtype": wl["dtype"], "k": k,
                   "queries_per_step": q, "normalize": True,
                   "parallelism": ("row-sharded x%d, one NCCL all-gather of per-shard top-k per step" % world) if sharded
                   else ("query-parallel replicas x%d" % world if world > 1 else "single GPU"),
                   "l2_policy": "corpus shard (%.0f MB per GPU) larger than L2; 4 rotating query batches" % (n * d * 2 / 1e6 / (world if sharded else 1))
                   if n * d * 2 / (world if sharded else 1) > 126e6 else "inputs fit L2 (config as specified by BASELINE.json)"},
        "e2e": res["e2e"], "gpu_launches": int(res["gpu_launches"]), "roofline": res["roofline"], "clocks": res["clocks"],
    }
    for key in ("phases", "sustained", "parity", "cpu_baseline"):
        if key in res:
            out[key] = res[key]

    if rank == 0 and world == 1 and not args.no_extra and args.workload == "cfg4":
        extra = out["extra"] = {}
        del index, corpus, res  # release the 24.6 GB corpus
        torch.cuda.empty_cache()
        for key in ("cfg3", "cfg2", "cfg1"):
            try:
                w2 = dict(WORKLOADS[key])
                r2 = run_workload(key, w2, args, dev, 0, 1, peaks, sharded=False,
                                  cpu_baseline=not args.no_cpu_baseline, parity=not args.no_parity)
                ix2 = r2.pop("_index")
                r2.pop("_corpus")
                r2.pop("_rows_local")
                cbytes = w2["n"] * w2["d"] * (2 if w2["dtype"] == "float16" else 4)
                r2["config"] = {"workload": w2["name"], "corpus_rows": w2["n"], "d": w2["d"], "k": w2["k"],
                                "queries_per_step": w2["q"], "steps": args.steps, "corpus_dtype": w2["dtype"],
                                "l2_policy": ("corpus (%.0f MB) larger than L2; 4 rotating query batches" % (cbytes / 1e6)) if cbytes > 126e6
                                else ("corpus (%.0f MB) fits the 126 MB L2 - the size BASELINE.json specifies; 4 rotating query batches" % (cbytes / 1e6))}
                extra[key] = r2
                if key == "cfg2":
                    extra["cfg2 query-batch sweep"] = extra_numbers(ix2, w2["d"], w2["k"], dev, peaks)
                del ix2
            except Exception as exc:  # noqa: BLE001 - secondary numbers must not lose the headline line
                extra[key] = {"error": repr(exc)}
            torch.cuda.empty_cache()
        try:
            # the shape of the index the reference ships: Qwen3-Embedding vectors (d = 1024) stored as
            # fp32 (extract/index.py:59-71), ~400 k declarations; one query with faiss_k = 1000 is the
            # engine's own request, the 1024-query batch shows the d = 1024 scan variant's throughput
            from lean_explore_b200 import GpuIndexFlatIP

            big = GpuIndexFlatIP.from_tensor(make_corpus_gpu(400_000, 1024, "float32", dev, seed=3))
            extra["shipped index shape: 400k x 1024 fp32"] = extra_numbers(big, 1024, 50, dev, peaks,
                                                                           sweep=((1, 1000), (64, 1000), (1024, 50)))
            del big
        except Exception as exc:  # noqa: BLE001
            extra["shipped index shape: 400k x 1024 fp32"] = {"error": repr(exc)}
        try:
            extra["encoder"] = encoder_numbers(dev, cpu=not args.no_cpu_baseline)
        except Exception as exc:  # noqa: BLE001
            extra["encoder"] = {"error": repr(exc)}
        try:
            extra["qwen3"] = decoder_numbers(dev, cpu=not args.no_cpu_baseline)
        except Exception as exc:  # noqa: BLE001
            extra["qwen3"] = {"error": repr(exc)}

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
            if b == 1:
                # the query path is a weight read: fraction of the HBM roofline, and the layered kernels beside it
                wbytes = 2.0 * g["num_hidden_layers"] * (4 * g["hidden_size"] ** 2 + 2 * g["hidden_size"] * g["intermediate_size"])
                entry[label]["weight_read_hbm_frac"] = round(wbytes / (ms / 1e3) / 1e9 / load_peaks()["hbm_gbs"], 4)
                enc.set_fused(False)
                for _ in range(3):
                    enc.encode_ids_torch(ids, mask)
                e0.record()
                for _ in range(iters):
                    enc.encode_ids_torch(ids, mask)
                e1.record()
                torch.cuda.synchronize()
                entry[label]["layered_kernels_ms_per_call"] = round(e0.elapsed_time(e1) / iters, 4)
                entry[label]["layered_kernels_launches"] = enc.last_launches()
                enc.set_fused(True)
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
    oracle/) on the host cores, same config/metric; rank 0 only.  Fixed sizes (no calibration): every
    step searches REF_Q queries; cfg4's 16M x 768 fp32 rows (49 GB) are represented by a 2M-row
    slice and the QPS is scaled by 1/8 (brute force is linear in the rows), stated in `sample`."""
    if rank != 0:
        return
    try:
        from threadpoolctl import threadpool_limits

        threadpool_limits(limits=os.cpu_count() or 1)
    except Exception:  # noqa: BLE001
        pass
    n, d, k = wl["n"], wl["d"], wl["k"]
    scale_rows = 1.0
    if n > 2_000_000:
        scale_rows = 2_000_000 / n
        n = 2_000_000
    rng = np.random.default_rng(0)
    c32 = np.empty((n, d), dtype=np.float32)
    for j0 in range(0, n, 65536):
        blk = rng.standard_normal((min(65536, n - j0), d), dtype=np.float32)
        blk /= np.linalg.norm(blk, axis=1, keepdims=True)
        c32[j0 : j0 + blk.shape[0]] = blk.astype(np.float16).astype(np.float32) if wl["dtype"] == "float16" else blk
    nq_step = min(wl["q"], max(16, int(0.8e12 / (2.0 * n * d))))  # ~0.8 TFLOP of sgemm per step
    xs = np.random.default_rng(1).standard_normal((nq_step, d), dtype=np.float32)
    cpu = CpuFlatIP()
    warm = min(args.warmup, 3)
    t_w = time.perf_counter()
    for _ in range(max(1, warm)):
        cpu.search(c32, xs, k)
    per_step = (time.perf_counter() - t_w) / max(1, warm)
    steps = max(1, min(args.steps, int(150.0 / max(per_step, 1e-3))))  # whole run bounded to a few minutes
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu.search(c32, xs, k)
    dt = time.perf_counter() - t0
    value = nq_step * steps / dt * scale_rows
    out = {
        "impl": "reference", "metric": "queries/sec top-50 over Nxd corpus" if k == 50 else f"queries/sec top-{k} over Nxd corpus",
        "value": round(value, 1), "unit": "queries/s", "n_gpus": args.gpus, "steps": steps, "warmup": max(1, warm),
        "ms_per_step": round(dt / steps * 1e3, 3), "higher_is_better": True,
        "scaling": "strong" if args.workload == "cfg4" else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "corpus_rows": wl["n"], "d": d, "corpus_dtype": wl["dtype"], "k": k,
                   "queries_per_step": nq_step, "normalize": True, "parallelism": "host CPU, all cores"},
        "cpu_baseline": {"value": round(value, 1), "unit": "queries/s", "cores": cpu.threads, "kind": "port",
                         "sample": "%d queries/step x %d steps over %d x %d fp32 rows%s; faiss-cpu is not installable "
                                   "offline, so this is its restatement (numpy/OpenBLAS sgemm + FAISS-style heaps)"
                                   % (nq_step, steps, n, d,
                                      " (2M-row slice of the 16M rows, QPS scaled by 1/8)" if scale_rows != 1.0 else ""),
                         "host_cpus": os.cpu_count(), "affinity": len(os.sched_getaffinity(0))},
        "e2e": {"value": round(value, 1), "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
