"""A few single-query k = 1000 searches over a 400k x 1024 fp32 corpus (the shipped index shape), for ncu."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bench import make_corpus_gpu, make_queries_gpu  # noqa: E402
from lean_explore_b200 import GpuIndexFlatIP  # noqa: E402

dev = torch.device("cuda", 0)
index = GpuIndexFlatIP.from_tensor(make_corpus_gpu(400_000, 1024, "float32", dev, seed=3))
xs = [make_queries_gpu(1, 1024, dev, seed=s) for s in range(4)]
for i in range(8):
    index.search_torch(xs[i % 4], 1000, normalize=True)
torch.cuda.synchronize()
print(index.last_stats())

# timing: the engine's request shape on both corpora (python scripts/gpu_merge_q1.py time)
if len(sys.argv) > 1 and sys.argv[1] == "time":
    import json

    out = {}
    for name, ix, d in (("400k x 1024 fp32", index, 1024),
                        ("500k x 384 fp16", GpuIndexFlatIP.from_tensor(make_corpus_gpu(500_000, 384, "float16", dev)), 384)):
        for q, k in ((1, 1000), (1, 50), (8, 1000), (64, 1000)):
            xq = [make_queries_gpu(q, d, dev, seed=10 + s) for s in range(4)]
            for i in range(3):
                ix.search_torch(xq[i], k, normalize=True)
            torch.cuda.synchronize()
            ix.set_timing(True)
            ix.get_timing()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(50):
                ix.search_torch(xq[i % 4], k, normalize=True)
            e1.record()
            torch.cuda.synchronize()
            tm = ix.get_timing()
            ix.set_timing(False)
            out[f"{name} Q={q} k={k}"] = {"ms_per_search": round(e0.elapsed_time(e1) / 50, 4), "scan_ms": round(tm["scan_ms"] / 50, 4),
                                          "merge_ms": round(tm["merge_ms"] / 50, 4), "launches": ix.last_stats()["kernel_launches"]}
    print(json.dumps(out, indent=1))
