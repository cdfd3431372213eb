"""Query-batch sweep (device-resident, CUDA events) for A/B runs of two builds of liblxg.so on one box:
LXG_LIB_PATH=... python scripts/gpu_sweep_ab.py   (prints one JSON object)"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bench import make_corpus_gpu, make_queries_gpu  # noqa: E402
from lean_explore_b200 import GpuIndexFlatIP  # noqa: E402

dev = torch.device("cuda", 0)
out = {}
for name, n, d, dt in (("500k x 384 fp16", 500_000, 384, "float16"), ("400k x 1024 fp32", 400_000, 1024, "float32"),
                       ("2M x 768 fp16", 2_000_000, 768, "float16")):
    ix = GpuIndexFlatIP.from_tensor(make_corpus_gpu(n, d, dt, dev, seed=3 if d == 1024 else 0))
    for q, k in ((1, 50), (8, 50), (64, 50), (256, 50), (512, 50), (1024, 50), (4096, 50), (1, 1000), (8, 1000), (64, 1000)):
        xq = [make_queries_gpu(q, d, dev, seed=100 + s) for s in range(4)]
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
        out[f"{name} Q={q} k={k}"] = [round(e0.elapsed_time(e1) / 50, 4), round(tm["scan_ms"] / 50, 4), round(tm["merge_ms"] / 50, 4)]
print(json.dumps(out))
