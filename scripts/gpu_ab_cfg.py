"""One build, one workload, device-resident: mean scan / step ms over N iterations (for alternating A/B runs).
python scripts/gpu_ab_cfg.py ROWS D Q K ITERS"""
import json
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bench import make_corpus_gpu, make_queries_gpu  # noqa: E402
from lean_explore_b200 import GpuIndexFlatIP  # noqa: E402

n, d, q, k, iters = (int(a) for a in sys.argv[1:6])
dev = torch.device("cuda", 0)
ix = GpuIndexFlatIP.from_tensor(make_corpus_gpu(n, d, "float16", dev))
xq = [make_queries_gpu(q, d, dev, seed=100 + s) for s in range(4)]
for i in range(10):
    ix.search_torch(xq[i % 4], k, normalize=True)
torch.cuda.synchronize()
ix.set_timing(True)
ix.get_timing()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(iters):
    ix.search_torch(xq[i % 4], k, normalize=True)
e1.record()
torch.cuda.synchronize()
tm = ix.get_timing()
print(json.dumps({"lib": os.path.basename(os.environ.get("LXG_LIB_PATH", "liblxg.so")), "rows": n, "d": d, "q": q,
                  "step_ms": round(e0.elapsed_time(e1) / iters, 4), "scan_ms": round(tm["scan_ms"] / iters, 4),
                  "merge_ms": round(tm["merge_ms"] / iters, 4)}))
