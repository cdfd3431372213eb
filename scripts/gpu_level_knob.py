"""Scan time at a few batch sizes (cfg2 corpus) for one setting of LXG_LVL_SLEEP, the pause between level-warp rounds."""
import json
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bench import make_corpus_gpu, make_queries_gpu  # noqa: E402
from lean_explore_b200 import GpuIndexFlatIP  # noqa: E402

dev = torch.device("cuda", 0)
out = {}
ix = GpuIndexFlatIP.from_tensor(make_corpus_gpu(500_000, 384, "float16", dev))
for q in (1, 64, 1024, 4096):
    xq = [make_queries_gpu(q, 384, dev, seed=100 + s) for s in range(4)]
    for i in range(3):
        ix.search_torch(xq[i], 50, normalize=True)
    torch.cuda.synchronize()
    ix.set_timing(True)
    ix.get_timing()
    for i in range(20):
        ix.search_torch(xq[i % 4], 50, normalize=True)
    torch.cuda.synchronize()
    tm = ix.get_timing()
    ix.set_timing(False)
    out[q] = round(tm["scan_ms"] / 20, 4)
print(os.environ.get("LXG_LVL_SLEEP"), json.dumps(out))
