"""Pass-1 candidate counts (LXG_DEBUG_COUNTS=1 -> stderr) for a few batch sizes; cfg2 and cfg3 corpora.
LXG_DEBUG_COUNTS=1 python scripts/gpu_level_counts.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bench import make_corpus_gpu, make_queries_gpu  # noqa: E402
from lean_explore_b200 import GpuIndexFlatIP  # noqa: E402

dev = torch.device("cuda", 0)
for n, d, dt in ((500_000, 384, "float16"), (2_000_000, 768, "float16"), (400_000, 1024, "float32")):
    index = GpuIndexFlatIP.from_tensor(make_corpus_gpu(n, d, dt, dev))
    for q, k in ((1, 50), (64, 50), (256, 50), (1024, 50), (4096, 50), (1, 1000), (64, 1000)):
        x = make_queries_gpu(q, d, dev, seed=5)
        print(f"--- {n} x {d} {dt}  Q={q} k={k}", file=sys.stderr, flush=True)
        for i in range(2):
            index.search_torch(x, k, normalize=True)
        torch.cuda.synchronize()
