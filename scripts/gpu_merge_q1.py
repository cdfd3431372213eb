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
