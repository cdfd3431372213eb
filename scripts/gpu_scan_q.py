"""A few searches of Q queries (k = 50) over the cfg2 corpus, for ncu captures of the small-batch regime."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bench import make_corpus_gpu, make_queries_gpu  # noqa: E402
from lean_explore_b200 import GpuIndexFlatIP  # noqa: E402

q = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda", 0)
index = GpuIndexFlatIP.from_tensor(make_corpus_gpu(500_000, 384, "float16", dev))
xs = [make_queries_gpu(q, 384, dev, seed=s) for s in range(4)]
for i in range(8):
    index.search_torch(xs[i % 4], 50, normalize=True)
torch.cuda.synchronize()
print(index.last_stats())
