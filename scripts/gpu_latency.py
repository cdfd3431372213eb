"""Single-query latency of the search for the reference's default faiss_k=1000 (engine.py:538)."""
import json, sys, time
import torch
sys.path.insert(0, ".")
from bench import make_corpus_gpu, make_queries_gpu
from lean_explore_b200 import GpuIndexFlatIP
dev = torch.device("cuda", 0)
out = {}
for name, n, d in (("500k x 384", 500_000, 384), ("2M x 768", 2_000_000, 768)):
    corpus = make_corpus_gpu(n, d, "float16", dev)
    index = GpuIndexFlatIP.from_tensor(corpus)
    for q, k in ((1, 50), (1, 200), (1, 1000), (8, 1000), (64, 1000)):
        xs = [make_queries_gpu(q, d, dev, seed=s) for s in range(4)]
        for i in range(3):
            index.search_torch(xs[i], k, normalize=True)
        torch.cuda.synchronize()
        index.set_timing(True); index.get_timing()
        steps = 50
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            index.search_torch(xs[i % 4], k, normalize=True)
        e1.record(); torch.cuda.synchronize()
        tm = index.get_timing(); index.set_timing(False)
        xh = xs[0].cpu().pin_memory().numpy()
        index.search(xh, k, normalize=True)
        t0 = time.perf_counter()
        for i in range(steps):
            index.search(xh, k, normalize=True)
        host = (time.perf_counter() - t0) / steps * 1e3
        st = index.last_stats()
        out[f"{name} Q={q} k={k}"] = dict(ms=round(e0.elapsed_time(e1) / steps, 4), host_call_ms=round(host, 4),
                                          **{kk: round(v / steps, 4) for kk, v in tm.items() if kk != "calls"}, slices=st["slices"], kp=st["kp"])
    del index, corpus
print(json.dumps(out, indent=1))
