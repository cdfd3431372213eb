#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr, body = rows[0], rows[1:]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in body:
    v = float(r[iv].replace(",", "")) / {"ns": 1e6, "us": 1e3, "ms": 1.0}.get(r[iu], 1.0)
    a = agg.setdefault(r[ik][:90], [0, 0.0])
    a[0] += 1
    a[1] += v
ours = sum(t for k, (n, t) in agg.items() if "lxg::" in k and "corpus_stats" not in k and "make_scan_copy" not in k)
print(f"{len(body)} launches; lxg search kernels total {ours:.3f} ms (cold-cache, serialised under ncu)")
for k, (n, t) in agg.items():
    share = f"{100 * t / ours:5.1f}% of the search step" if "lxg::" in k and "corpus_stats" not in k else ""
    print(f"{n:5d} x {t / n * 1e3:9.1f} us = {t:9.3f} ms  {share:28s} {k}")
