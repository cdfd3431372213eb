#!/usr/bin/env python
"""Top SASS instructions by warp-stall samples from `ncu -i rep --page source --csv` output."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, isamp, iexe = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
body = rows[2:]
tot = sum(int(r[isamp]) for r in body)
print("total samples", tot)
top = sorted(range(len(body)), key=lambda i: -int(body[i][isamp]))[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for i in sorted(top):
    r = body[i]
    st = sorted(((int(r[c]), hdr[c]) for c in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {int(r[isamp])*100/tot:5.1f}% exec={r[iexe]:>10} {r[isrc].strip()[:80]:80s} {st}")
