#!/usr/bin/env python
"""Warp-stall samples of one kernel aggregated per source line.

usage: ncu_lines.py REPORT.ncu-rep CUBIN MANGLED_SUBSTRING [TOP]
Joins `ncu -i REPORT --page source --csv --print-source sass` (samples per SASS instruction, in
program order) with `nvdisasm -g -c CUBIN` (line info per SASS instruction of the same build).
"""
import csv, re, subprocess, sys
from collections import Counter

rep, cubin, sym = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
insts, cur, on = [], None, False
for l in dis:
    if l.startswith(".text."):
        on = sym in l
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", l):
        insts.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, body = rows[1], rows[2:]
isamp, iexe = hdr.index("# Samples"), hdr.index("Instructions Executed")
assert len(body) == len(insts), (len(body), len(insts))
agg, exe = Counter(), Counter()
for r, key in zip(body, insts):
    agg[key] += int(r[isamp])
    exe[key] += int(r[iexe])
tot = sum(agg.values())
print("kernel", rows[0][1][:80], "samples", tot, "warp-instructions", sum(exe.values()))
src = {}
for k, v in agg.most_common(top):
    t = ""
    if k:
        try:
            src.setdefault(k[0], open("lean_explore_b200/csrc/" + k[0]).read().split("\n"))
            t = src[k[0]][k[1] - 1].strip()[:90]
        except OSError:
            pass
    print(f"{str(k):30s} {v * 100 / tot:5.1f}% exec={exe[k]:>10} {t}")
