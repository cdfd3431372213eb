#!/usr/bin/env python
"""Print the metrics DESIGN.md / profiles/ quote from an ncu report (ncu -i ... --page raw --csv)."""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct",
        "sm__inst_executed_pipe_tensor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum ", "sm__cycles_elapsed.avg ", "sm__cycles_active.avg ", "smsp__inst_executed.sum ",
        "l1tex__throughput.avg.pct", "lts__throughput.avg.pct", "sm__pipe_tensor_subpipe", "smsp__issue_active.avg.pct",
        "smsp__average_warp", "smsp__warp_issue_stalled", "sm__cycles_elapsed.avg.per_second"]
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u = rows[0], rows[1]
for v in rows[2:]:
    print("== kernel:", v[h.index("Kernel Name")][:70], "grid", v[h.index("Grid Size")], "block", v[h.index("Block Size")])
    for i, name in enumerate(h):
        if any(name.startswith(k.strip()) if k.endswith(" ") else k in name for k in KEYS):
            if v[i] not in ("", "0", "n/a"):
                print(f"  {name} [{u[i]}] = {v[i]}")
