set -x
T=${TAG:-r1sq}
timeout 400 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 5 -c 1 -o gpurun_out/${T}_prof_scan_q64 python scripts/gpu_scan_q.py 64 > gpurun_out/${T}_ncu.log 2>&1
tail -3 gpurun_out/${T}_ncu.log
python scripts/ncu_summary.py gpurun_out/${T}_prof_scan_q64.ncu-rep 2>&1 | grep -E "time_duration|dram__bytes_read.sum \[|per_second|stalled|issue_active|warps_active|grid" | head -40
cuobjdump -xelf all lean_explore_b200/liblxg.so > /dev/null 2>&1; ls *.cubin | head
for c in *.cubin; do python scripts/ncu_lines.py gpurun_out/${T}_prof_scan_q64.ncu-rep $c "scan_topk_kernelILi128ELb0ELi0" 30 > gpurun_out/${T}_lines_$c.txt 2>/dev/null && head -40 gpurun_out/${T}_lines_$c.txt; done
rm -f *.cubin
