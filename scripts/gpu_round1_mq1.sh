set -x
T=${TAG:-r1mq}
timeout 400 ncu --set full --clock-control none --import-source on -k regex:merge_rescore -s 5 -c 1 -o gpurun_out/${T}_prof_merge_q1 python scripts/gpu_merge_q1.py > gpurun_out/${T}_ncu.log 2>&1
tail -3 gpurun_out/${T}_ncu.log
python scripts/ncu_summary.py gpurun_out/${T}_prof_merge_q1.ncu-rep 2>&1 | grep -E "time_duration|grid|block_size" | head
cuobjdump -xelf all lean_explore_b200/liblxg.so > /dev/null 2>&1
python scripts/ncu_lines.py gpurun_out/${T}_prof_merge_q1.ncu-rep lxg_search.sm_100a.cubin "merge_rescore_kernelILi1024" 40 > gpurun_out/${T}_lines.txt 2>&1; cat gpurun_out/${T}_lines.txt
rm -f *.cubin
