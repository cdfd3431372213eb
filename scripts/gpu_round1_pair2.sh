set -x
T=${TAG:-r1pair2}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_rerank.csv python scripts/gpu_decoder_prof.py > /dev/null 2>&1
python scripts/ncu_launches.py gpurun_out/${T}_launches_rerank.csv > gpurun_out/${T}_launches_rerank.txt 2>&1; cat gpurun_out/${T}_launches_rerank.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_pair -s 6 -c 1 -o gpurun_out/${T}_prof_gemm_gu python scripts/gpu_decoder_prof.py > gpurun_out/${T}_ncu.log 2>&1
tail -3 gpurun_out/${T}_ncu.log
ncu -i gpurun_out/${T}_prof_gemm_gu.ncu-rep --page raw --csv > gpurun_out/${T}_prof_gemm_gu_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/${T}_prof_gemm_gu.ncu-rep 2>&1 | tail -40
