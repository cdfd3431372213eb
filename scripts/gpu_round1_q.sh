set -x
T=${TAG:-r1q}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file gpurun_out/${T}_launches_query.csv python scripts/gpu_decoder_prof.py 1 24 > /dev/null 2>&1
python scripts/ncu_launches.py gpurun_out/${T}_launches_query.csv > gpurun_out/${T}_launches_query.txt 2>&1; cat gpurun_out/${T}_launches_query.txt
