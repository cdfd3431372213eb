T=${TAG:-enc}
for cfg in "minilm 256 64" "minilm 1 16" "bge 256 64"; do
  set -- $cfg
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_$1_$2x$3.csv python scripts/gpu_encoder_prof.py $1 $2 $3 > /dev/null 2>&1
  python scripts/ncu_launches.py gpurun_out/${T}_launches_$1_$2x$3.csv | grep -v "at::" 
done
