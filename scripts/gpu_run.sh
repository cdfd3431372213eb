#!/bin/bash
# One parametrised GPU-box runner (replaces the round-1 one-off scripts).
#   scripts/gpu_run.sh TAG STEP [STEP ...]       (run through gpurun from the repo root)
# Steps: tests | tests:<pytest args> | bench[:<bench args>] | ref[:<args>] | benchN:<n>[:<args>] |
#        launches[:<bench args>] | ncu:<kernel regex>:<bench args> | san:<tool>:<pytest args> | py:<script> [args]
# Everything lands in gpurun_out/<TAG>_*.
set -u
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/${TAG}_smi.txt 2>&1
i=0
for step in "$@"; do
  i=$((i+1))
  kind=${step%%:*}
  rest=""; [[ "$step" == *:* ]] && rest=${step#*:}
  case $kind in
    tests)
      args=${rest:-"tests -m gpu -x -q"}
      timeout 1500 bash -c "python -m pytest $args" > $OUT/${TAG}_pytest_$i.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest_$i.log; tail -3 $OUT/${TAG}_pytest_$i.log ;;
    bench)
      timeout 900 python bench.py $rest > $OUT/${TAG}_bench_$i.json 2> $OUT/${TAG}_bench_$i.err; echo "bench rc=$?"; head -c 600 $OUT/${TAG}_bench_$i.json; echo ;;
    ref)
      timeout 900 python bench.py --impl reference $rest > $OUT/${TAG}_ref_$i.json 2> $OUT/${TAG}_ref_$i.err; echo "ref rc=$?"; head -c 400 $OUT/${TAG}_ref_$i.json; echo ;;
    benchN)
      n=${rest%%:*}; a=""; [[ "$rest" == *:* ]] && a=${rest#*:}
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n $a > $OUT/${TAG}_bench_n${n}_$i.json 2> $OUT/${TAG}_bench_n${n}_$i.err; echo "benchN rc=$?"; tail -c 1500 $OUT/${TAG}_bench_n${n}_$i.json; echo ;;
    launches)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:scan_topk|merge_|prep_queries|exact_|normalize_l2|make_scan|bert_fused|gemm_|attention_|rmsnorm|layernorm|embed_|pool_|qk_norm|last_token" -c 400 --csv --log-file $OUT/${TAG}_launches_$i.csv python bench.py --steps 2 --warmup 1 --no-extra --no-cpu-baseline --no-parity --no-sustained $rest > $OUT/${TAG}_launches_$i.log 2>&1; echo "launches rc=$?"
      python scripts/ncu_launches.py $OUT/${TAG}_launches_$i.csv > $OUT/${TAG}_launches_$i.txt 2>&1; tail -12 $OUT/${TAG}_launches_$i.txt ;;
    ncu)
      kern=${rest%%:*}; a=${rest#*:}
      timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:$kern" -s 3 -c 1 -f -o $OUT/${TAG}_prof_$i python bench.py --steps 2 --warmup 1 --no-extra --no-cpu-baseline --no-parity --no-sustained $a > $OUT/${TAG}_ncu_$i.log 2>&1; echo "ncu rc=$?"
      python scripts/ncu_summary.py $OUT/${TAG}_prof_$i.ncu-rep > $OUT/${TAG}_ncu_$i.txt 2>&1; tail -30 $OUT/${TAG}_ncu_$i.txt ;;
    san)
      tool=${rest%%:*}; a=${rest#*:}
      timeout ${SAN_TIMEOUT:-600} compute-sanitizer --tool $tool --target-processes all --error-exitcode 9 --log-file $OUT/${TAG}_san_${tool}_$i.log bash -c "python -m pytest $a -x -q" > $OUT/${TAG}_san_${tool}_$i.pytest.log 2>&1; echo "san $tool rc=$?"; tail -5 $OUT/${TAG}_san_${tool}_$i.log; tail -3 $OUT/${TAG}_san_${tool}_$i.pytest.log ;;
    py)
      timeout 1200 python $rest > $OUT/${TAG}_py_$i.log 2>&1; echo "py rc=$?"; tail -40 $OUT/${TAG}_py_$i.log ;;
    *) echo "unknown step $step" ;;
  esac
done
