N=${N:-2}
T=${TAG:-r1m}
set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/${T}_bench_cfg2_n$N.json 2> gpurun_out/${T}_bench_cfg2_n$N.err
cat gpurun_out/${T}_bench_cfg2_n$N.json; tail -3 gpurun_out/${T}_bench_cfg2_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload cfg4 --steps 10 --warmup 3 > gpurun_out/${T}_bench_cfg4_n$N.json 2> gpurun_out/${T}_bench_cfg4_n$N.err
cat gpurun_out/${T}_bench_cfg4_n$N.json; tail -3 gpurun_out/${T}_bench_cfg4_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref_n$N.json 2> gpurun_out/${T}_bench_ref_n$N.err
cat gpurun_out/${T}_bench_ref_n$N.json; tail -3 gpurun_out/${T}_bench_ref_n$N.err
