N=${N:-2}
T=${TAG:-r1H2}
set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/${T}_bench_cfg2_n$N.json 2> gpurun_out/${T}_bench_cfg2_n$N.err
cut -c1-1500 gpurun_out/${T}_bench_cfg2_n$N.json; tail -3 gpurun_out/${T}_bench_cfg2_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload cfg4 --steps 10 --warmup 3 > gpurun_out/${T}_bench_cfg4_n$N.json 2> gpurun_out/${T}_bench_cfg4_n$N.err
cut -c1-1500 gpurun_out/${T}_bench_cfg4_n$N.json; tail -3 gpurun_out/${T}_bench_cfg4_n$N.err
