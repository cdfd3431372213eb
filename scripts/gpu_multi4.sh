N=${N:-4}
set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r1H3_bench_cfg2_n$N.json 2> gpurun_out/r1H3_bench_cfg2_n$N.err
cut -c1-700 gpurun_out/r1H3_bench_cfg2_n$N.json; tail -2 gpurun_out/r1H3_bench_cfg2_n$N.err
