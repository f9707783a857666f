#!/bin/bash
N=$1
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/scale_bench_n$N.json 2> gpurun_out/scale_bench_n$N.err
cut -c1-220 gpurun_out/scale_bench_n$N.json
