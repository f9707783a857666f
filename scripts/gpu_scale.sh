#!/bin/bash
# usage: gpu_scale.sh N
N=$1
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/scale_bench_n$N.json 2> gpurun_out/scale_bench_n$N.err
cut -c1-200 gpurun_out/scale_bench_n$N.json; tail -3 gpurun_out/scale_bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  scripts/bench_pipeline.py > gpurun_out/scale_pipeline_n$N.txt 2> gpurun_out/scale_pipeline_n$N.err
grep config gpurun_out/scale_pipeline_n$N.txt | cut -c1-330; tail -3 gpurun_out/scale_pipeline_n$N.err
