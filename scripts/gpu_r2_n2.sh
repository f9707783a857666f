#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_n2_bench.json 2> gpurun_out/r2_n2_bench.err
echo rc=$?
cut -c1-400 gpurun_out/r2_n2_bench.json; tail -5 gpurun_out/r2_n2_bench.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2_n2_ref.json 2> gpurun_out/r2_n2_ref.err
echo rc=$?; cut -c1-300 gpurun_out/r2_n2_ref.json
