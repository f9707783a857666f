#!/bin/bash
mkdir -p gpurun_out
for B in 32 48 64 96 128; do timeout 300 python scripts/bench_hubert.py $B 96000 8 2>&1 | grep "hubert encode"; done | tee gpurun_out/r02_hubert_batch_sweep.txt
