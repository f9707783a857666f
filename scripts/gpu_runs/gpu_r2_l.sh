#!/bin/bash
mkdir -p gpurun_out
R=$PWD/dissc_b200
timeout 900 python scripts/ab_layers.py --rounds 3 ns2000: ns200:DISSC_LIB=$R/libdissc_b200_ns200.so ns20000:DISSC_LIB=$R/libdissc_b200_ns20000.so > gpurun_out/r2l_ab.txt 2>&1
grep -E "^s0 |^s1 |^s2 |^s3 |^s4 |TOTAL|^ups " gpurun_out/r2l_ab.txt
for v in "" $R/libdissc_b200_ns200.so $R/libdissc_b200_ns20000.so; do DISSC_LIB=$v python scripts/bench_hubert.py 32 96000 8 2>&1 | head -1; done
