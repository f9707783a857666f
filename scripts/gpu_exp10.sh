#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/ab_layers.py --rounds 3 na3:DISSC_TC_NA=3 na2:DISSC_TC_NA=2 > gpurun_out/exp10_ab.txt 2>&1
grep "^s0 \|^s1 \|^s2 \|^s3 \|^ups \|^conv_pre\|TOTAL\|layer  " gpurun_out/exp10_ab.txt
grep "s1.rb2.c1.0\|s1.rb1.c1.0\|s1.rb0.c1.0\|s0.rb2.c1.0\|ups.0\|ups.1" gpurun_out/exp10_ab.txt
timeout 600 python -m pytest tests/test_layers_gpu.py tests/test_generator_gpu.py -m gpu -x -q 2>&1 | tail -2
