#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_layers_gpu.py tests/test_generator_gpu.py -m gpu -q -rf --no-header -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/r2d_tests.log
tail -6 gpurun_out/r2d_tests.log
python scripts/ab_layers.py --rounds 3 pk2: nopk2:DISSC_TC_PACK2=0 > gpurun_out/r2d_ab.txt 2>&1
grep -E "^s4|^s3|TOTAL|^s0 |^s1 |^s2 " gpurun_out/r2d_ab.txt
