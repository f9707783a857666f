#!/bin/bash
mkdir -p gpurun_out
DISSC_TC_KB64=1 timeout 900 python -m pytest tests/test_layers_gpu.py tests/test_generator_gpu.py tests/test_hubert_gpu.py -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -4
timeout 900 python scripts/ab_layers.py --rounds 3 base: kb64:DISSC_TC_KB64=1 single:DISSC_TC_SPLIT256=0,DISSC_TC_SINGLE_ACC=1 > gpurun_out/r2j_ab.txt 2>&1
grep -E "^s0 |^s1 |^s2 |^s3 |^s4 |TOTAL|^ups |^conv_pre |^s0.rb0.c1.0|^s0.rb2.c1.0|^s1.rb0.c1.0|^s1.rb2.c1.0" gpurun_out/r2j_ab.txt
DISSC_TC_KB64=1 python scripts/bench_hubert.py 32 96000 5 2>&1 | head -1
python scripts/bench_hubert.py 32 96000 5 2>&1 | head -1
