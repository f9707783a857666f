#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_layers_gpu.py tests/test_generator_gpu.py -m gpu -q -x -rf --no-header -p no:cacheprovider -k "pair or generator or golden or oracle or config2" 2>&1 | tail -8 > gpurun_out/r2g_tests.log
echo "tests rc=$?"; tail -4 gpurun_out/r2g_tests.log
timeout 600 python scripts/ab_layers.py --rounds 3 g3: g2:DISSC_TC_PACK2_GROUPS=2 > gpurun_out/r2g_ab.txt 2>&1
grep -E "^s4|^s3 |^s2 |^s1 |^s0 |TOTAL" gpurun_out/r2g_ab.txt
