#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hubert_gpu.py -m gpu -q -x -s -rf --no-header -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/r2h_tests.log
echo "tests rc=$?"; grep -E "attention|hubert|passed|failed|Error|error" gpurun_out/r2h_tests.log | tail -12
timeout 300 python scripts/bench_hubert.py 32 96000 5 2>&1 | tail -3
DISSC_HUB_ATTN_TC=0 timeout 300 python scripts/bench_hubert.py 32 96000 5 2>&1 | tail -3
