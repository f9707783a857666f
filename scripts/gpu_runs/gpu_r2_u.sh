#!/bin/bash
# HuBERT GEMM tuning sweep (plan-time knobs)
mkdir -p gpurun_out
run() { env "$@" timeout 300 python scripts/bench_hubert.py 32 96000 8 2>&1 | grep "hubert encode\|rror" | sed "s/^/$* : /" | cut -c1-120; }
{
run DISSC_HUB_NC256=1
run DISSC_HUB_NC256=1 DISSC_TC_KB64_256=1
run DISSC_HUB_NC256=1 DISSC_TC_KB64_256=1 DISSC_HUB_CLUSTER2=0
run DISSC_HUB_NC256=0
} | tee gpurun_out/r02_hubert_tuning_sweep2.txt
DISSC_TC_KB64_256=1 timeout 900 python -m pytest tests/test_hubert_gpu.py -m gpu -q -x --no-header -p no:cacheprovider -s 2>&1 | grep -E "passed|failed|err|Error|hubert:|assert" | head -20
