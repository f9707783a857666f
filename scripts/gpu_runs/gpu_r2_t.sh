#!/bin/bash
# HuBERT GEMMs: 256-column single-accumulator chunks (default) vs 128-column dual-accumulator chunks
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hubert_gpu.py -m gpu -q -x --no-header -p no:cacheprovider -s 2>&1 | grep -E "passed|failed|err|Error|hubert:|assert" | head -20
for v in 1 0; do DISSC_HUB_NC256=$v timeout 300 python scripts/bench_hubert.py 32 96000 8 2>&1 | grep "hubert encode" | sed "s/^/NC256=$v /"; done | tee gpurun_out/r02_hubert_nc256_ab.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02o_hubert_launches.csv python scripts/bench_hubert.py 32 96000 1 > /dev/null 2>&1
