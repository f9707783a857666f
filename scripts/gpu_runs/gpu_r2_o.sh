#!/bin/bash
# HuBERT packed transformer layout: parity tests, then timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hubert_gpu.py -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/r02o_tests.log
cat gpurun_out/r02o_tests.log
timeout 300 python scripts/bench_hubert.py 32 96000 8 > gpurun_out/r02o_hubert_bench.txt 2>&1; cat gpurun_out/r02o_hubert_bench.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02o_hubert_launches.csv python scripts/bench_hubert.py 32 96000 1 > /dev/null 2>&1
