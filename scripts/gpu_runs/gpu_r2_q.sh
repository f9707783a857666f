#!/bin/bash
# HuBERT after the packed-row / wave-moment / GEMM k-means changes: parity tests, then compute-sanitizer memcheck
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hubert_gpu.py -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -5
export PYTHONUNBUFFERED=1
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 9 \
  python -m pytest tests/test_hubert_gpu.py -m gpu -x -q -k "varlen_batch or large_codebook or attention_tensor_core" \
  > gpurun_out/r02_sanitize_memcheck_hubert.log 2>&1; echo "memcheck rc=$?"
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/r02_sanitize_memcheck_hubert.log; tail -6 gpurun_out/r02_sanitize_memcheck_hubert.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 --error-exitcode 9 \
  python -m pytest tests/test_hubert_gpu.py -m gpu -x -q -k "varlen_batch" \
  > gpurun_out/r02_sanitize_racecheck_hubert.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r02_sanitize_racecheck_hubert.log
timeout 300 python scripts/bench_hubert.py 32 96000 8 > gpurun_out/r02o_hubert_bench.txt 2>&1; cat gpurun_out/r02o_hubert_bench.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02o_hubert_launches.csv python scripts/bench_hubert.py 32 96000 1 > /dev/null 2>&1
