#!/bin/bash
# compute-sanitizer over the final HuBERT path (zero-stage skipping, 256-column chunks) and the opt-in CTA-pair form
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 9 \
  python -m pytest tests/test_hubert_gpu.py -m gpu -x -q -k "varlen_batch or large_codebook or attention_tensor_core" \
  > gpurun_out/r02_sanitize_memcheck_hubert.log 2>&1; echo "memcheck rc=$?"
tail -4 gpurun_out/r02_sanitize_memcheck_hubert.log
DISSC_HUB_PAIR2=1 timeout 900 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 9 \
  python -m pytest tests/test_hubert_gpu.py -m gpu -x -q -k "varlen_batch" \
  > gpurun_out/r02_sanitize_memcheck_hubert_pair2.log 2>&1; echo "memcheck pair2 rc=$?"
tail -4 gpurun_out/r02_sanitize_memcheck_hubert_pair2.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 --error-exitcode 9 \
  python -m pytest tests/test_hubert_gpu.py -m gpu -x -q -k "varlen_batch" \
  > gpurun_out/r02_sanitize_racecheck_hubert.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r02_sanitize_racecheck_hubert.log
