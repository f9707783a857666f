#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 9 \
  python -m pytest tests/test_layers_gpu.py tests/test_generator_gpu.py -m gpu -x -q \
  -k "pair64_ragged_time or (pair_resblock_shapes and 64 and 11) or tc_conv_transpose1d_lengths or golden_tiny or ragged" \
  > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/sanitize_memcheck.log; tail -8 gpurun_out/sanitize_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 --error-exitcode 9 \
  python -m pytest tests/test_layers_gpu.py -m gpu -x -q -k "pair64_ragged_time and 118" \
  > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -8 gpurun_out/sanitize_racecheck.log
