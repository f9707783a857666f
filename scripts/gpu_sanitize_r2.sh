#!/bin/bash
# compute-sanitizer over the kernels added in round 2: packed C = 16 pair kernel, tcgen05 attention, 64-channel blocks,
# two-chunk 256-column GEMMs, cluster multicast (opt-in), embedding range checks.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 9 \
  python -m pytest tests/test_layers_gpu.py tests/test_generator_gpu.py tests/test_hubert_gpu.py -m gpu -x -q \
  -k "(pair_ragged_time and (117 or 1025)) or (pair_resblock_shapes and 16 and 11) or (tc_conv_resblock_shapes and 256 and 11) or golden_tiny or golden_ragged or out_of_range or varlen_batch" \
  > gpurun_out/r02_sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/r02_sanitize_memcheck.log; tail -6 gpurun_out/r02_sanitize_memcheck.log
DISSC_TC_CLUSTER2=1 timeout 900 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 9 \
  python -m pytest tests/test_layers_gpu.py -m gpu -x -q -k "tc_conv_resblock_shapes and 128 and 11" \
  > gpurun_out/r02_sanitize_memcheck_cluster2.log 2>&1; echo "memcheck cluster2 rc=$?"; tail -4 gpurun_out/r02_sanitize_memcheck_cluster2.log
timeout 1200 compute-sanitizer --tool racecheck --print-limit 20 --error-exitcode 9 \
  python -m pytest tests/test_layers_gpu.py tests/test_hubert_gpu.py -m gpu -x -q -k "(pair_ragged_time and 118) or varlen_batch" \
  > gpurun_out/r02_sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -6 gpurun_out/r02_sanitize_racecheck.log
