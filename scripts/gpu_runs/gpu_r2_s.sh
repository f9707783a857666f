#!/bin/bash
# ncu captures of the HuBERT GEMMs: conv1 (first conv_tc<128,8,0> launch of a forward) and a QKV projection (first <128,8,1>)
mkdir -p gpurun_out
# one forward = 13 + 19 conv_tc<128,...> launches + 1 <64>; warm-up 2 forwards, capture in the third
for spec in "conv1 66 0" "qkv 74 1" "fc2 77 2"; do
  set -- $spec
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_tc_kernel --launch-skip $2 --launch-count 1 \
    -o /tmp/h_$1 -f python scripts/bench_hubert.py 32 96000 1 > gpurun_out/ncu_h_$1.log 2>&1
  ncu -i /tmp/h_$1.ncu-rep --page details > gpurun_out/r02_hub_$1_details.txt 2>&1
  ncu -i /tmp/h_$1.ncu-rep --page source --csv > gpurun_out/r02_hub_$1_source.csv 2>&1
  grep -m1 "conv_tc_kernel" gpurun_out/r02_hub_$1_details.txt | cut -c1-200
  grep -E "^    Duration|Grid Size" gpurun_out/r02_hub_$1_details.txt
done
