#!/bin/bash
# ncu source-level captures of the HuBERT GEMMs in the final configuration (256-column single-accumulator chunks, 64-channel blocks)
mkdir -p gpurun_out
for spec in "conv1 66" "qkv 74"; do
  set -- $spec
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_tc_kernel --launch-skip $2 --launch-count 1 \
    -o /tmp/h_$1 -f python scripts/bench_hubert.py 32 96000 1 > gpurun_out/ncu_h_$1.log 2>&1
  ncu -i /tmp/h_$1.ncu-rep --page details > gpurun_out/r02x_hub_$1_details.txt 2>&1
  ncu -i /tmp/h_$1.ncu-rep --page source --csv > gpurun_out/r02x_hub_$1_source.csv 2>&1
  grep -m1 "conv_tc_kernel" gpurun_out/r02x_hub_$1_details.txt | cut -c1-160
  grep -E "^    Duration|TC is" gpurun_out/r02x_hub_$1_details.txt
done
