#!/bin/bash
mkdir -p gpurun_out
for spec in "6 k11d1" "0 k3d1"; do
  set -- $spec
  timeout 600 ncu --set full --clock-control none -k regex:resblock_pair16 --launch-skip $1 --launch-count 1 \
    -o /tmp/p16_$2 -f python scripts/one_forward.py 64 300 1 > gpurun_out/ncu7_$2.log 2>&1
  ncu -i /tmp/p16_$2.ncu-rep --page details > gpurun_out/p16_$2_details.txt 2>&1
  echo == $2; grep -n "Duration\|SM Frequency\|DRAM Throughput\|Executed Ipc Active\|Issue Slots Busy\|highest-utilized\|Mem Pipes Busy\|Executed Instructions  \|bank conflict" gpurun_out/p16_$2_details.txt | cut -c1-150
done
