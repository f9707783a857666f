#!/bin/bash
mkdir -p gpurun_out
for spec in "0 s3k3" "6 s3k11" "9 s4k3" "15 s4k11"; do
  set -- $spec
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:resblock_pair_tc --launch-skip $1 --launch-count 1 \
    -o /tmp/p_$2 -f python scripts/one_forward.py 64 300 1 > gpurun_out/ncu_p_$2.log 2>&1
  ncu -i /tmp/p_$2.ncu-rep --page details > gpurun_out/r01d_$2_details.txt 2>&1
  ncu -i /tmp/p_$2.ncu-rep --page source --csv > gpurun_out/r01d_$2_source.csv 2>&1
  tail -1 gpurun_out/ncu_p_$2.log
done
du -sh gpurun_out
