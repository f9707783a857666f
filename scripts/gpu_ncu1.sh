#!/bin/bash
mkdir -p gpurun_out
for spec in "0 k3d1" "6 k11d1"; do
  set -- $spec
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:resblock_pair64 --launch-skip $1 --launch-count 1 \
    -o /tmp/p64_$2 -f python scripts/one_forward.py 64 300 1 > gpurun_out/ncu_p64_$2.log 2>&1
  ncu -i /tmp/p64_$2.ncu-rep --page details > gpurun_out/p64_$2_details.txt 2>&1
  ncu -i /tmp/p64_$2.ncu-rep --page source --csv > gpurun_out/p64_$2_source.csv 2>&1
  tail -2 gpurun_out/ncu_p64_$2.log
done
du -sh gpurun_out
