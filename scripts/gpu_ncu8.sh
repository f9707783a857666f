#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_tc_kernel --launch-skip 33 --launch-count 1 \
    -o /tmp/s1 -f python scripts/one_forward.py 64 300 1 > gpurun_out/ncu8.log 2>&1
ncu -i /tmp/s1.ncu-rep --page source --csv > gpurun_out/s1k11_source.csv 2>&1
ncu -i /tmp/s1.ncu-rep --page details | grep -n "conv_tc_kernel\|Duration\|highest-utilized\|Dynamic Shared" | cut -c1-150
