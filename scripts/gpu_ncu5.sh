#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:hub_attention --launch-skip 12 --launch-count 1 \
    -o /tmp/attn -f python scripts/bench_hubert.py 32 96000 1 > gpurun_out/ncu5.log 2>&1
ncu -i /tmp/attn.ncu-rep --page details > gpurun_out/r01_d_hub_attention_details.txt 2>&1
ncu -i /tmp/attn.ncu-rep --page source --csv > gpurun_out/attn_source.csv 2>&1
grep -n "Duration\|Executed Ipc Active\|Issue Slots Busy\|highest-utilized\|Achieved Occupancy\|Registers Per\|Theoretical Occ\|L1/TEX Hit\|Bank conflicts\|bank conflict" gpurun_out/r01_d_hub_attention_details.txt | cut -c1-160
