#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none -k regex:hub_layernorm --launch-skip 30 --launch-count 1 \
    -o /tmp/ln -f python scripts/bench_hubert.py 32 96000 1 > gpurun_out/ncu6.log 2>&1
ncu -i /tmp/ln.ncu-rep --page details > gpurun_out/hub_ln_details.txt 2>&1
grep -n "Duration\|DRAM Throughput\|Executed Ipc Active\|Issue Slots Busy\|Achieved Occupancy\|Registers Per\|Theoretical Occ\|L1/TEX Hit\|L2 Hit\|Memory Throughput\|Sectors/Req\|uncoalesced\|Uncoalesced\|excessive" gpurun_out/hub_ln_details.txt | cut -c1-170
