#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:hub_attention_tc --launch-skip 8 --launch-count 1 -o /tmp/attn -f python scripts/bench_hubert.py 32 96000 1 > gpurun_out/r2i_ncu.log 2>&1
ncu -i /tmp/attn.ncu-rep --page details > gpurun_out/r2i_attn_details.txt 2>&1
ncu -i /tmp/attn.ncu-rep --page source --csv > gpurun_out/r2i_attn_source.csv 2>&1
grep -E "Duration|Registers Per|Theoretical Occ|Achieved Occ|Issue Slots Busy|Executed Ipc Active|Warp Cycles Per Issued|Shared Memory Configuration|Dynamic Shared|Block Limit" gpurun_out/r2i_attn_details.txt | head -20
du -sh gpurun_out/r2i_attn_source.csv
