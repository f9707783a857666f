#!/bin/bash
# source-level capture of the HuBERT attention kernel
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:hub_attention_tc --launch-skip 8 --launch-count 1 -o /tmp/c_attn -f python scripts/bench_hubert.py 32 96000 1 > gpurun_out/ncu_c_attn.log 2>&1
ncu -i /tmp/c_attn.ncu-rep --page details > gpurun_out/r02p_hub_attention_tc_details.txt 2>&1
ncu -i /tmp/c_attn.ncu-rep --page source --csv > gpurun_out/r02p_hub_attention_tc_source.csv 2>&1
ls -la gpurun_out/r02p*
