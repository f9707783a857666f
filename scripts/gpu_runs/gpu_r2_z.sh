#!/bin/bash
# vocoder: weight-stage size of the streamed-weight layers (32 KB default vs 64 KB / 16 KB), per-layer A/B on one box
mkdir -p gpurun_out
timeout 1200 python scripts/ab_layers.py --rounds 3 s32: s64:DISSC_TC_STAGE_BYTES=65536 s16:DISSC_TC_STAGE_BYTES=16384 > gpurun_out/r02_ab_stage_bytes.txt 2>&1
grep -E "^s0 |^s1 |^s2 |^s3 |^s4 |TOTAL|^ups |^conv_pre |label|variant" gpurun_out/r02_ab_stage_bytes.txt | head -20
