#!/bin/bash
# CTA-pair (cta_group::2) GEMMs in the HuBERT encoder: parity first (short timeout: a protocol error hangs), then timing
mkdir -p gpurun_out
export DISSC_HUB_PAIR2=1
timeout 600 python -m pytest tests/test_hubert_gpu.py -m gpu -q --no-header -p no:cacheprovider -s 2>&1 | grep -E "passed|failed|rror|hubert|assert|attention" | cut -c1-400 | head -30
python /tmp/ll.py gpurun_out/r02o_hubert_launches.csv 2>/dev/null | head -0
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02v_hubert_launches.csv python scripts/bench_hubert.py 32 96000 1 > /dev/null 2>&1
