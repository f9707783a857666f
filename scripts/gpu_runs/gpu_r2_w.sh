#!/bin/bash
# CTA-pair GEMMs: relay vs direct remote-mbarrier signalling vs the single-CTA form (HuBERT encoder)
mkdir -p gpurun_out
# (the DISSC_TC_PAIR_DIRECT experiment was removed after this run: it hangs)
# DISSC_HUB_PAIR2=1 DISSC_TC_PAIR_DIRECT=1 timeout 300 python -m pytest tests/test_hubert_gpu.py -m gpu -q -x --no-header -p no:cacheprovider -s -k "varlen_batch or config4" 2>&1 | grep -E "passed|failed|rror|hubert|assert" | cut -c1-300 | head
echo "direct rc=${PIPESTATUS[0]}"
run() { env "$@" timeout 200 python scripts/bench_hubert.py 32 96000 8 2>&1 | grep "hubert encode\|rror" | sed "s/^/$* : /" | cut -c1-130; }
{
run DISSC_HUB_PAIR2=0
run DISSC_HUB_PAIR2=1
run DISSC_HUB_PAIR2=1 DISSC_TC_PAIR_DIRECT=1
run DISSC_HUB_PAIR2=0
} | tee gpurun_out/r02_hubert_pair2_ab.txt
