#!/bin/bash
# half-tap weight slots (three activation buffers + three weight slots) for the HuBERT GEMMs: parity, then A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_hubert_gpu.py -m gpu -q -x --no-header -p no:cacheprovider -s 2>&1 | grep -E "passed|failed|rror|hubert|assert|attention" | cut -c1-300 | head -20
run() { env "$@" timeout 200 python scripts/bench_hubert.py 32 96000 8 2>&1 | grep "hubert encode\|rror" | sed "s/^/$* : /" | cut -c1-130; }
{
run DISSC_TC_WSPLIT=1
run DISSC_TC_WSPLIT=0
run DISSC_TC_WSPLIT=1
run DISSC_TC_WSPLIT=0
} | tee gpurun_out/r02_hubert_wsplit_ab.txt
