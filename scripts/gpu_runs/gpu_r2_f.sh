#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_layers_gpu.py tests/test_generator_gpu.py tests/test_hubert_gpu.py -m gpu -q -x -rf --no-header -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/r2f_tests.log
echo "tests rc=$?"; tail -5 gpurun_out/r2f_tests.log
timeout 600 python scripts/ab_layers.py --rounds 3 cl2: nocl2:DISSC_TC_CLUSTER2=0 > gpurun_out/r2f_ab.txt 2>&1
grep -E "^s0 |^s1 |^s2 |^s3 |^s4 |TOTAL|^ups |^conv_pre " gpurun_out/r2f_ab.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-configs > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
DISSC_TC_CLUSTER2=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-configs > gpurun_out/r2f_bench_nocl2.json 2>> gpurun_out/r2f_bench.err
python - <<'PY'
import json
for f in ('r2f_bench','r2f_bench_nocl2'):
    try:
        d=json.load(open(f'gpurun_out/{f}.json')); print(f, d['ms_per_step'], d['e2e']['ms_per_step'], d['gathered']['ms_per_step'], d['clocks']['sm_mhz'])
    except Exception as e: print(f, 'failed', e)
PY
