#!/bin/bash
# GPU box, round 2 step b: tests, smoke, bench (new line), accuracy / prefetch experiments.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rf --no-header -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/r2b_tests.log
tail -4 gpurun_out/r2b_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_smoke.log 2>&1; tail -4 gpurun_out/r2b_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
cut -c1-300 gpurun_out/r2b_bench.json; tail -3 gpurun_out/r2b_bench.err
DISSC_TC_SINGLE_ACC=0 python -m pytest tests/test_generator_gpu.py -q -s -k "rows_vs_oracle" -p no:cacheprovider 2>&1 | grep -E "max-abs|passed|failed" > gpurun_out/r2b_dualacc.log
cat gpurun_out/r2b_dualacc.log
python scripts/ab_layers.py --rounds 2 base: na3:DISSC_TC_NA=3 na4:DISSC_TC_NA=4 > gpurun_out/r2b_ab_na.txt 2>&1
tail -14 gpurun_out/r2b_ab_na.txt
