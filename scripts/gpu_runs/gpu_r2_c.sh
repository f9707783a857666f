#!/bin/bash
mkdir -p gpurun_out
python -m dissc_b200.build >/dev/null 2>&1
python scripts/ab_layers.py --rounds 4 base: na4:DISSC_TC_NA=4 split:DISSC_TC_SPLIT256=1,DISSC_TC_SINGLE_ACC=0 \
   split3:DISSC_TC_SPLIT256=1,DISSC_TC_SINGLE_ACC=0,DISSC_TC_NA=3 split4:DISSC_TC_SPLIT256=1,DISSC_TC_SINGLE_ACC=0,DISSC_TC_NA=4 \
   > gpurun_out/r2c_ab.txt 2>&1
tail -14 gpurun_out/r2c_ab.txt
DISSC_TC_SPLIT256=1 DISSC_TC_SINGLE_ACC=0 python -m pytest tests/test_generator_gpu.py -q -s -k "rows_vs_oracle or golden" -p no:cacheprovider 2>&1 | grep -E "max-abs|passed|failed" > gpurun_out/r2c_split_acc.log
cat gpurun_out/r2c_split_acc.log
python bench.py --steps 20 --warmup 5 --no-configs > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c_bench.json'))
print('bench', d['ms_per_step'], d['e2e']['ms_per_step'], d['gathered']['ms_per_step'], d['clocks'])
PY
