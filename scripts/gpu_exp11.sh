#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/exp11_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/exp11_pytest.log
timeout 900 python scripts/ab_layers.py --rounds 3 na2:DISSC_TC_NA=2 na3:DISSC_TC_NA=3 > gpurun_out/exp11_ab.txt 2>&1
grep "^s0 \|^s1 \|^s2 \|^s3 \|^ups \|TOTAL\|layer  " gpurun_out/exp11_ab.txt
grep "s1.rb2.c1.0\|s1.rb1.c1.0\|s1.rb0.c1.0\|s0.rb2.c1.0\|s0.rb1.c1.0\|ups.0\|ups.1" gpurun_out/exp11_ab.txt
for v in 2 3; do DISSC_TC_NA=$v timeout 600 python bench.py --no-cpu-baseline > gpurun_out/exp11_bench_$v.json 2> gpurun_out/exp11_bench.err
  python -c "import json;d=json.load(open('gpurun_out/exp11_bench_$v.json'));print('NA=$v ms_per_step',round(d['ms_per_step'],3))"; done
