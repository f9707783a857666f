#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_layers_gpu.py -m gpu -x -q -k "pair" > gpurun_out/exp9_pytest_pair.log 2>&1; echo "pair pytest rc=$?"
tail -12 gpurun_out/exp9_pytest_pair.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/exp9_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/exp9_pytest.log
timeout 900 python scripts/ab_layers.py --rounds 3 pair16off:DISSC_TC_PAIR16=0 pair16on:DISSC_TC_PAIR16=1 > gpurun_out/exp9_ab.txt 2>&1
grep "s4\.\|^s4\|^s3 \|^s0 \|TOTAL" gpurun_out/exp9_ab.txt
for r in 1 2; do for v in 0 1; do
  DISSC_TC_PAIR16=$v timeout 600 python bench.py --no-cpu-baseline > gpurun_out/exp9_bench_$v_$r.json 2> gpurun_out/exp9_bench.err
  python -c "import json;d=json.load(open('gpurun_out/exp9_bench_$v_$r.json'));print('PAIR16=$v run $r ms_per_step',round(d['ms_per_step'],3))"
done; done
