#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/exp8_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/exp8_pytest.log
for r in 1 2 3; do
  timeout 600 python bench.py --no-cpu-baseline > gpurun_out/exp8_bench_$r.json 2> gpurun_out/exp8_bench.err
  python -c "import json;d=json.load(open('gpurun_out/exp8_bench_$r.json'));print('run $r ms_per_step',round(d['ms_per_step'],3),'e2e ms',round(d['e2e']['ms_per_step'],3),d['clocks']['sm_mhz'])"
done
timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:resblock_pair64 --launch-skip 6 --launch-count 1 python scripts/one_forward.py 64 300 1 2>&1 | grep -i "inst_executed\|duration" 
