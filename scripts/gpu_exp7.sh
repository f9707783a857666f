#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/exp7_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/exp7_pytest.log
for r in 1 2; do
for v in 0 1; do
  DISSC_PDL=$v timeout 600 python bench.py --no-cpu-baseline > gpurun_out/exp7_bench_pdl${v}_$r.json 2> gpurun_out/exp7_bench.err
  python -c "import json;d=json.load(open('gpurun_out/exp7_bench_pdl${v}_$r.json'));print('PDL=$v run $r ms_per_step',round(d['ms_per_step'],3),'e2e ms',round(d['e2e']['ms_per_step'],3),d['clocks']['sm_mhz'])"
done; done
DISSC_PDL=0 timeout 300 python scripts/latency_config1.py 2>&1 | tail -3 | cut -c1-120
DISSC_PDL=1 timeout 300 python scripts/latency_config1.py 2>&1 | tail -3 | cut -c1-120
