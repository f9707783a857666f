#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_layers_gpu.py -m gpu -x -q -k "pair" > gpurun_out/exp2_pytest_pair.log 2>&1; echo "pair pytest rc=$?"
tail -15 gpurun_out/exp2_pytest_pair.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/exp2_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/exp2_pytest.log
timeout 900 python scripts/ab_layers.py --rounds 3 unfused64:DISSC_TC_PAIR64=0 pair64:DISSC_TC_PAIR64=1 > gpurun_out/exp2_ab.txt 2>&1
tail -14 gpurun_out/exp2_ab.txt
timeout 600 python bench.py > gpurun_out/exp2_bench.json 2> gpurun_out/exp2_bench.err; cut -c1-300 gpurun_out/exp2_bench.json
