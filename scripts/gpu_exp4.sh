#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/exp4_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/exp4_pytest.log
timeout 900 python scripts/ab_layers.py --rounds 2 cur: > gpurun_out/exp4_ab.txt 2>&1
tail -14 gpurun_out/exp4_ab.txt
timeout 600 python bench.py > gpurun_out/exp4_bench.json 2> gpurun_out/exp4_bench.err; cut -c1-300 gpurun_out/exp4_bench.json
