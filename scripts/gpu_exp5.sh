#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/exp5_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/exp5_pytest.log
timeout 300 python scripts/latency_config1.py > gpurun_out/exp5_latency.txt 2>&1; cat gpurun_out/exp5_latency.txt | tail -5
timeout 600 python scripts/bench_pipeline.py > gpurun_out/exp5_pipeline.txt 2>&1; tail -5 gpurun_out/exp5_pipeline.txt
timeout 300 python scripts/bench_hubert.py 32 96000 5 > gpurun_out/exp5_hubert.txt 2>&1; tail -2 gpurun_out/exp5_hubert.txt
timeout 600 python bench.py > gpurun_out/exp5_bench.json 2> gpurun_out/exp5_bench.err; cut -c1-300 gpurun_out/exp5_bench.json
