#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; cut -c1-260 gpurun_out/final_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_ref.json 2>> gpurun_out/final_bench.err; cut -c1-200 gpurun_out/final_bench_ref.json
timeout 300 python scripts/latency_config1.py > gpurun_out/final_latency.txt 2>&1; cut -c1-150 gpurun_out/final_latency.txt
timeout 300 python scripts/bench_hubert.py 32 96000 5 > gpurun_out/final_hubert.txt 2>&1; tail -2 gpurun_out/final_hubert.txt
timeout 600 python scripts/bench_pipeline.py > gpurun_out/final_pipeline_n1.txt 2>&1; grep config gpurun_out/final_pipeline_n1.txt | cut -c1-200
timeout 600 python scripts/ab_layers.py --rounds 2 cur: > gpurun_out/final_layers.txt 2>&1; tail -13 gpurun_out/final_layers.txt
