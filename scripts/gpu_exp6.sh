#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hubert_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q > gpurun_out/exp6_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/exp6_pytest.log
timeout 300 python scripts/bench_hubert.py 32 96000 5 > gpurun_out/exp6_hubert.txt 2>&1; tail -2 gpurun_out/exp6_hubert.txt
timeout 600 python scripts/bench_pipeline.py > gpurun_out/exp6_pipeline.txt 2>&1; grep config gpurun_out/exp6_pipeline.txt | cut -c1-360
