#!/bin/bash
# one-shot GPU experiment: parity tests, then A/B per-layer timings of the epilogue variants, then the bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/exp1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/exp1_pytest.log
tail -5 gpurun_out/exp1_pytest.log
DISSC_TC_FAST=0 DISSC_TC_EPW64=4 timeout 300 python scripts/profile_layers.py 64 300 > gpurun_out/exp1_layers_generic.txt 2>&1
DISSC_TC_FAST=1 DISSC_TC_EPW64=4 timeout 300 python scripts/profile_layers.py 64 300 > gpurun_out/exp1_layers_fast_epw4.txt 2>&1
DISSC_TC_FAST=1 DISSC_TC_EPW64=8 timeout 300 python scripts/profile_layers.py 64 300 > gpurun_out/exp1_layers_fast_epw8.txt 2>&1
tail -12 gpurun_out/exp1_layers_generic.txt gpurun_out/exp1_layers_fast_epw4.txt gpurun_out/exp1_layers_fast_epw8.txt
timeout 600 python bench.py > gpurun_out/exp1_bench.json 2> gpurun_out/exp1_bench.err; cat gpurun_out/exp1_bench.json | cut -c1-400
