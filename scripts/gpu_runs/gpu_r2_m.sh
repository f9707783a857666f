#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_generator_gpu.py tests/test_cli_gpu.py tests/test_integration_doc.py -m gpu -q -x -rf --no-header -p no:cacheprovider 2>&1 | tail -15
