#!/bin/bash
# GPU box: full -m gpu test suite (all failures listed), then the default bench line.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rf --no-header -p no:cacheprovider 2>&1 | tail -80 > gpurun_out/${1:-r2}_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${1:-r2}_bench.json 2> gpurun_out/${1:-r2}_bench.err
tail -5 gpurun_out/${1:-r2}_tests.log
cat gpurun_out/${1:-r2}_bench.json | cut -c1-600
