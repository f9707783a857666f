#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hubert_gpu.py tests/test_pipeline_gpu.py -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -4
python scripts/bench_hubert.py 32 96000 8 2>&1 | head -1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; tail -2 gpurun_out/r2n_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2n_bench.json'))
print(d['ms_per_step'], d['sustained']['ms_per_step'], d['configs'].get('configs[0]'), d['configs'].get('error'))
print({k:(round(v.get('ms_per_step',0),2), v.get('clips_per_s') or v.get('utterances_per_s')) for k,v in d['configs'].items() if isinstance(v,dict) and 'ms_per_step' in v})
PY
