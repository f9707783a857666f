#!/bin/bash
# Final round-2 refresh: full GPU tests, smoke, the bench line (+ reference arm), per-layer times, HuBERT bench + launch list.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rf --no-header -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/r02_tests.log
tail -3 gpurun_out/r02_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -3 gpurun_out/r02_smoke.log
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench.json 2>> gpurun_out/r02_bench.err
python scripts/ab_layers.py --rounds 2 cur: > gpurun_out/r02_layers_B64_T300.txt 2>&1
tail -12 gpurun_out/r02_layers_B64_T300.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_hubert_launches.csv python scripts/bench_hubert.py 32 96000 1 > /dev/null 2>&1
python scripts/bench_hubert.py 32 96000 8 > gpurun_out/r02_hubert_bench.txt 2>&1; cat gpurun_out/r02_hubert_bench.txt
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench.json'))
print('value', d['ms_per_step'], d['value']/1e6, 'e2e', d['e2e']['ms_per_step'], 'gathered', d['gathered']['ms_per_step'], d['clocks'])
print(d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['hbm']['frac'], d['roofline']['forward']['tensor']['frac'], d['roofline']['forward']['hbm']['frac'])
print({k:(round(v.get('ms_per_step',0),2), v.get('clips_per_s') or v.get('utterances_per_s')) for k,v in d['configs'].items() if isinstance(v,dict)})
PY
