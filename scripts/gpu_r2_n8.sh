#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
echo rc=$?
python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench_n$N.json'))
print('value', d['ms_per_step'], d['value']/1e9, 'e2e', d['e2e']['ms_per_step'], d['e2e']['value']/1e9)
print('gathered', d['gathered'])
print({k:(round(v.get('ms_per_step',0),2), v.get('clips_per_s') or v.get('utterances_per_s')) for k,v in d['configs'].items() if isinstance(v,dict)})
PY
tail -3 gpurun_out/r02_bench_n$N.err
