#!/bin/bash
for i in 1 2; do
python scripts/bench_hubert.py 32 96000 8 2>&1 | head -1
DISSC_TC_CLUSTER2=1 python scripts/bench_hubert.py 32 96000 8 2>&1 | head -1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2k_hubert_launches.csv python scripts/bench_hubert.py 32 96000 1 > /dev/null 2>&1
python - <<'PY'
import csv,re,collections
lines=open('gpurun_out/r2k_hubert_launches.csv').read().splitlines()
st=next(i for i,l in enumerate(lines) if l.startswith('"ID"'))
rows=list(csv.DictReader(lines[st:]))
ids=[i for i,r in enumerate(rows) if 'hub_conv0_stats' in r['Kernel Name']]
rows=rows[ids[-1]:]
for r in rows:
    k=re.sub(r'^(void )?(dissc::)?','',r['Kernel Name']); k=re.sub(r'\(.*$','',k)
    v=float(r['Metric Value'].replace(',','')); u=r['Metric Unit']
    us=v*{'ns':1e-3,'us':1,'ms':1e3}.get(u,1)
    print(f"{k[:60]:60s} {us:9.1f} us")
PY
