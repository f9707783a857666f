#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
   --log-file gpurun_out/r01_d_hubert_launches.csv python scripts/bench_hubert.py 32 96000 1 > gpurun_out/ncu4.log 2>&1
tail -3 gpurun_out/ncu4.log
python - <<'PY'
import csv
lines=open('gpurun_out/r01_d_hubert_launches.csv').read().splitlines()
start=next(i for i,l in enumerate(lines) if l.startswith('"ID"'))
rows=list(csv.DictReader(lines[start:]))
L={}
for r in rows:
    d=L.setdefault(int(r['ID']),{'k':r['Kernel Name'],'grid':r.get('Grid Size','')})
    if r['Metric Name']=='gpu__time_duration.sum':
        v=float(r['Metric Value'].replace(',',''));u=r['Metric Unit']
        d['us']=v*{'ns':1e-3,'us':1,'ms':1e3}.get(u,1)
ids=sorted(L)
# last forward = last third of hub launches; print the final 70 launches
for i in ids[-75:]:
    print(i, L[i]['k'][:60], L[i]['grid'], round(L[i].get('us',0),1))
PY
