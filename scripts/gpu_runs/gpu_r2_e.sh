#!/bin/bash
mkdir -p gpurun_out
python scripts/ab_layers.py --rounds 3 base: halfw:DISSC_EXP_HALFW=1 quarterw:DISSC_EXP_HALFW=2 > gpurun_out/r2e_ab.txt 2>&1
grep -E "^s0\.rb[02]\.c1\.0|^s1\.rb[02]\.c[12]\.0|^s0 |^s1 |^s2 |^s3 |^s4 |TOTAL|^ups|conv_pre " gpurun_out/r2e_ab.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2e_hubert_launches.csv python scripts/bench_hubert.py 32 96000 1 > gpurun_out/r2e_hubert_ncu.log 2>&1
python - <<'PY'
import csv,re,collections
lines=open('gpurun_out/r2e_hubert_launches.csv').read().splitlines()
st=next(i for i,l in enumerate(lines) if l.startswith('"ID"'))
rows=list(csv.DictReader(lines[st:]))
# last forward only: take launches after the last conv0_stats kernel
ids=[i for i,r in enumerate(rows) if 'hub_conv0_stats' in r['Kernel Name']]
rows=rows[ids[-1]:]
agg=collections.OrderedDict()
for r in rows:
    k=re.sub(r'^(void )?(dissc::)?','',r['Kernel Name']); k=re.sub(r'\(.*$','',k)
    v=float(r['Metric Value'].replace(',','')); u=r['Metric Unit']
    us=v*{'ns':1e-3,'us':1,'ms':1e3,'nsecond':1e-3,'usecond':1,'msecond':1e3}.get(u,1)
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=us
tot=sum(a[1] for a in agg.values())
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1]): print(f"{k[:90]:90s} {a[0]:3d} {a[1]/1e3:8.3f} ms {100*a[1]/tot:5.1f}%")
print('total', tot/1e3)
PY
