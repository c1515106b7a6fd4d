#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02h_launches_c5_share0of8.csv python scripts/c5_share_once.py 0 8 > gpurun_out/r02h_ncu.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02h_launches_c5_share0of8.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
cols=rows[hdr]; data=rows[hdr+1:]
iN=cols.index('Kernel Name'); iV=cols.index('Metric Value'); iU=cols.index('Metric Unit')
half=len(data)//2
agg={}
for r in data[half:]:
    v=float(r[iV].replace(',','')); u=r[iU]
    ms = v/1e6 if u=='ns' else (v/1e3 if u=='us' else v)
    k=r[iN][:70]
    a=agg.setdefault(k,[0,0.]); a[0]+=1; a[1]+=ms
for k,(c,ms) in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    print(f"{ms:9.3f} ms  x{c:3d}  {k}")
PY
