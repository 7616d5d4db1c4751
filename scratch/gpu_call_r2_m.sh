#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/m_dyg_launches.csv python scratch/dyg_probe.py 3 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/m_dyg_launches.csv')) if len(r)>10]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value'); 
data=rows[1:]
n=len(data)//3
last=data[-n:]
tot=0; agg=collections.OrderedDict()
for r in last:
    v=float(r[iv].replace(',','')); k=r[ik][:70]; tot+=v
    agg.setdefault(k,[0,0]); agg[k][0]+=v; agg[k][1]+=1
print('launches per forward', n, 'sum us', tot/1e3)
for k,(v,c) in sorted(agg.items(), key=lambda kv:-kv[1][0]): print(f'{v/1e3:8.1f} us  x{c:2d}  {k}')
PY
