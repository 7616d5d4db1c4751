#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_nn.py -q -x 2>&1 | tail -2
python bench_rows.py --rows dygformer 2>&1 | cut -c1-200
for P in dyg tgat; do
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 600 --csv --log-file gpurun_out/n_${P}_launches.csv python scratch/${P}_probe.py 3 > /dev/null 2>&1
python - $P <<'PY'
import csv, collections, sys
P=sys.argv[1]
rows=[r for r in csv.reader(open(f'gpurun_out/n_{P}_launches.csv')) if len(r)>10]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value'); im=hdr.index('Metric Name'); iid=hdr.index('ID')
t=collections.OrderedDict()
for r in rows[1:]:
    d=t.setdefault(r[iid],{'k':r[ik][:64]}); d[r[im]]=float(r[iv].replace(',',''))
L=list(t.values()); n=len(L)//3; last=L[-n:]
agg=collections.OrderedDict()
for d in last:
    a=agg.setdefault(d['k'],[0,0,0]); a[0]+=d['gpu__time_duration.sum']; a[1]+=1; a[2]+=d.get('smsp__inst_executed.sum',0)
print(P,'launches per forward', n, 'sum us', sum(a[0] for a in agg.values())/1e3)
for k,(v,c,i) in sorted(agg.items(), key=lambda kv:-kv[1][0]): print(f'{v/1e3:8.1f} us  x{c:2d}  {i/1e6:8.2f} Minst  {k}')
PY
done
