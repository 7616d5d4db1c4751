#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_attn_folded.py tests/test_gpu_nn.py tests/test_gpu_lazy_edge_x.py -q -x > gpurun_out/x_tests.log 2>&1; tail -25 gpurun_out/x_tests.log
python bench_rows.py --rows tgat 2>&1 | cut -c1-400
python bench_configs.py --config 3 2>/dev/null | cut -c1-300
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 200 --csv --log-file gpurun_out/x_tgat_launches.csv python scratch/tgat_probe.py 3 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/x_tgat_launches.csv') if l.startswith('"')))
h = rows[0]; ki, mi, vi = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value')
idi = h.index('ID')
per = collections.OrderedDict()
for r in rows[1:]:
    per.setdefault(r[idi], [r[ki], 0, 0])
    if 'time' in r[mi]: per[r[idi]][1] = float(r[vi].replace(',', ''))
    else: per[r[idi]][2] = float(r[vi].replace(',', ''))
ids = list(per)
n = len(ids) // 3
last = ids[-n:]
agg = collections.OrderedDict()
for i in last:
    k, t, c = per[i]
    a = agg.setdefault(k[:70], [0, 0, 0]); a[0] += t; a[1] += 1; a[2] += c
print('launches per forward', n, 'sum us', sum(a[0] for a in agg.values()) / 1e3)
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f'{a[0]/1e3:8.1f} us  x{a[1]:2d} {a[2]/1e6:8.2f} Minst  {k}')
PY
