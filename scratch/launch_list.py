"""Per-launch list of the last of N identical passes in an ncu --csv launch log."""
import collections
import csv
import sys

path, passes = sys.argv[1], int(sys.argv[2])
rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
h = rows[0]
ki, mi, vi, idi, gi = (h.index(c) for c in ('Kernel Name', 'Metric Name', 'Metric Value', 'ID', 'Grid Size'))
per = collections.OrderedDict()
for r in rows[1:]:
    e = per.setdefault(r[idi], [r[ki], 0.0, 0.0, r[gi]])
    v = float(r[vi].replace(',', ''))
    if 'time' in r[mi]:
        e[1] = v
    else:
        e[2] = v
ids = list(per)
n = len(ids) // passes
print('launches per pass', n, 'sum us', sum(per[i][1] for i in ids[-n:]) / 1e3)
for i in ids[-n:]:
    k, t, c, g = per[i]
    print(f'{t / 1e3:7.1f} us {c / 1e6:7.2f} Minst grid {g:>14s}  {k[:64]}')
