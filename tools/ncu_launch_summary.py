#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: ncu_launch_summary.py launches.csv [divide_by_calls]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
h = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[h]
agg = collections.OrderedDict()
for r in rows[h + 1:]:
    if len(r) < len(hdr): continue
    d = dict(zip(hdr, r))
    name = d['Kernel Name'].split('(')[0].replace('<unnamed>::', '').replace('void ', '')
    v = float(d['Metric Value'].replace(',', '')); u = d['Metric Unit']
    v = {'ns': v / 1e6, 'us': v / 1e3, 'ms': v, 's': v * 1e3}.get(u, v)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
for k, a in agg.items():
    print(f"{a[1] / div:9.3f} ms {100 * a[1] / tot:5.1f}%  x{a[0]:3d}  {k[:110]}")
print(f"{tot / div:9.3f} ms total")
