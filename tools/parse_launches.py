#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel, launches and
the duration of the LAST instance (steady state), plus the sum."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hi]
ik, iv, iu = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    try:
        v = float(r[iv].replace(',', ''))
    except ValueError:
        continue
    if r[iu] == 'ns':
        v /= 1e3
    agg.setdefault(r[ik].split('(')[0][:70], []).append(v)
tot = sum(v[-1] for v in agg.values())
for k, v in agg.items():
    print(f"{k:70s} n={len(v):3d} last={v[-1]:9.2f} us  {100*v[-1]/tot:5.1f}%")
print(f"sum of last instances: {tot:.1f} us")
