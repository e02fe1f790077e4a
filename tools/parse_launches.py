#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list as a markdown table: per kernel the number of
launches, the median duration and the kernel's share of the summed device time."""
import collections
import csv
import statistics
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')) if len(r) > 10]
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
    agg.setdefault(r[ik].split('(')[0].replace('void ', '').replace('ssd::', '')[:60], []).append(v)
tot = sum(sum(v) for v in agg.values())
print("| kernel | launches | median us | share of the summed device time |")
print("|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"| `{k}` | {len(v)} | {statistics.median(v):.1f} | {100 * sum(v) / tot:.1f} % |")
print(f"\nsum over {sum(len(v) for v in agg.values())} launches: {tot / 1e3:.2f} ms")
