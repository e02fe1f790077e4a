#!/usr/bin/env python
"""Summarise the ncu CSVs of the exchange kernels (tools/multi_gpu.sh N tag ncu): per kernel the median device time,
warp instructions and NVLink bytes sent / received per launch.  usage: tools/ncu_exchange_summary.py <csv prefix>"""
import csv
import glob
import statistics
import sys

prefix = sys.argv[1]
vals = {}
for path in glob.glob(prefix + "*.csv"):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    if not rows:
        continue
    h = rows[0]
    ik, im, iv, ig, ib = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("Block Size")
    for r in rows[1:]:
        k = r[ik].split("(")[0]
        vals.setdefault((k, r[ig], r[ib]), {}).setdefault(r[im], []).append(float(r[iv].replace(",", "")))
print("| kernel | grid x block | launches profiled | median us | warp instructions | NVLink bytes sent per launch | NVLink bytes received |")
print("|---|---|---|---|---|---|---|")
for (k, g, b), m in sorted(vals.items()):
    med = lambda n: statistics.median(m[n]) if n in m else float("nan")   # noqa: E731
    n = max(len(v) for v in m.values())
    print(f"| `{k}` | {g} x {b} | {n} | {med('gpu__time_duration.sum') / 1e3:.1f} | {med('smsp__inst_executed.sum'):,.0f} | "
          f"{med('nvltx__bytes.sum'):,.0f} | {med('nvlrx__bytes.sum'):,.0f} |")
