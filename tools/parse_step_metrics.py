#!/usr/bin/env python
"""Last step of a tools/ncu_step.sh launch list: kernel, us, warp instructions, DRAM MB read / written."""
import csv, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]
ik, im, iv, iid = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value'), h.index('ID')
by = {}
for r in rows[1:]:
    by.setdefault(int(r[iid]), {'k': r[ik].split('(')[0][-40:]})[r[im]] = float(r[iv].replace(',', ''))
ids = sorted(by)
# the last step = everything after the last zero_kernel / first kernel of the step
start = max(i for i in ids if 'zero_kernel' in by[i]['k']) if any('zero_kernel' in by[i]['k'] for i in ids) else ids[-9]
tot_i = tot_t = 0
print("| kernel | us | warp-inst (M) | dram rd MB | dram wr MB |\n|---|---|---|---|---|")
for i in ids:
    if i < start: continue
    b = by[i]
    tot_i += b.get('smsp__inst_executed.sum', 0); tot_t += b.get('gpu__time_duration.sum', 0)
    print(f"| {b['k']} | {b.get('gpu__time_duration.sum',0)/1e3:.1f} | {b.get('smsp__inst_executed.sum',0)/1e6:.2f} | {b.get('dram__bytes_read.sum',0)/1e6:.2f} | {b.get('dram__bytes_write.sum',0)/1e6:.2f} |")
print(f"| total | {tot_t/1e3:.1f} | {tot_i/1e6:.2f} | | |")
