#!/usr/bin/env python
"""Summarise an ncu --set full report (read with `ncu -i rep --page raw --csv`): one row per kernel."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
idx = {n: i for i, n in enumerate(h)}
def g(r, k, d=""):
    return r[idx[k]] if k in idx else d
stall = [k for k in h if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued")]
print("| kernel | us | grid x block | regs | warp-inst | dram rd MB | dram wr MB | sm busy % | issue active % | top stalls |")
print("|---|---|---|---|---|---|---|---|---|---|")
for r in rows[2:]:
    name = g(r, "Kernel Name").split("(")[0][-34:]
    st = []
    tot = 0
    for k in stall:
        try:
            v = float(r[idx[k]].replace(",", ""))
        except ValueError:
            v = 0
        tot += v
        st.append((v, k.replace("smsp__pcsamp_warps_issue_stalled_", "")))
    st.sort(reverse=True)
    tops = ", ".join(f"{n} {100*v/max(tot,1):.0f}%" for v, n in st[:3])
    def f(k, scale=1.0, fmt="{:.1f}"):
        try:
            return fmt.format(float(g(r, k).replace(",", "")) * scale)
        except ValueError:
            return "?"
    units = {"Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6, "Gbyte": 1e3}
    sc = units.get(rows[1][idx["dram__bytes_read.sum"]], 1.0)
    scw = units.get(rows[1][idx["dram__bytes_write.sum"]], 1.0)
    print(f"| {name} | {f('gpu__time_duration.sum', 1e-3 if rows[1][idx['gpu__time_duration.sum']]=='ns' else 1.0)} | "
          f"{g(r,'launch__grid_size')}x{g(r,'launch__block_size')} | {g(r,'launch__registers_per_thread')} | "
          f"{f('smsp__inst_executed.sum', 1e-6, '{:.2f}M')} | {f('dram__bytes_read.sum', sc, '{:.2f}')} | {f('dram__bytes_write.sum', scw, '{:.2f}')} | "
          f"{f('sm__throughput.avg.pct_of_peak_sustained_elapsed')} | {f('smsp__issue_active.avg.pct_of_peak_sustained_active')} | {tops} |")
