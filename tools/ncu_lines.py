#!/usr/bin/env python
"""Hottest CUDA source lines of one kernel in an ncu report (needs --import-source on, -lineinfo).
usage: tools/ncu_lines.py report.ncu-rep kernel-regex [top-n]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
fname, hdr, out, nfunc = None, None, [], 0
first = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        if first is None: first = r[1]
        continue
    if r[0] == "Line No":
        hdr = r; isamp = hdr.index("# Samples"); iex = hdr.index("Instructions Executed"); continue
    if hdr and r[0].isdigit() and len(r) > iex:
        try:
            s = int(r[isamp]); e = int(r[iex])
        except ValueError:
            continue
        if s or e:
            out.append((fname, int(r[0]), r[1].strip(), s, e))
# several instances of the kernel repeat the same lines: aggregate
agg = {}
for f, ln, src, s, e in out:
    k = (f, ln)
    a = agg.setdefault(k, [src, 0, 0]); a[1] += s; a[2] += e
ts = sum(a[1] for a in agg.values()) or 1; te = sum(a[2] for a in agg.values()) or 1
print(f"{first[:80] if first else kern}: samples {ts} warp-inst {te}")
key_i = 2 if len(sys.argv) > 4 and sys.argv[4] == "inst" else 1
top = sorted(agg.items(), key=lambda kv: -kv[1][key_i])[:ntop]
for (f, ln), (src, s, e) in sorted(top):
    print(f"{f}:{ln:<5} {100*s/ts:5.1f}% samp {100*e/te:5.1f}% inst  {src[:105]}")
