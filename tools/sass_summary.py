#!/usr/bin/env python
"""Instruction mix of every kernel in libssd_b200.so (cuobjdump -sass): counts of the opcodes that show what the
kernels are built from -- UBLKCP (TMA bulk copy), SYNCS (mbarrier), CREDUX / REDUX (warp reductions), FMNMX3,
MUFU, ATOMS / ATOMG / RED, BAR, UCGABAR (cluster barrier), MAPA / ST.E to shared::cluster -- as a markdown table.

    python tools/sass_summary.py > profiles/r02_sass_mix.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "single_shot_detection_b200", "libssd_b200.so")
WATCH = ["UBLKCP", "SYNCS", "CREDUX", "REDUX", "FMNMX3", "MUFU", "ATOMS", "ATOMG", "RED", "BAR", "UCGABAR", "MAPA",
         "LDS", "STS", "LDG", "STG", "SHFL", "VOTE", "FADD", "FMUL", "FFMA", "IMAD", "ACQBULK", "MEMBAR", "CCTL"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    name = None
    for ln in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("void ", "").replace("ssd::", "")
            kernels[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
        if m and name:
            kernels[name][m.group(1).split(".")[0]] += 1
            kernels[name]["_total"] += 1
    cols = [w for w in WATCH if any(c[w] for c in kernels.values())]
    print("# round 2: static SASS instruction mix per kernel (`cuobjdump -sass libssd_b200.so`, sm_100a)\n")
    print("Static counts (instructions in the binary, not executed): they show what each kernel is built from -- "
          "`UBLKCP` = TMA bulk copy, `SYNCS` = mbarrier, `CREDUX` = single-instruction warp max, `FMNMX3` = three-input "
          "max, `UCGABAR` / `MAPA` = cluster barrier / distributed shared memory.  No tensor-core opcodes: nothing on the "
          "path is a contraction.\n")
    print("| kernel | total | " + " | ".join(cols) + " |")
    print("|---|---|" + "---|" * len(cols))
    for k, c in kernels.items():
        if c["_total"] < 40:
            continue
        print(f"| `{k[:70]}` | {c['_total']} | " + " | ".join(str(c[w]) if c[w] else "" for w in cols) + " |")


if __name__ == "__main__":
    main()
