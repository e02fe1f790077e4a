#!/usr/bin/env python
"""profiles/ncu_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per launch of the roofline kernel
(score_pass1_kernel) out of an `ncu --set full` report, keyed by workload, together with the digest of the kernel
sources the capture saw -- bench.py quotes the number only while the sources still hash to it.

    python tools/ncu_traffic.py <report.ncu-rep> <workload> [kernel-regex]
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    rep, workload = sys.argv[1], sys.argv[2]
    kern = sys.argv[3] if len(sys.argv) > 3 else "score_pass1"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    head = rows[0]
    col = {n: i for i, n in enumerate(head)}
    vals = []
    for r in rows[2:]:
        if kern not in r[col["Kernel Name"]]:
            continue
        rd = float(r[col["dram__bytes_read.sum"]].replace(",", ""))
        wr = float(r[col["dram__bytes_write.sum"]].replace(",", ""))
        unit_rd, unit_wr = rows[1][col["dram__bytes_read.sum"]], rows[1][col["dram__bytes_write.sum"]]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        vals.append(rd * scale.get(unit_rd, 1.0) + wr * scale.get(unit_wr, 1.0))
    if not vals:
        raise SystemExit(f"no launch of {kern} in {rep}")
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    table = json.load(open(path)) if os.path.exists(path) else {}
    table[workload] = {"traffic_bytes": sorted(vals)[len(vals) // 2], "launches": len(vals), "kernel": kern,
                       "source_digest": bench.source_digest(), "source": f"profiles/ ({os.path.basename(rep)})"}
    with open(path, "w") as f:
        json.dump(table, f, indent=1, sort_keys=True)
    print(json.dumps(table[workload]))


if __name__ == "__main__":
    main()
