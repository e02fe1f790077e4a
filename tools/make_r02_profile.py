#!/usr/bin/env python
"""Assemble profiles/r02_*.md from what tools/profile_all.sh and tools/multi_gpu.sh left in gpurun_out/.
usage: tools/make_r02_profile.py <tag>"""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
WL = ["ssd300_voc_b32", "ssd_mb2_coco_b64", "ssd512_coco_b32", "retina500_coco_b32", "m2det512_coco_b256"]


def last_json(path):
    try:
        return json.loads(open(path).read().strip().splitlines()[-1])
    except Exception:  # noqa: BLE001
        return None


def rnd(x, n=1):
    return "-" if x is None else f"{x:.{n}f}"


out = [f"# round 2: every BASELINE configuration on one B200 (`tools/profile_all.sh {tag}`)\n",
       "Bench lines (`bench.py --workload W --no-config5 --no-cpu-baseline --no-e2e`, final code of the round; CUDA events, "
       "inputs rotating over sets larger than L2) and, per configuration, the `ncu --set full` table of the last eager step "
       "(`tools/ncu_step.sh`: cold cache, serialised launches -- compare shares, not absolute times).\n",
       "| workload | img/s (8 steps in flight) | us/step | us/step serial | pass 1 alone: us, fraction of the copy peak | sampler kernel alone: fraction | step fraction (SURVEY 8d) in flight / serial | kernels in context (us) |",
       "|---|---|---|---|---|---|---|---|"]
for w in WL:
    d = last_json(os.path.join(G, f"{tag}_bench_{w}.json"))
    if not d:
        out.append(f"| {w} | (no run) | | | | | | |")
        continue
    r = d["roofline"]
    other = list(r.get("other_streaming_kernels", {}).values())
    ku = d.get("kernels_us") or {}
    out.append(f"| {w} | {d['value']:,.0f} | {rnd(1e3 * d['ms_per_step'])} | {rnd(1e3 * d['serial']['ms_per_step'])} | "
               f"{rnd(r['us_per_launch'])}, {rnd(r['frac'], 3)} | {rnd(other[0]['frac'], 3) if other else '-'} | "
               f"{rnd(r['step']['in_flight']['frac'], 3)} / {rnd(r['step']['serial']['frac'], 3)} | "
               + ", ".join(f"{k} {v:.1f}" for k, v in ku.items()) + " |")
out.append("")
for w in WL:
    for kind, title in (("full.md", "`ncu --set full`, last eager step"), ("launches.md", "launch list of the same step (`--metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_*`)")):
        path = os.path.join(G, f"{tag}_ncu_{w}.{kind}")
        if os.path.exists(path):
            out.append(f"## {w}: {title}\n")
            out.append(open(path).read().strip() + "\n")
open(os.path.join(ROOT, "profiles", "r02_ncu_all_configs.md"), "w").write("\n".join(out) + "\n")

# multi-GPU
rows = []
for f in sorted(glob.glob(os.path.join(G, f"{tag}_bench_n*.json"))):
    d = last_json(f)
    if d:
        rows.append(d)
one = last_json(os.path.join(G, f"{tag}_bench_ssd300_voc_b32.json"))
mg = [f"# round 2: multi-GPU (`tools/multi_gpu.sh N {tag}`, one box, one rank per GPU)\n",
      "Weak scaling of the headline step (32 images per GPU, 8 steps in flight, the exchange kernels are the first / last nodes of every "
      "step graph and also run at N = 1), and BASELINE configs[4] as stated: M2Det-512 b256 sharded by image, with the "
      "gathered buffer of every rank compared bit for bit with a single-GPU run of the same global batch.\n",
      "| N | img/s | us/step | vs N = 1 | serial us/step | e2e img/s (host buffers) | config 5 strong: img/s, us/step, images per GPU, gather parity | exchange_check |",
      "|---|---|---|---|---|---|---|---|"]
base = one["ms_per_step"] if one else None
if one:
    mg.append(f"| 1 | {one['value']:,.0f} | {rnd(1e3 * one['ms_per_step'])} | 1.00 | {rnd(1e3 * one['serial']['ms_per_step'])} | (see BENCH line) | | |")
for d in sorted(rows, key=lambda x: x["n_gpus"]):
    n = d["n_gpus"]
    c5 = d.get("config5_strong") or {}
    chk = "-"
    try:
        chk = open(os.path.join(G, f"{tag}_xchk_n{n}.log")).read().strip().splitlines()[-1]
    except Exception:  # noqa: BLE001
        pass
    eff = f"{base / d['ms_per_step']:.3f}" if base else "-"
    e2e = d["e2e"]["value"]
    mg.append(f"| {n} | {d['value']:,.0f} | {rnd(1e3 * d['ms_per_step'])} | {eff} | {rnd(1e3 * d['serial']['ms_per_step'])} | "
              f"{e2e:,.0f} | " + (f"{c5.get('value', 0):,.0f}, {rnd(1e3 * c5.get('ms_per_step', 0))}, {c5.get('images_per_gpu')}, {c5.get('gather_parity')}" if c5 else "-")
              + f" | {chk} |")
mg.append("")
mg.append("## ncu on the exchange kernels (N = 2, `tools/multi_gpu.sh 2 <tag> ncu`)\n")
mg.append("Kernel replay fails (`==ERROR== UnknownError`, `Failed to profile \"exchange_open_kernel\"`: Nsight Compute cannot save / restore the "
          "CUDA-IPC mappings the kernels write through); APPLICATION replay works (`--replay-mode application`, one run of "
          "`tools/exchange_check.py` per metric group, shard shape of the headline step: 32 images per rank, T = 200).  This VM's "
          "`nvidia-smi nvlink -gt d` counters read N/A, the `nvltx / nvlrx` counters of ncu do count:\n")
xs = os.path.join(G, f"{tag}_ncu_exchange_n2.md")
if os.path.exists(xs):
    mg.append(open(xs).read().strip() + "\n")
else:
    mg.append("(no capture in this run)\n")
mg.append("Algorithmic bytes: every image row (T x 24 + 20 bytes padded to 16: 4 832 at T = 200) is read once and stored `world` times with "
          "16-byte stores, plus one 8-byte release flag per (rank, row) and one 8-byte ack per (slot, rank) from the open kernel; at N = 8 "
          "and 32 images per GPU that is 1.08 MB per step and GPU over NVLink.  The times above are ncu's serialised, profiler-attached "
          "launches (the pack kernel's spin on its peers' acks included); in the step graph the exchange is hidden behind the other "
          "steps in flight -- the scaling column above is its real cost.\n")
# e2e: what the platform's host -> device path delivers (tools/h2d_probe.py) and the pipeline depth sweep
probe = os.path.join(G, f"{tag}_h2d_probe_n8.log")
if os.path.exists(probe):
    mg.append("## e2e scaling: the platform's H2D ceiling (`tools/h2d_probe.py` under torchrun, 8 ranks, same box type)\n")
    mg.append("The e2e leg copies 27.9 MB of pinned logits + locs per batch and rank inside its timed region.  Pinned-memory copies alone, "
              "1 / 2 / 4 / 8 ranks copying at once (the others idle):\n")
    mg.append("```")
    mg.extend(l.rstrip() for l in open(probe) if "copying" in l)
    mg.append("```\n")
    dl = []
    for d in (2, 4, 6):
        j = last_json(os.path.join(G, f"{tag}_e2e_depth{d}.json"))
        if j:
            dl.append(f"depth {d}: {j['e2e']['value']:,.0f} img/s")
    mg.append("So four GPUs of this VM share ~116 GB/s and eight ~239 GB/s of host -> device bandwidth: H2D ALONE caps the e2e metric at "
              "63 k / 126 k / 132 k / 274 k img/s for N = 1 / 2 / 4 / 8, i.e. at 4.3x the N = 1 value at N = 8 -- the 6x of linear scaling is "
              "not reachable on this host.  Measured e2e: 59 k / 119 k / 128 k / 199 k = 0.94 / 0.94 / 0.97 / 0.73 of that ceiling.  The N = 8 gap "
              "is not the software pipeline's depth (`bench.py --gpus 8 --e2e-depth D`: " + "; ".join(dl) + "): eight ranks also share 32 host "
              "cores for the ground-truth packing, the graph launches and the D2H read-backs.\n")
open(os.path.join(ROOT, "profiles", "r02_multi_gpu.md"), "w").write("\n".join(mg) + "\n")
with open(os.path.join(ROOT, "profiles", "r02_sass_mix.md"), "w") as f:
    f.write(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_summary.py")], capture_output=True, text=True).stdout)
print("written")
