#!/usr/bin/env python
"""Benchmark of the anchor pipeline (BASELINE.json metric: images/sec target-assign+NMS, SSD300 b32).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference] [--in-flight F]

A step = one pass of the hot path over one batch of synthetic input:
    encode_ground_truth -> sampler -> to_centroids + encode_box (in place) -> postprocess
    [-> exchange of detections + statistics between the ranks, N > 1].

ours:       `value`  = device-resident inputs, every step replayed from a CUDA graph, F consecutive steps in
                       flight on F streams (default 4), timed with CUDA events (max over ranks);
            `serial` = the same K steps strictly one after the other on one stream;
            `e2e`    = the same step through the reference-shaped Python API with HOST buffers:
                       pinned scores/locs and the ground-truth list are copied H2D and the padded
                       detections + counts + statistics are read back D2H inside the timed region;
            `roofline` = the logit-streaming kernel of the step (score_pass1, with the sampler's
                       criterion output) timed alone with CUDA events against MEASURED_PEAKS.json,
                       the sampler's own streaming kernel beside it;
            `cpu_baseline` = the CPU oracle (same torch CPU ops as the reference) on a bounded sample.
reference:  the reference's CPU algorithm (oracle port, torch CPU ops + torchvision NMS) on the
            host cores, same metric / config.

Inputs rotate over several independent input sets whose total size exceeds the 126 MB L2, so a
timed iteration never finds its logits in cache.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

from single_shot_detection_b200 import workloads as wl  # noqa: E402

L2_BYTES = 126 * 1024 * 1024


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--workload", default=wl.HEADLINE)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample-images", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--in-flight", type=int, default=4,
                    help="step graphs in flight on as many streams (1 = strictly one step after the other)")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi while the timed region runs
# --------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [x for x in sm if x >= 0.5 * max(sm)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm on the host cores (oracle port)
# --------------------------------------------------------------------------------------------
def time_cpu_oracle(w: wl.Workload, images: int, steps: int, warmup: int, threads: int):
    from oracle import anchor_pipeline_oracle as ora
    torch.set_num_threads(threads)
    anchors, gt, scores, locs = wl.make_inputs(w, seed=23, batch=images)
    cfg = w.cfg()
    for _ in range(warmup):
        ora.run_step(gt, anchors, scores, locs, cfg, canonical=False, use_torchvision=True)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        ora.run_step(gt, anchors, scores, locs, cfg, canonical=False, use_torchvision=True)
        times.append(time.perf_counter() - t0)
    return images / min(times), images / (sum(times) / len(times)), sum(times)


def run_reference(args):
    """--impl reference: rank 0 only, CPU."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = wl.WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    images = min(args.cpu_sample_images, w.batch)
    steps = max(1, min(args.steps, 5))
    warm = max(1, min(args.warmup, 1))
    best, mean, total = time_cpu_oracle(w, images, steps, warm, cores)
    anchors = wl.build_anchors(w)
    sample = (f"{images} of {w.batch} images per step, {steps} steps + {warm} warm-up, oracle port of the reference "
              f"(torch CPU ops + torchvision.ops.nms), {cores} torch threads")
    line = {
        "impl": "reference", "metric": "images/sec target-assign+NMS", "value": mean, "unit": "images/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": 1e3 * images / mean,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w.name, "anchors": int(anchors.shape[0]), "score_cols": w.num_score_cols,
                   "batch": w.batch, "sample_images": images},
        "cpu_baseline": {"value": mean, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample,
                         "best": best},
        "e2e": {"value": mean, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from single_shot_detection_b200 import _native as N
    from single_shot_detection_b200 import sharding
    from single_shot_detection_b200.pipeline import AnchorPipeline, matched_stats
    from single_shot_detection_b200.target_assigner import pack_ground_truth
    from single_shot_detection_b200 import sampler as S

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    N.require_device()

    w = wl.WORKLOADS[args.workload]
    B = w.batch                       # images per GPU (weak scaling: every rank has its own batch)
    anchors = wl.build_anchors(w)
    A, C = int(anchors.shape[0]), w.num_score_cols
    per_set = B * A * (C + 4) * 4
    nsets = max(2, -(-int(1.5 * L2_BYTES) // per_set))
    nsets = min(max(nsets, args.in_flight), 16)          # one input set per step in flight (HBM is not the limit)

    # synthetic inputs, one seed per (rank, set)
    host_sets = []
    for s in range(nsets):
        a_, gt, scores, locs = wl.make_inputs(w, seed=23 + 1000 * rank + s)
        host_sets.append((gt, scores.pin_memory(), locs.pin_memory()))
    anchors_dev = anchors.to(dev)
    dev_sets = []
    for gt, scores, locs in host_sets:
        packed = pack_ground_truth(gt, dev)
        packed.rows = packed.rows.clone()
        packed.offsets = packed.offsets.clone()
        dev_sets.append((packed, scores.to(dev), locs.to(dev)))
    torch.cuda.synchronize()

    # one pipeline (and one CUDA graph) per input set
    pipes, outs = [], []
    launches0 = N.lib().ssd_b200_launch_count()
    in_flight = max(1, min(args.in_flight, nsets))
    for k, (packed, scores_d, locs_d) in enumerate(dev_sets):
        # steps of different slots may run concurrently: each slot has its own scratch buffers
        pipe = AnchorPipeline(w.cfg(), workspace_slot=k % in_flight)
        pipes.append(pipe)
    # kernels per step, counted on one eager step
    launches_before = N.lib().ssd_b200_launch_count()
    pipes[0].step_device(dev_sets[0][0], anchors_dev, dev_sets[0][1], dev_sets[0][2])
    torch.cuda.synchronize()
    launches_per_step = int(N.lib().ssd_b200_launch_count() - launches_before)
    cap = B if world > 1 else None
    # The exchange step: by default the last kernel of every step graph packs the shard and writes it into every
    # rank's gathered buffer over NVLink peer memory (sharding.PeerExchange, csrc/exchange.cu); SSD_EXCHANGE=nccl
    # packs locally and calls NCCL's all-gather after every replay instead (capturing the NCCL collective into the
    # graph hung on this stack).
    px = None
    if world > 1 and os.environ.get("SSD_EXCHANGE", "peer") != "nccl":
        try:
            px = sharding.PeerExchange(B * world, w.max_total, slots=nsets)
            ok = torch.tensor([1], device=dev)
        except Exception as e:  # noqa: BLE001  (no peer access / IPC on this box: every rank must take the same route)
            print(f"[bench] peer exchange unavailable on rank {rank}: {e!r}", file=sys.stderr)
            px, ok = None, torch.tensor([0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok) == 0:
            px = None
    conc = in_flight > 1
    for k, (pipe, (packed, scores_d, locs_d)) in enumerate(zip(pipes, dev_sets)):
        if px is not None:
            outs.append(pipe.capture(packed, anchors_dev, scores_d, locs_d, exchange=(px, k), concurrent=conc))
        else:
            outs.append(pipe.capture(packed, anchors_dev, scores_d, locs_d, shard_capacity=cap, concurrent=conc))
    # the strictly serial number uses graphs captured for a step that runs alone (single GPU; with several ranks the
    # serial loop replays the in-flight graphs)
    pipes_serial = pipes
    if conc and world == 1:
        pipes_serial = []
        for packed, scores_d, locs_d in dev_sets:
            ps = AnchorPipeline(w.cfg())
            ps.capture(packed, anchors_dev, scores_d, locs_d)
            pipes_serial.append(ps)
    torch.cuda.synchronize()

    # the one exchange step of the path (detections + counts + stats, one collective per step) runs
    # asynchronously: the all-gather of step i overlaps the kernels of step i+1
    gather = sharding.OverlappedGather(B * world, w.max_total, ring=max(3, nsets + 1)) if world > 1 and px is None else None

    class _PeerFlush:                      # same surface as OverlappedGather for the loops below
        def flush(self):
            for k in range(nsets):
                px.wait(k)
            return px.gathered(0)

    if px is not None:
        gather = _PeerFlush()

    # Throughput mode: `in_flight` consecutive steps run concurrently, each replayed on its own stream (input
    # set k always on stream k % in_flight, with that slot's scratch buffers).  A step alone leaves most of the
    # GPU idle (its longest kernel, the NMS, runs 640 small CTAs at ~35 % issue utilisation), so the next
    # batch's streaming kernels fill the gaps.  --in-flight 1 is the strictly serial number (reported beside it).
    streams = [torch.cuda.Stream() for _ in range(in_flight)]

    def device_step(i, serial=False):
        k = i % nsets
        if serial or in_flight == 1:
            pipes_serial[k].replay()
            if world > 1 and px is None:
                return gather.submit(outs[k].shard)
            return outs[k].dets, outs[k].counts, outs[k].assign_stats
        with torch.cuda.stream(streams[k % in_flight]):
            pipes[k].replay()
            if world > 1 and px is None and not os.environ.get("SSD_BENCH_NO_GATHER"):   # (diagnostics)
                return gather.submit_nowait(outs[k].shard)
        return outs[k].dets, outs[k].counts, outs[k].assign_stats

    def fork():
        main = torch.cuda.current_stream()
        for s_ in streams:
            s_.wait_stream(main)

    def join():
        main = torch.cuda.current_stream()
        for s_ in streams:
            main.wait_stream(s_)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- serial reference: one step after the other on one stream ----
    for i in range(args.warmup):
        device_step(i, serial=True)
    if world > 1:
        gather.flush()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        device_step(i, serial=True)
    if world > 1:
        gather.flush()
    e1.record()
    barrier()
    serial_ms = e0.elapsed_time(e1)

    # ---- value: device-resident, CUDA events ----
    fork()
    for i in range(args.warmup):
        device_step(i)
    join()
    if world > 1:
        gather.flush()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fork()
    for i in range(args.steps):
        device_step(i)
    join()
    if world > 1:
        gather.flush()                    # the last step's exchange completes inside the timed region
    e1.record()
    barrier()
    dev_ms = e0.elapsed_time(e1)

    # ---- e2e: reference-shaped API, host buffers, H2D + D2H inside the timed region ----
    pipe_e = AnchorPipeline(w.cfg())
    T = w.max_total

    # AnchorPipeline.stream: the reference-shaped step over a sequence of host batches, the H2D copies
    # of batch i+1 overlapping the kernels of batch i; every batch's detections, counts and statistics
    # are read back to pinned host memory before it is yielded (under torchrun: after the all-gather).
    def e2e_run(n):
        seen = 0
        batches = (host_sets[i % nsets] for i in range(n))
        for target, mask, dets in pipe_e.stream(batches, anchors, gather_batch=B * world if world > 1 else None):
            seen += len(dets)
        assert seen == n * B * world
        torch.cuda.current_stream().synchronize()

    e2e_steps = max(3, min(args.steps, 50))
    e2e_run(max(3, min(args.warmup, 5)))
    barrier()
    t0 = time.perf_counter()
    e2e_run(e2e_steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    clock_info = clocks.stop() if rank == 0 else None

    gt0 = host_sets[0][0]
    gt_bytes = sum(g.numel() * 4 for g in gt0) + (B + 1) * 4
    h2d = B * A * C * 4 + B * A * 16 + gt_bytes
    d2h = B * world * (T * 6 * 4 + 5 * 4) + B * 4 + 16 + B * 16

    # ---- roofline: the dominant kernel alone, CUDA events, rotating (cold-L2) inputs ----
    roof = roofline_probe(w, dev_sets, anchors_dev, pipes[0], B, A, C, nsets)

    # ---- per-kernel times in context (library event timers around every launch of an eager step) ----
    import ctypes
    lib = N.lib()
    for i in range(3):
        k = i % nsets
        pipes[0].step_device(dev_sets[k][0], anchors_dev, dev_sets[k][1], dev_sets[k][2])
    torch.cuda.synchronize()
    lib.ssd_b200_timing_enable(1)
    for i in range(20):
        k = i % nsets
        pipes[0].step_device(dev_sets[k][0], anchors_dev, dev_sets[k][1], dev_sets[k][2])
    buf = ctypes.create_string_buffer(4096)
    lib.ssd_b200_timing_report(buf, 4096)
    lib.ssd_b200_timing_enable(0)
    kernels_us = {kv.split(":")[0]: float(kv.split(":")[1]) for kv in buf.value.decode().split(",") if kv}
    num_fg = C - (1 if w.converter == "SOFTMAX" else 0)
    pair_tests = B * num_fg * w.max_per_class * (w.max_per_class - 1) // 2
    nms_info = {"kernel": "segment_nms_kernel", "us_in_step": kernels_us.get("nms"),
                "pair_tests_per_launch": pair_tests,
                "pair_tests_per_s": pair_tests / (kernels_us["nms"] * 1e-6) if kernels_us.get("nms") else None,
                "note": "K(K-1)/2 IoU tests per (image, class) at K = max_per_class; bound by the ALU pipe, not HBM"}

    # max over ranks
    t = torch.tensor([dev_ms, e2e_s, serial_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_s, serial_ms = float(t[0]), float(t[1]), float(t[2])

    if rank == 0:
        ms_per_step = dev_ms / args.steps
        value = B * world / (ms_per_step * 1e-3)
        e2e_value = B * world * e2e_steps / e2e_s
        line = {
            "metric": "images/sec target-assign+NMS", "value": value, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w.name, "anchors": A, "score_cols": C, "images_per_gpu": B,
                       "global_batch": B * world, "converter": w.converter, "sampler": w.sampler,
                       "region": "encode_ground_truth + sampler + to_centroids/encode_box + postprocess"
                                 + (" + all_gather(dets,stats)" if world > 1 else ""),
                       "l2": f"inputs rotate over {nsets} sets = {nsets * per_set / 2**20:.0f} MiB > 126 MiB L2",
                       "steps_in_flight": in_flight,
                       "device_path": (f"CUDA graph replay per step, {in_flight} consecutive steps in flight on "
                                       f"{in_flight} streams (own scratch buffers per slot)" if in_flight > 1
                                       else "CUDA graph replay per step, one step after the other") + (
                           "" if world == 1 else (", shard packed and written into every rank's gathered buffer over NVLink "
                                                  "peer memory by the last kernel of the step graph" if px is not None
                                                  else ", NCCL all-gather of step i overlapping step i+1"))},
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
                    "path": "AnchorPipeline.stream(batches of (list of GT, pinned host scores, locs), CPU anchors) -> "
                            "(target, mask, list of host detections) per batch; H2D of batch i+1 overlaps batch i"},
            "serial": {"ms_per_step": serial_ms / args.steps, "value": B * world / (serial_ms / args.steps * 1e-3),
                       "note": "the same K steps strictly one after the other on one stream (--in-flight 1)"},
            "gpu_launches": launches_per_step * args.steps,
            "launches_per_step": launches_per_step,
            "clocks": clock_info,
            "roofline": roof,
            "kernels_us": kernels_us,
            "nms": nms_info,
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            images = min(args.cpu_sample_images, B)
            best, mean, total = time_cpu_oracle(w, images, 3, 1, cores)
            line["cpu_baseline"] = {
                "value": mean, "unit": "images/s", "cores": cores, "kind": "port", "best": best,
                "sample": f"{images} of {B} images per step, 3 steps + 1 warm-up, oracle port of the reference "
                          f"(torch CPU ops + torchvision.ops.nms), {cores} torch threads"}
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.cuda.synchronize()
        if px is not None:
            px.close()
        dist.barrier()
        dist.destroy_process_group()


def roofline_probe(w, dev_sets, anchors_dev, pipe, B, A, C, nsets, iters: int = 50):
    """Time the logit-streaming kernels alone (CUDA events on the launching stream, launches back to back,
    inputs rotating over sets larger than L2).  The roofline entry is the one that runs in the step: the
    post-processor's first pass (row statistics + gate bookkeeping + the sampler's criterion); the sampler's
    own streaming kernel (train steps, no post-processor) is reported beside it."""
    import ctypes
    from single_shot_detection_b200 import _native as N
    from single_shot_detection_b200 import ops
    peak, peak_src = measured_peaks()
    dev = anchors_dev.device
    stream = torch.cuda.current_stream().cuda_stream
    keys = torch.empty((B, A), dtype=torch.int32, device=dev)
    cls = [torch.zeros((B, A), dtype=torch.int64, device=dev) for _ in range(nsets)]
    lib = N.lib()

    def timed(launch):
        for i in range(5):
            launch(i % nsets)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            launch(i % nsets)
        e1.record()
        torch.cuda.synchronize()
        return 1e3 * e0.elapsed_time(e1) / iters

    ws = torch.empty((max(lib.ssd_hard_negative_workspace_bytes(B, A), 256),), dtype=torch.uint8, device=dev)

    def mining(k):
        N.check(lib.ssd_mining_keys(dev_sets[k][1].data_ptr(), cls[k].data_ptr(), B, A, C, keys.data_ptr(),
                                    ws.data_ptr(), ws.numel(), stream))

    algo_bytes = B * A * (4 * C + 8 + 4)          # logits + int64 class read, uint32 key written
    us_mining = timed(mining)
    other = {"mining_loss_kernel (ssd_mining_keys, the sampler alone)": {
        "us_per_launch": us_mining, "achieved": algo_bytes / (us_mining * 1e-6) / 1e9,
        "frac": algo_bytes / (us_mining * 1e-6) / 1e9 / peak, "algorithmic_bytes_per_launch": algo_bytes}}
    if w.converter != "SOFTMAX" or w.sampler != "hard_negative_mining":
        # no shared pass in this configuration: the sampler's kernel is not in the step either (naive sampler);
        # pass 1 without the criterion output is the step's streaming kernel
        want_keys = False
        algo_p1 = B * A * (4 * C)
    else:
        want_keys = True
        algo_p1 = B * A * (4 * C + 8 + 4)          # logits read, (max, sum) float2 + uint32 criterion key written
    post = pipe.postprocessor
    conv, first_fg = {"SOFTMAX": (N.CONVERT_SOFTMAX, 1), "SIGMOID": (N.CONVERT_SIGMOID, 0)}[w.converter]
    p = ops._post_params(dev_sets[0][1], dev_sets[0][2], conv, first_fg, N.BOXES_ENCODED, float(w.xy_scale),
                         float(w.wh_scale), float(w.score_threshold), int(w.max_per_class), float(w.overlap_threshold),
                         int(w.max_total or 0), 0.0)
    pws = ops._post_workspace(p, dev)

    def pass1(k):
        N.check(lib.ssd_postprocess_pass1(ctypes.byref(p), dev_sets[k][1].data_ptr(), keys.data_ptr() if want_keys else None,
                                          pws.data_ptr(), pws.numel(), stream))

    us = timed(pass1)
    achieved = algo_p1 / (us * 1e-6) / 1e9
    # dram__bytes_read.sum + dram__bytes_write.sum of this kernel, one `ncu --set full` capture per workload
    # (profiles/r01c_ncu_full.md); None for a workload that has no committed capture
    traffic = NCU_TRAFFIC.get(w.name)
    return {"bound": "hbm", "kernel": "score_pass1_kernel (ssd_postprocess_pass1)", "achieved": achieved, "peak": peak,
            "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "traffic_source": "profiles/r01c_ncu_full.md" if traffic else None,
            "algorithmic_bytes_per_launch": algo_p1, "us_per_launch": us, "other_streaming_kernels": other,
            "note": "the streaming kernel of the step: one read of the logits yields the row statistics, the gate "
                    "bookkeeping and the sampler's criterion (4C read + 12 written bytes per anchor); the call is timed "
                    "back to back on one stream INCLUDING its scratch-zeroing kernel, inputs rotating over sets larger "
                    "than L2; the largest kernel of the step, segment_nms_kernel, is ALU bound and reported under 'nms'"}


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the roofline kernel (profiles/r01c_ncu_full.md)
NCU_TRAFFIC = {"ssd300_voc_b32": 23524608, "ssd512_coco_b32": 254768384 + 15038208}


def main():
    args = parse_args()
    # stdout carries exactly one JSON line: everything libraries print there (NCCL's version banner)
    # goes to stderr instead
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import builtins
    original_print = builtins.print

    def emit(*a, **k):
        if not k.get("file"):
            os.write(real_stdout, (" ".join(str(x) for x in a) + "\n").encode())      # straight to the real stdout
        else:
            original_print(*a, **k)

    builtins.print = emit
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_ours(args)
    finally:
        builtins.print = original_print
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)


if __name__ == "__main__":
    main()
