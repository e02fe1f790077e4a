#!/usr/bin/env python
"""The candidate-selection launch(es) of the post-processor alone: per-launch time (CUDA events, cold-L2 rotating
inputs) of the fused cluster kernel for every cluster size against the streaming pass 1 -> gates -> pass 2 chain, and
the phase marks of one fused CTA (%globaltimer).  `python tools/fused_probe.py [workload] [batch]`."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from single_shot_detection_b200 import _native as N  # noqa: E402
from single_shot_detection_b200 import ops, workloads as wl  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else wl.HEADLINE
    w = wl.WORKLOADS[name]
    B = int(sys.argv[2]) if len(sys.argv) > 2 else w.batch
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    anchors = wl.build_anchors(w).to(dev)
    A, C = int(anchors.shape[0]), w.num_score_cols
    nsets = min(16, max(2, -(-int(1.5 * 126 * 2**20) // (B * A * (C + 4) * 4))))
    sets = []
    for s in range(nsets):
        _, gt, scores, locs = wl.make_inputs(w, seed=23 + s, batch=B)
        sets.append((scores.to(dev), locs.to(dev)))
    lib = N.lib()
    conv, first_fg = {"SOFTMAX": (N.CONVERT_SOFTMAX, 1), "SIGMOID": (N.CONVERT_SIGMOID, 0)}[w.converter]
    stream = torch.cuda.current_stream().cuda_stream
    keys = torch.empty((B, A), dtype=torch.int32, device=dev)
    out = {}
    for mode in (0, -1, 1, 2, 4, 8):
        N.check(lib.ssd_b200_set_fused_select(mode))
        p = ops._post_params(sets[0][0], sets[0][1], conv, first_fg, N.BOXES_ENCODED, float(w.xy_scale), float(w.wh_scale),
                             float(w.score_threshold), int(w.max_per_class), float(w.overlap_threshold), int(w.max_total or 0), 0.0)
        ws = ops.workspace(lib.ssd_postprocess_workspace_bytes(ctypes.byref(p)), dev, f"probe{mode}")
        want_keys = w.converter == "SOFTMAX"

        def launch(k, full=False):
            N.check(lib.ssd_postprocess_pass1(ctypes.byref(p), sets[k][0].data_ptr(), keys.data_ptr() if want_keys else None,
                                              ws.data_ptr(), ws.numel(), stream))

        for i in range(5):
            launch(i % nsets)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 100
        e0.record()
        for i in range(iters):
            launch(i % nsets)
        e1.record()
        torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / iters
        # phase marks of the middle CTA
        nslots = lib.ssd_b200_trace_slots()
        buf = torch.empty((nslots, 2), dtype=torch.int64, device=dev)
        buf[:, 0] = torch.iinfo(torch.int64).max
        buf[:, 1] = 0
        torch.cuda.synchronize()
        assert lib.ssd_b200_trace_enable(buf.data_ptr()) == 0
        launch(1 % nsets)
        torch.cuda.synchronize()
        lib.ssd_b200_trace_enable(None)
        t = buf.cpu()
        used = [(int(t[i, 0]), int(t[i, 1])) for i in range(16) if int(t[i, 1]) > 0]
        t0 = min(u[0] for u in used)
        marks = [round((int(t[i, 0]) - t0) / 1e3, 2) for i in range(16, nslots) if int(t[i, 1]) == 1]
        out[str(mode)] = {"us_per_call": round(us, 2), "span_us": round((max(u[1] for u in used) - t0) / 1e3, 2),
                          "marks_us": marks}
        print(mode, out[str(mode)], flush=True)
    N.check(lib.ssd_b200_set_fused_select(-1))
    print(json.dumps({name: out}))


if __name__ == "__main__":
    main()
