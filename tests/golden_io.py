"""Helpers to read tests/golden/*.npz (written by tests/golden/make_golden.py)."""
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

PIPELINE_CASES = ["tiny_voc_b3", "tiny_sigmoid_b2", "edge_tiny", "mixup_ignoreband_tiny",
                  "ssd_mb2_coco_b64", "ssd300_voc_b8"]
SMALL_CASES = ["tiny_voc_b3", "tiny_sigmoid_b2", "edge_tiny", "mixup_ignoreband_tiny"]


def load(name):
    return np.load(os.path.join(GOLDEN_DIR, name), allow_pickle=False)


def split_ragged(flat, off):
    flat = torch.from_numpy(np.ascontiguousarray(flat))
    return [flat[int(off[i]):int(off[i + 1])].clone() for i in range(len(off) - 1)]


class PipelineCase:
    def __init__(self, name):
        from single_shot_detection_b200 import workloads as wl
        z = load(f"pipeline_{name}.npz")
        self.name = name
        self.z = z
        self.w = wl.WORKLOADS[str(z["workload"])]
        self.matched, self.unmatched = (float(x) for x in z["thresholds"])
        self.anchors = torch.from_numpy(z["anchors"])
        self.gt = split_ragged(z["gt_flat"], z["gt_off"])
        self.scores = torch.from_numpy(z["scores"])
        self.locs = torch.from_numpy(z["locs"])
        self.B = len(self.gt)
        self.A = self.anchors.shape[0]
        self.C = self.scores.shape[1] // self.A
        self.target = torch.from_numpy(z["target"])
        self.match_idx = torch.from_numpy(z["match_idx"])
        self.hnm_mask = torch.from_numpy(np.unpackbits(z["hnm_mask"], axis=1)[:, :self.A].astype(bool))
        self.naive_mask = torch.from_numpy(np.unpackbits(z["naive_mask"], axis=1)[:, :self.A].astype(bool))
        self.enc_inplace = torch.from_numpy(z["enc_inplace"])
        self.dets = split_ragged(z["det_flat"], z["det_off"])
        self.dets_all = split_ragged(z["det_all_flat"], z["det_all_off"])
        self.loss3 = z["loss3"]
        self.full = "probs" in z.files

    def t(self, key):
        return torch.from_numpy(self.z[key])

    def cfg(self):
        c = self.w.cfg()
        c["matched_threshold"] = self.matched
        c["unmatched_threshold"] = self.unmatched
        return c
