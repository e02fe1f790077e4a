"""Helpers to read tests/golden/*.npz (written by tests/golden/make_golden.py)."""
import hashlib
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# BASELINE.json configs at full anchor / class shapes, B = 2 (make_golden.py: run_config_case): inputs regenerated
# from the seed and pinned by SHA-256, compact reference outputs stored
CONFIG_CASES = ["ssd300_voc_b32", "ssd512_coco_b32", "retina500_coco_b32", "m2det512_coco_b256"]
PIPELINE_CASES = ["tiny_voc_b3", "tiny_sigmoid_b2", "edge_tiny", "mixup_ignoreband_tiny",
                  "ssd_mb2_coco_b64", "ssd300_voc_b8"] + CONFIG_CASES
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
        self.gt = split_ragged(z["gt_flat"], z["gt_off"])
        self.regenerated = "regen_seed" in z.files
        if self.regenerated:
            # inputs come from the recipe that made the fixture; the digests prove they are the same bits
            anchors, gt, scores, locs = wl.make_inputs(self.w, seed=int(z["regen_seed"]), batch=int(z["regen_batch"]))
            for key, t in (("sha_scores", scores), ("sha_locs", locs), ("sha_anchors", anchors)):
                digest = hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()
                assert digest == str(z[key]), f"{name}: regenerated {key[4:]} differ from the ones the fixture was made with"
            assert all(torch.equal(a, b) for a, b in zip(gt, self.gt)), f"{name}: regenerated ground truth differs"
            self.anchors, self.scores, self.locs = anchors, scores, locs
        else:
            self.anchors = torch.from_numpy(z["anchors"])
            self.scores = torch.from_numpy(z["scores"])
            self.locs = torch.from_numpy(z["locs"])
        self.B = len(self.gt)
        self.A = self.anchors.shape[0]
        self.C = self.scores.shape[1] // self.A
        self.target = torch.from_numpy(z["target"])
        self.match_idx = torch.from_numpy(z["match_idx"].astype(np.int64))
        self.hnm_mask = torch.from_numpy(np.unpackbits(z["hnm_mask"], axis=1)[:, :self.A].astype(bool))
        self.naive_mask = torch.from_numpy(np.unpackbits(z["naive_mask"], axis=1)[:, :self.A].astype(bool))
        if "enc_rows" in z.files:       # coded boxes stored for the matched rows + a sample of the others
            self.enc_rows = torch.from_numpy(z["enc_rows"].astype(np.int64))
            self.enc_inplace = torch.from_numpy(z["enc_vals"])
        else:
            self.enc_rows = None
            self.enc_inplace = torch.from_numpy(z["enc_inplace"])
        self.det_all_anchor = (split_ragged(z["det_all_anchor"], z["det_all_off"])
                               if "det_all_anchor" in z.files else None)
        self.dets = split_ragged(z["det_flat"], z["det_off"])
        self.dets_all = split_ragged(z["det_all_flat"], z["det_all_off"])
        self.loss3 = z["loss3"]
        self.full = "probs" in z.files

    def enc_view(self, boxes):
        """The rows of a [B, A, 4] coded-box tensor that ``enc_inplace`` holds reference values for."""
        return boxes if self.enc_rows is None else boxes.reshape(-1, 4)[self.enc_rows.to(boxes.device)]

    def t(self, key):
        return torch.from_numpy(self.z[key])

    def cfg(self):
        c = self.w.cfg()
        c["matched_threshold"] = self.matched
        c["unmatched_threshold"] = self.unmatched
        return c
