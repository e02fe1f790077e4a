import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from single_shot_detection_b200 import workloads as wl
from single_shot_detection_b200.pipeline import AnchorPipeline
from single_shot_detection_b200.target_assigner import pack_ground_truth
name = sys.argv[1] if len(sys.argv) > 1 else "ssd300_voc_b32"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
w = wl.WORKLOADS[name]
dev = torch.device("cuda", 0)
anchors, gt, scores, locs = wl.make_inputs(w, seed=23)
anchors_d, scores_d, locs_d = anchors.to(dev), scores.to(dev), locs.to(dev)
packed = pack_ground_truth(gt, dev)
pipe = AnchorPipeline(w.cfg())
for i in range(steps):
    out = pipe.step_device(packed, anchors_d, scores_d, locs_d)
torch.cuda.synchronize()
print("ok", out.counts.tolist()[:4], out.status.tolist())
