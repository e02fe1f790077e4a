import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from oracle import anchor_pipeline_oracle as ora
from single_shot_detection_b200 import workloads as wl
from single_shot_detection_b200 import target_assigner as ta
import golden_io as gio

w = wl.WORKLOADS["ssd300_voc_b8"]
anchors, gt, scores, locs = wl.make_inputs(w, seed=23, batch=2)
case = gio.PipelineCase("ssd300_voc_b8")
print("gt equal to golden:", all(torch.equal(a, b) for a, b in zip(gt, case.gt)), "scores equal:", torch.equal(scores, case.scores))
for trial in range(3):
    assigner = ta.TargetAssigner(0.5, 0.5)
    target = assigner.encode_ground_truth(gt, anchors)
    ref, ref_match = ora.assign_targets(gt, anchors, 0.5, 0.5, return_match=True)
    m = assigner.last_match.cpu().long()
    r = torch.stack(ref_match)
    bad = (m != r).nonzero()
    print("trial", trial, "mismatches", bad.shape[0], bad[:10].tolist(), [(int(m[i, j]), int(r[i, j])) for i, j in bad[:10].tolist()])
    print(" stats", assigner.last_stats.cpu().tolist())

# sigmoid top-k case
from single_shot_detection_b200 import box_coder, postprocessor
case = gio.PipelineCase("tiny_sigmoid_b2")
w = case.w
coder = box_coder.BoxCoder(w.xy_scale, w.wh_scale, w.eps)
for mt, ref in ((w.max_total, case.dets), (None, case.dets_all)):
    post = postprocessor.Postprocessor(coder, w.score_threshold, {"max_per_class": w.max_per_class, "overlap_threshold": w.overlap_threshold}, score_converter=w.converter, max_total=mt)
    dets = post.postprocess((case.scores.cuda(), case.locs.cuda()), case.anchors)
    for i, (d, r) in enumerate(zip(dets, ref)):
        d = d.cpu()
        print("mt", mt, "img", i, d.shape, r.shape)
        if d.shape == r.shape:
            diff = (d - r).abs().max(dim=1)[0]
            badrows = (diff > 1e-4).nonzero().view(-1)
            print("  bad rows", badrows.tolist()[:20])
            for k in badrows.tolist()[:5]:
                print("   mine", d[k].tolist(), "\n   ref ", r[k].tolist())
        else:
            print("  classes mine", torch.bincount(d[:,4].long()).tolist(), "ref", torch.bincount(r[:,4].long()).tolist())
