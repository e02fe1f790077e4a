import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from single_shot_detection_b200 import box_utils, _native as N
dev = torch.device("cuda", 0)
lib = N.lib()
lib.ssd_b200_timing_enable(1)
gen = torch.Generator().manual_seed(1)
for n, soft in [(1, False), (1, True), (40, False), (40, True), (300, False)]:
    c = torch.rand((n, 2), generator=gen) * 200
    s = torch.rand((n, 2), generator=gen) * 70 + 10
    boxes = torch.cat([c - s / 2, c + s / 2], 1).to(dev)
    scores = torch.rand((n,), generator=gen).to(dev)
    try:
        (bk, sk), keep = box_utils.nms(boxes, scores, .45, .01, soft=soft)
        torch.cuda.synchronize()
        print(n, soft, "ok", keep.numel())
    except Exception as e:
        print(n, soft, "FAIL", str(e)[:100])
        break
