for n in 1 2 5 40 300; do
python - <<PY 2>&1 | tail -1
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from single_shot_detection_b200 import box_utils
dev = torch.device("cuda", 0)
gen = torch.Generator().manual_seed(1)
n = $n
c = torch.rand((n, 2), generator=gen) * 200
s = torch.rand((n, 2), generator=gen) * 70 + 10
boxes = torch.cat([c - s / 2, c + s / 2], 1).to(dev)
scores = torch.rand((n,), generator=gen).to(dev)
try:
    (bk, sk), keep = box_utils.nms(boxes, scores, .45, .01)
    torch.cuda.synchronize()
    print(n, "ok", keep.numel())
except Exception as e:
    print(n, "FAIL", str(e)[:90])
PY
done
python scratch/prof_step.py ssd300_voc_b32 2 2>&1 | tail -1 | cut -c1-150
