run() { echo "== $*"; env "$@" timeout 120 python tools/graph_timeline.py ssd512_coco_b32 1 2>&1 | grep -v "^frame" | tail -2 | cut -c1-300; }
echo "== eager multi-stream"; timeout 120 python scratch/prof_step.py ssd512_coco_b32 3 2>&1 | tail -1 | cut -c1-200
run A=1
run SSD_SHARE_PASS=0
run SSD_GRAPH_PRIORITY=0
run SSD_SHARE_PASS=0 SSD_GRAPH_PRIORITY=0
