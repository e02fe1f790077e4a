#!/bin/bash
# Bench line + per-kernel ncu metrics (+ the table of an `ncu --set full` capture) of every BASELINE configuration
# on one GPU.  usage: tools/profile_all.sh <tag>  -> gpurun_out/<tag>_bench_<workload>.json, <tag>_ncu_<workload>.*
TAG=${1:-r02}
for W in ssd300_voc_b32 ssd_mb2_coco_b64 ssd512_coco_b32 retina500_coco_b32 m2det512_coco_b256; do
  STEPS=100; [ $W = m2det512_coco_b256 ] && STEPS=20
  python bench.py --workload $W --steps $STEPS --warmup 10 --no-config5 --no-cpu-baseline --no-e2e \
      > gpurun_out/${TAG}_bench_$W.json 2> gpurun_out/${TAG}_bench_$W.err
  KEEP=; [ $W = ssd300_voc_b32 ] && KEEP=keep
  timeout 600 bash tools/ncu_step.sh $W gpurun_out/${TAG}_ncu_$W full $KEEP
done
du -sh gpurun_out
