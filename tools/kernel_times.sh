#!/bin/bash
# Per-kernel device times of one eager step (ncu, cold cache, serialised: compare SHARES).
# usage: tools/kernel_times.sh <workload> <out-prefix>
W=${1:-ssd300_voc_b32}; OUT=${2:-gpurun_out/kt_$W}
python scratch/prof_step.py $W 4 > $OUT.plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:_kernel -c 80 --csv --log-file $OUT.csv python scratch/prof_step.py $W 4 > $OUT.ncu.log 2>&1
python tools/parse_launches.py $OUT.csv
