#!/bin/bash
# ncu launch list of the BENCH command itself (the --metrics gpu__time_duration.sum --clock-control none pass of
# B200_PROFILING.md): cold-cache, serialised per-launch times -- each kernel's SHARE of the step must agree with the
# bench line's kernels_us, not the absolute.  usage: tools/bench_launch_list.sh <out-prefix>
OUT=${1:-gpurun_out/bench_launches}
CMD="python bench.py --steps 2 --warmup 1 --no-config5 --no-cpu-baseline --no-e2e"
$CMD > $OUT.plain.json 2> $OUT.plain.err || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:_kernel -c 2000 --csv --log-file $OUT.csv $CMD > $OUT.ncu.log 2>&1
python tools/parse_launches.py $OUT.csv > $OUT.md 2>&1
