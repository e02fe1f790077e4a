#!/bin/bash
# Per-kernel device time, executed warp instructions and DRAM bytes of one eager step (ncu: cold cache,
# serialised launches -- compare SHARES, not absolute times); `full` adds an `ncu --set full` capture of the
# last step: its raw page as CSV + the per-kernel table (tools/ncu_summary.py); `keep` also keeps the report
# itself (~25 MB -- gpurun_out/ travels back only below 64 MiB) for tools/ncu_lines.py / ncu_traffic.py.
# usage: tools/ncu_step.sh <workload> <out-prefix> [full [keep]]
W=${1:-ssd300_voc_b32}; OUT=${2:-gpurun_out/ncu_$W}
python tools/prof_step.py $W 4 > $OUT.plain.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -k regex:_kernel -c 60 --csv --log-file $OUT.launches.csv python tools/prof_step.py $W 4 > $OUT.ncu.log 2>&1
python tools/parse_step_metrics.py $OUT.launches.csv > $OUT.launches.md 2>&1
if [ "$3" = full ]; then
    REP=/tmp/$(basename $OUT)
    [ "$4" = keep ] && REP=$OUT
    ncu --set full --import-source on --clock-control none -k regex:_kernel --launch-skip 16 -c 8 -f -o $REP \
        python tools/prof_step.py $W 3 > $OUT.ncufull.log 2>&1
    python tools/ncu_summary.py $REP.ncu-rep > $OUT.full.md 2>&1
fi
