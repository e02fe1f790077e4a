#!/bin/bash
# round profile: plain bench, launch list, ncu --set full of one eager step (both after a plain run exited 0)
set -x
python bench.py --steps 50 --warmup 5 2>gpurun_out/bench8.err > gpurun_out/bench8.json; echo "bench rc=$?"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r01c_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r01c_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r01c_ncu_bench.log 2>&1
python scratch/prof_step.py ssd300_voc_b32 3 > gpurun_out/r01c_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:assign|mining|box_transform|score_pass|class_gate|segment_nms|image_topk|zero_kernel" -s 20 -c 10 \
    -o gpurun_out/r01c_full_ssd300 -f python scratch/prof_step.py ssd300_voc_b32 3 > gpurun_out/r01c_ncu_full.log 2>&1
ls -la gpurun_out/r01b*
