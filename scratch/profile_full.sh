#!/bin/bash
python scratch/prof_step.py ssd300_voc_b32 3 > gpurun_out/r01c_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:assign|mining|box_transform|score_pass|class_gate|segment_nms|image_topk|zero_kernel" -s 7 -c 7 \
    -o gpurun_out/r01c_full_ssd300 -f python scratch/prof_step.py ssd300_voc_b32 3 > gpurun_out/r01c_ncu_full.log 2>&1
python scratch/prof_step.py ssd512_coco_b32 3 > gpurun_out/r01c_plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:assign|mining|box_transform|score_pass|class_gate|segment_nms|image_topk|zero_kernel" -s 8 -c 8 \
    -o gpurun_out/r01c_full_ssd512 -f python scratch/prof_step.py ssd512_coco_b32 3 > gpurun_out/r01c_ncu_full512.log 2>&1
ls -la gpurun_out/r01c_full*
