python tools/fused_probe.py ssd300_voc_b32 2>&1 | grep "^-1" > gpurun_out/r2_fprobe.log
python tools/fused_probe.py ssd300_voc_b32 2>&1 | grep "^-1" >> gpurun_out/r2_fprobe.log
