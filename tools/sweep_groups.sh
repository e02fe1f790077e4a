#!/bin/bash
# Steps in flight x steps per graph launch (bench.py --in-flight F --group G): profiles/r02_tuning_log.md section 2.
for cfg in "8 1" "8 2" "8 4" "8 8" "16 4" "16 8" "12 4"; do
  set -- $cfg
  python bench.py --steps 200 --warmup 20 --no-config5 --no-cpu-baseline --no-e2e --in-flight $1 --group $2 \
      > gpurun_out/sweep_if$1_g$2.json 2>> gpurun_out/sweep.err
done
