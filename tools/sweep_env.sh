#!/bin/bash
# A/B of tuning knobs through the bench line (diagnostics, one B200): every argument is "tag VAR=VALUE [VAR=VALUE ...]";
# results in gpurun_out/sweep_<tag>.json.  Example (profiles/r02_tuning_log.md section 3):
#   tools/sweep_env.sh "base X=1" "late SSD_ASSIGN_AFTER_PASS1=1" "topkf SSD_TOPK=f" "nt256 SSD_NMS_THREADS=256"
for spec in "$@"; do
  set -- $spec; tag=$1; shift
  env "$@" python bench.py --steps 100 --warmup 10 --no-config5 --no-cpu-baseline --no-e2e \
      > gpurun_out/sweep_$tag.json 2>> gpurun_out/sweep.err
done
