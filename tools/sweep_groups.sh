for cfg in "8 1" "8 2" "8 4" "8 8" "16 4" "16 8" "12 4"; do
  set -- $cfg
  python bench.py --steps 200 --warmup 20 --no-config5 --no-cpu-baseline --no-e2e --in-flight $1 --group $2 > gpurun_out/r2b_g_$1_$2.json 2>> gpurun_out/r2b_g.err
done
