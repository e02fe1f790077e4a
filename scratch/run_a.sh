for f in 2 3 6 8 12 16; do
  python bench.py --steps 200 --warmup 20 --no-config5 --no-cpu-baseline --no-e2e --in-flight $f > gpurun_out/r2_bench_if$f.json 2>> gpurun_out/r2_if.err
done
