# bench variants under tuning knobs (diagnostics): results in gpurun_out/r2b_env_<tag>.json
run() { tag=$1; shift; env "$@" python bench.py --steps 200 --warmup 20 --no-config5 --no-cpu-baseline --no-e2e > gpurun_out/r2b_env_$tag.json 2>> gpurun_out/r2b_env.err; }
run c1 SSD_CONCURRENT_CTAS=1
run c2 SSD_CONCURRENT_CTAS=2
run c3 SSD_CONCURRENT_CTAS=3
