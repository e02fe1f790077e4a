# bench variants under tuning knobs (diagnostics): results in gpurun_out/r2d_env_<tag>.json
run() { tag=$1; shift; env "$@" python bench.py --steps 100 --warmup 10 --no-config5 --no-cpu-baseline --no-e2e > gpurun_out/r2d_env_$tag.json 2>> gpurun_out/r2d_env.err; }
run nt128 SSD_NMS_THREADS=128
run nt64 SSD_NMS_THREADS=64
run nt32 SSD_NMS_THREADS=32
run nt256 SSD_NMS_THREADS=256
