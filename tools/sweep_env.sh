# bench variants under tuning knobs (diagnostics): results in gpurun_out/r2c_env_<tag>.json
run() { tag=$1; shift; env "$@" python bench.py --steps 200 --warmup 20 --no-config5 --no-cpu-baseline --no-e2e > gpurun_out/r2c_env_$tag.json 2>> gpurun_out/r2c_env.err; }
run chain1 SSD_SERIAL_CHAIN=1
run chain0 SSD_SERIAL_CHAIN=0
run chain1p SSD_SERIAL_CHAIN=1 SSD_PASS1_FIRST=1
