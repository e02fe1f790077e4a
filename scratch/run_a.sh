python -m pytest tests -m gpu -q -x 2>&1 | tail -3 > gpurun_out/r2_t5.log
python bench.py --steps 200 --warmup 20 --no-config5 --no-cpu-baseline > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err
python tools/graph_timeline.py ssd300_voc_b32 > gpurun_out/r2_timeline4.log 2>&1
SSD_ASSIGN_AFTER_PASS1=0 python bench.py --steps 200 --warmup 20 --no-config5 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench4b.json 2>> gpurun_out/r2_bench4.err
