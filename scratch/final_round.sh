python bench.py --steps 200 --warmup 20 2>gpurun_out/bench9.err > gpurun_out/bench9.json; echo "bench rc=$?"
python tools/graph_timeline.py ssd300_voc_b32 3 2>&1 | tail -3 > gpurun_out/timeline9.txt; cat gpurun_out/timeline9.txt | cut -c1-400
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null > gpurun_out/bench9_ref.json; cat gpurun_out/bench9_ref.json | cut -c1-300
