for p in 0 1; do
  echo "SSD_GRAPH_PRIORITY=$p"
  SSD_GRAPH_PRIORITY=$p python tools/graph_timeline.py ssd300_voc_b32 3 2>&1 | tail -3 | cut -c1-420
  SSD_GRAPH_PRIORITY=$p python bench.py --steps 100 --warmup 10 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"
done
