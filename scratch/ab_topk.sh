for rep in 1 2; do
for f in separate fused; do
  echo -n "SSD_TOPK=$f: "
  SSD_TOPK=$f python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(1e3*d['ms_per_step'],2), d['launches_per_step'])"
done
done
