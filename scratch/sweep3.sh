run() { echo -n "$* : "; env "$@" python bench.py --steps 400 --warmup 40 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['steps_in_flight'], round(d['value']), round(1e3*d['ms_per_step'],2), 'serial', round(1e3*d['serial']['ms_per_step'],2))"; }
run A=1
run SSD_GRAPH_PRIORITY=0
run SSD_TOPK=fused
run SSD_GATE_KERNEL=0
run A=1
