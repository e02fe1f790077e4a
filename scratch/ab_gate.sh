run() { echo -n "$* : "; env "$@" python bench.py --steps 300 --warmup 30 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['steps_in_flight'], round(d['value']), round(1e3*d['ms_per_step'],2), 'serial', round(1e3*d['serial']['ms_per_step'],2), d['launches_per_step'])"; }
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
run SSD_GATE_KERNEL=1
run SSD_GATE_KERNEL=0
run SSD_GATE_KERNEL=1
run SSD_GATE_KERNEL=0
