python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for w in ssd_mb2_coco_b64 retina500_coco_b32 m2det512_coco_b256 ssd512_coco_b32; do
  python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench_$w.err > gpurun_out/bench_$w.json; echo "$w rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_$w.json')); r=d['roofline']
print(round(d['value']), round(1e3*d['ms_per_step'],1), 'e2e', round(d['e2e']['value']), 'launches', d['launches_per_step'], 'roof', r['kernel'][:20], round(r['frac'],3), round(r['us_per_launch'],1), {k: round(v['frac'],3) for k,v in r['other_streaming_kernels'].items()})
print(d['kernels_us'])
" 2>&1 | tail -3
done
