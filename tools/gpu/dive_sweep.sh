mkdir -p gpurun_out
for r in 0.90 0.94 0.97 0.99; do
  CSM_DIVE_RATIO=$r timeout 300 python bench.py --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/dive_$r.json 2> gpurun_out/dive_$r.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/dive_$r.json').read().strip().splitlines()[-1])
print('$r', round(d['ms_per_step'],3), round(d['value']/1e9,3), round(d['constraints_per_sec'],1), {k:(v['ms'],int(v['candidates'])) for k,v in d['roofline']['kernels'].items() if k in ('k_dive','k_expand_lattice')})
PY
done
