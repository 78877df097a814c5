bash tools/gpu/manifest.sh check || exit 9
mkdir -p gpurun_out
timeout 70 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_last.json 2> gpurun_out/bench_r2_last.err; tail -c 300 gpurun_out/bench_r2_last.err; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r2_last.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','parity_checked','parity_failed')}, d['e2e']['value'], d['roofline'].get('frac'), d['clocks'])
PY
