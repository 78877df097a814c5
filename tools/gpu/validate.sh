# Full GPU validation: every -m gpu test, the default bench line, configs 3/4/5.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_gpu.log
tail -4 gpurun_out/tests_gpu.log
timeout 400 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 600 gpurun_out/bench_full.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_full.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','constraints_per_sec','gpu_launches','clocks')}, d['e2e'], d['cpu_baseline'])
PY
timeout 900 python benchmarks/run_configs.py --configs 1,3,4,5 > gpurun_out/configs_latest.jsonl 2> gpurun_out/configs_latest.err
python - <<PY
import json
for ln in open('gpurun_out/configs_latest.jsonl'):
    ln=ln.strip()
    if not ln.startswith('{'): continue
    d=json.loads(ln)
    print(d.get('config'), {k:(round(v,2) if isinstance(v,float) else v) for k,v in d.items() if k.startswith('gpu_') and not isinstance(v,dict)}, d.get('parity_ok'))
PY
