# Round-end evidence in one call: every GPU test, the default bench line, configs, ncu captures.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_gpu.log
tail -3 gpurun_out/tests_gpu.log
timeout 400 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_full.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','constraints_per_sec','gpu_launches','clocks')}, d['e2e'], d['cpu_baseline'])
PY
timeout 600 python benchmarks/run_configs.py --configs 1,3,4,5 > gpurun_out/configs_latest.jsonl 2> gpurun_out/configs_latest.err
python - <<PY
import json
for ln in open('gpurun_out/configs_latest.jsonl'):
    if ln.startswith('{'):
        d=json.loads(ln)
        print(d.get('config'), {k:(round(v,2) if isinstance(v,float) else v) for k,v in d.items() if k.startswith('gpu_') and not isinstance(v,dict)}, d.get('parity_ok'))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_score_top_tile -c 1 -o gpurun_out/prof_r1f_top -f python bench.py --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/ncu_top.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_expand_lattice -s 5 -c 1 -o gpurun_out/prof_r1f_lattice -f python bench.py --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/ncu_lat.log 2>&1
ls gpurun_out | head -40
