bash tools/gpu/manifest.sh check || exit 9
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/tests_gpu.log 2>&1; tail -6 gpurun_out/tests_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2h.json 2> gpurun_out/bench_r2h.err; tail -c 600 gpurun_out/bench_r2h.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2h_2.json 2>> gpurun_out/bench_r2h.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_expand_lattice|k_score_top_tile' -c 8 -o gpurun_out/r2h_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python - <<PY
import json
for f in ('gpurun_out/bench_r2h.json','gpurun_out/bench_r2h_2.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f,{k:d.get(k) for k in ('value','ms_per_step','device_ms_per_step','constraints_per_sec','gpu_launches','host_syncs_per_batch','parity_checked','parity_failed')}, d['e2e'])
        print(d['step_ms_per_rank'])
        print({k:d['roofline'].get(k) for k in ('kernel','bound','frac','ncu_pct_of_peak','algorithmic_GBps','hbm')}, {k:v['ms'] for k,v in d['roofline']['kernels'].items()})
    except Exception as e:
        print(f,'ERR',e)
PY
