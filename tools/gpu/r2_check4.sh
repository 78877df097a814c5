bash tools/gpu/manifest.sh check || exit 9
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rt2d.py tests/test_gpu_parity_2d.py -q -k "rt" > gpurun_out/rt_v2_tests.log 2>&1; tail -4 gpurun_out/rt_v2_tests.log
timeout 300 python bench.py --config 1 --steps 5 --warmup 3 > gpurun_out/bench_r2c_rt.json 2> gpurun_out/bench_r2c.err; tail -c 600 gpurun_out/bench_r2c.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_rt_match' -c 1 -o gpurun_out/r2c_rt python bench.py --config 1 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_rt.log 2>&1; tail -2 gpurun_out/ncu_rt.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2c.json 2>> gpurun_out/bench_r2c.err
timeout 900 python bench.py --config 3 --steps 2 --warmup 1 > gpurun_out/bench_r2c_c3.json 2>> gpurun_out/bench_r2c.err
timeout 900 python bench.py --config 5 --scale 0.04 --steps 2 --warmup 1 > gpurun_out/bench_r2c_c5_s004.json 2>> gpurun_out/bench_r2c.err
tail -c 1500 gpurun_out/bench_r2c.err
python - <<PY
import json
for f in ('gpurun_out/bench_r2c_rt.json','gpurun_out/bench_r2c.json','gpurun_out/bench_r2c_c3.json','gpurun_out/bench_r2c_c5_s004.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f,{k:d.get(k) for k in ('value','ms_per_step','device_ms_per_step','constraints_per_sec','matches_per_sec','single_call_ms','host_syncs_per_batch','parity_checked','parity_failed')}, d['e2e'].get('value'), d.get('cpu_baseline'))
        if d.get('roofline'): print({k:d['roofline'].get(k) for k in ('kernel','bound','frac','achieved','ncu_pct_of_peak','algorithmic_GBps')}, {k:v['ms'] for k,v in d['roofline']['kernels'].items()})
        print(d['config'])
    except Exception as e:
        print(f,'ERR',e)
PY
ls -la gpurun_out | head -30
