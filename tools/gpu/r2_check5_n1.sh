bash tools/gpu/manifest.sh check || exit 9
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err
timeout 1200 python bench.py --config 4 --steps 2 --warmup 1 > gpurun_out/bench_r2d_c4_n1.json 2>> gpurun_out/bench_r2d.err
timeout 900 python bench.py --config 3 --steps 2 --warmup 1 > gpurun_out/bench_r2d_c3.json 2>> gpurun_out/bench_r2d.err
timeout 1200 python bench.py --config 5 --scale 0.1 --steps 1 --warmup 1 > gpurun_out/bench_r2d_c5_s01_n1.json 2>> gpurun_out/bench_r2d.err
tail -c 1500 gpurun_out/bench_r2d.err
python - <<PY
import json
for f in ('gpurun_out/bench_r2d.json','gpurun_out/bench_r2d_c4_n1.json','gpurun_out/bench_r2d_c3.json','gpurun_out/bench_r2d_c5_s01_n1.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f,{k:d.get(k) for k in ('value','ms_per_step','constraints_per_sec','parity_checked','parity_failed')}, d['e2e'].get('value'), d.get('cpu_baseline'))
        if d.get('roofline'): print({k:d['roofline'].get(k) for k in ('kernel','bound','frac','ncu_pct_of_peak','algorithmic_GBps','hbm')})
        print({k:d['config'].get(k) for k in ('jobs','found','submaps','nodes','stack_build_s_per_rank','matcher_build_s_per_rank')})
    except Exception as e:
        print(f,'ERR',e)
PY
