bash tools/gpu/manifest.sh check || exit 9
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/tests_gpu.log 2>&1; tail -6 gpurun_out/tests_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2i.json 2> gpurun_out/bench_r2i.err; tail -c 400 gpurun_out/bench_r2i.err
timeout 900 python bench.py --config 3 --steps 2 --warmup 1 > gpurun_out/bench_r2i_c3.json 2>> gpurun_out/bench_r2i.err
timeout 900 python bench.py --config 5 --scale 0.1 --steps 1 --warmup 1 > gpurun_out/bench_r2i_c5_s01.json 2>> gpurun_out/bench_r2i.err
python - <<PY
import json
for f in ('gpurun_out/bench_r2i.json','gpurun_out/bench_r2i_c3.json','gpurun_out/bench_r2i_c5_s01.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f,{k:d.get(k) for k in ('value','ms_per_step','constraints_per_sec','host_syncs_per_batch','parity_checked','parity_failed')}, d['e2e'].get('value'), d.get('cpu_baseline'))
        if d.get('roofline'): print({k:d['roofline'].get(k) for k in ('kernel','bound','frac','ncu_pct_of_peak')})
    except Exception as e:
        print(f,'ERR',e)
PY
# compute-sanitizer memcheck over the kernels that are new in round 2 (bounded)
timeout 420 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_rt2d.py tests/test_gpu_rt3d.py tests/test_gpu_ingest.py tests/test_gpu_edge_cases.py -q -x > gpurun_out/r2_memcheck_new_kernels.log 2>&1; tail -8 gpurun_out/r2_memcheck_new_kernels.log
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity_3d.py -q -x -k "reference_3d or batch_equals or constraint_builder" > gpurun_out/r2_memcheck_3d.log 2>&1; tail -6 gpurun_out/r2_memcheck_3d.log
