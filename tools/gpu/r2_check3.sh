bash tools/gpu/manifest.sh check || exit 9
set -x
mkdir -p gpurun_out
# 1. the TMA fault: with box origins aligned to 8 cells; else origin (0,0) only; sanitizer on a failing case
timeout 600 python -m pytest tests/test_gpu_rt2d.py -q > gpurun_out/rt_tma_aligned.log 2>&1; tail -8 gpurun_out/rt_tma_aligned.log
if ! grep -q "7 passed" gpurun_out/rt_tma_aligned.log; then
  CSM_RT_ORIGIN0=1 timeout 600 python -m pytest tests/test_gpu_rt2d.py -q > gpurun_out/rt_tma_origin0.log 2>&1; tail -8 gpurun_out/rt_tma_origin0.log
  timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_rt2d.py -x -q -k default_options > gpurun_out/rt_tma_sanitizer.log 2>&1; grep -v "^$" gpurun_out/rt_tma_sanitizer.log | head -60
  export CSM_RT_NO_TMA=1
fi
# 2. every GPU test
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/tests_gpu.log 2>&1; tail -40 gpurun_out/tests_gpu.log
# 3. headline bench + tiled A/B + captures
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; tail -c 800 gpurun_out/bench_r2b.err
CSM_B200_LIB=$PWD/cartographer_b200/libcsm_b200_tiled.so timeout 600 python -m pytest tests/test_gpu_parity_2d.py tests/test_gpu_edge_cases.py tests/test_gpu_golden.py -q > gpurun_out/tests_tiled.log 2>&1; tail -5 gpurun_out/tests_tiled.log
CSM_B200_LIB=$PWD/cartographer_b200/libcsm_b200_tiled.so timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2b_tiled.json 2>> gpurun_out/bench_r2b.err
python - <<PY
import json
for f in ('gpurun_out/bench_r2b.json','gpurun_out/bench_r2b_tiled.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f,{k:d.get(k) for k in ('value','ms_per_step','device_ms_per_step','constraints_per_sec','gpu_launches','host_syncs_per_batch','parity_checked','parity_failed')}, d['e2e'])
        print({k:v['ms'] for k,v in d['roofline']['kernels'].items()})
    except Exception as e:
        print(f,'ERR',e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_expand_lattice|k_score_top_tile' -c 14 -o gpurun_out/r2b_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
CSM_B200_LIB=$PWD/cartographer_b200/libcsm_b200_tiled.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_expand_lattice' -c 6 -o gpurun_out/r2b_full_tiled python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_tiled.log 2>&1
# 4. other configs
timeout 900 python benchmarks/run_configs.py --configs 3,5 > gpurun_out/configs_35.jsonl 2> gpurun_out/configs_35.err; tail -c 600 gpurun_out/configs_35.err
python - <<PY
import json
for ln in open('gpurun_out/configs_35.jsonl'):
    if ln.startswith('{'):
        d=json.loads(ln); print(d['config'], {k:(round(v,3) if isinstance(v,float) else v) for k,v in d.items() if k.startswith(('gpu_','cpu_','parity'))}); print(d.get('kernels_of_one_found_match'))
PY
timeout 300 python bench.py --config 1 --steps 5 --warmup 3 > gpurun_out/bench_r2b_rt.json 2>> gpurun_out/bench_r2b.err; tail -c 1500 gpurun_out/bench_r2b_rt.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_rt_match' -c 2 -o gpurun_out/r2b_rt python bench.py --config 1 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_rt.log 2>&1; tail -3 gpurun_out/ncu_rt.log
ls -la gpurun_out | head -50
