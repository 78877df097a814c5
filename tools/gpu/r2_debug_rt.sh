bash tools/gpu/manifest.sh check || exit 9
set -x
mkdir -p gpurun_out
CSM_RT_NO_TMA=1 timeout 600 python -m pytest tests/test_gpu_rt2d.py -x -q > gpurun_out/rt_notma.log 2>&1; tail -15 gpurun_out/rt_notma.log
timeout 600 python -m pytest tests/test_gpu_rt2d.py -x -q -k config1 > gpurun_out/rt_tma.log 2>&1; tail -15 gpurun_out/rt_tma.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_rt2d.py -x -q -k config1 > gpurun_out/rt_tma_sanitizer.log 2>&1; grep -v "^$" gpurun_out/rt_tma_sanitizer.log | head -60
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_rt2d.py > gpurun_out/tests_gpu.log 2>&1; tail -25 gpurun_out/tests_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; tail -c 1500 gpurun_out/bench_r2a.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2a_2.json 2>> gpurun_out/bench_r2a.err
python - <<PY
import json
for f in ('gpurun_out/bench_r2a.json','gpurun_out/bench_r2a_2.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f,{k:d.get(k) for k in ('value','ms_per_step','device_ms_per_step','constraints_per_sec','gpu_launches','host_syncs_per_batch','parity_checked','parity_failed','clocks')}, d['e2e'], d.get('cpu_baseline'))
        print({k:v['ms'] for k,v in d['roofline']['kernels'].items()})
    except Exception as e:
        print(f,'ERR',e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_expand_lattice|k_score_top_tile' -s 6 -c 6 -o gpurun_out/r2a_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
