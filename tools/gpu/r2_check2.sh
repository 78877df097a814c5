bash tools/gpu/manifest.sh check || exit 9
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader; nproc
# 1. the TMA question first: global-gather form, then the TMA form, then under the sanitizer
CSM_RT_NO_TMA=1 timeout 600 python -m pytest tests/test_gpu_rt2d.py -x -q > gpurun_out/rt_notma.log 2>&1; tail -8 gpurun_out/rt_notma.log
timeout 600 python -m pytest tests/test_gpu_rt2d.py -x -q -k config1 > gpurun_out/rt_tma.log 2>&1; tail -8 gpurun_out/rt_tma.log
if ! grep -q "passed" gpurun_out/rt_tma.log; then
  timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_rt2d.py -x -q -k config1 > gpurun_out/rt_tma_sanitizer.log 2>&1; grep -v "^$" gpurun_out/rt_tma_sanitizer.log | head -40
fi
# 2. every GPU test (the RT file separately so that one failure does not hide the rest)
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_rt2d.py > gpurun_out/tests_gpu.log 2>&1; tail -30 gpurun_out/tests_gpu.log
timeout 600 python -m pytest tests/test_gpu_rt2d.py -q > gpurun_out/tests_gpu_rt.log 2>&1; tail -12 gpurun_out/tests_gpu_rt.log
# 3. headline bench twice (+ launch list + full captures of the two dominant kernels)
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
# 4. the other configs (3D after the sync-free rewrite; config 4 at 1/10 size)
timeout 900 python benchmarks/run_configs.py --configs 3,5 > gpurun_out/configs_35.jsonl 2> gpurun_out/configs_35.err; tail -c 600 gpurun_out/configs_35.err
python - <<PY
import json
for ln in open('gpurun_out/configs_35.jsonl'):
    if ln.startswith('{'):
        d=json.loads(ln); print(d['config'], {k:(round(v,3) if isinstance(v,float) else v) for k,v in d.items() if k.startswith(('gpu_','cpu_','parity'))})
PY
timeout 900 python bench.py --config 4 --scale 0.1 --steps 2 --warmup 1 > gpurun_out/bench_c4_s01.json 2> gpurun_out/bench_c4.err; tail -c 800 gpurun_out/bench_c4.err; tail -c 1500 gpurun_out/bench_c4_s01.json
timeout 300 python bench.py --config 1 --steps 5 --warmup 3 > gpurun_out/bench_r2a_rt.json 2>> gpurun_out/bench_r2a.err; tail -c 1200 gpurun_out/bench_r2a_rt.json
ls -la gpurun_out | head -40
