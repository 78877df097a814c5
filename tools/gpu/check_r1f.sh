set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_2d.py tests/test_gpu_edge_cases.py tests/test_gpu_constraint_builder.py -m gpu -x -q > gpurun_out/t2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t2.log
tail -5 gpurun_out/t2.log
for v in default minb10; do
  if [ $v = minb10 ]; then export CSM_B200_LIB=$PWD/cartographer_b200/csrc/_build/libcsm_b200_minb10.so; fi
  timeout 300 python bench.py --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/bench_r1f_$v.json 2> gpurun_out/bench_r1f_$v.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r1f_$v.json').read().strip().splitlines()[-1])
print('$v', d['ms_per_step'], d['value'], {k:(v['ms'],v['launches']) for k,v in d['roofline']['kernels'].items()})
PY
  timeout 600 python benchmarks/run_configs.py --configs 4 > gpurun_out/cfg4_r1f_$v.jsonl 2> gpurun_out/cfg4_r1f_$v.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/cfg4_r1f_$v.jsonl').read().strip().splitlines()[-1])
print('$v cfg4', d['gpu_constraints_per_s'], d['gpu_wall_ms'], d['gpu_device_ms'], d['parity_ok'], d['gpu_kernels_one_batch'])
PY
done
