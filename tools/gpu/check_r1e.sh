set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_2d.py tests/test_gpu_edge_cases.py tests/test_gpu_constraint_builder.py -m gpu -x -q > gpurun_out/t2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t2.log
tail -5 gpurun_out/t2.log
timeout 300 python bench.py --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/bench_r1e.json 2> gpurun_out/bench_r1e.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r1e.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], {k:(v['ms'],v['launches']) for k,v in d['roofline']['kernels'].items()})
PY
timeout 600 python benchmarks/run_configs.py --configs 4 > gpurun_out/cfg4_r1e.jsonl 2> gpurun_out/cfg4_r1e.err
tail -2 gpurun_out/cfg4_r1e.jsonl
