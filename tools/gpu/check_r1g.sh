mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_2d.py tests/test_gpu_edge_cases.py tests/test_gpu_constraint_builder.py -m gpu -x -q > gpurun_out/t2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t2.log
tail -3 gpurun_out/t2.log
AB_CFG4=1 bash tools/gpu/ab_variants.sh default
for v in default t512 t1024; do
  unset CSM_B200_LIB
  if [ "$v" != default ]; then export CSM_B200_LIB=$PWD/cartographer_b200/csrc/_build/libcsm_b200_$v.so; fi
  CSM_SKIP_CPU=1 timeout 400 python benchmarks/run_configs.py --configs 3,5 > gpurun_out/ab3d_$v.jsonl 2> gpurun_out/ab3d_$v.err
  python - <<PY
import json
for ln in open('gpurun_out/ab3d_$v.jsonl'):
    if ln.startswith('{'):
        d=json.loads(ln)
        print('$v', d['config'], round(d['gpu_matches_per_s'],1), round(d['gpu_wall_ms_per_match'],3), {k:v['ms'] for k,v in d['kernels_of_one_found_match'].items()})
PY
done
