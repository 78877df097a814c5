# A/B timing of alternative builds of libcsm_b200 (cartographer_b200/csrc/_build/libcsm_b200_<name>.so)
# usage: bash tools/gpu/ab_variants.sh name1 name2 ...   ("default" = the product library)
mkdir -p gpurun_out
for v in "$@"; do
  unset CSM_B200_LIB
  if [ "$v" != default ]; then export CSM_B200_LIB=$PWD/cartographer_b200/csrc/_build/libcsm_b200_$v.so; fi
  timeout 300 python bench.py --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/ab_$v.json').read().strip().splitlines()[-1])
    print('$v', round(d['ms_per_step'],3), {k:(v['ms'],v['launches']) for k,v in d['roofline']['kernels'].items()})
except Exception as e:
    print('$v FAILED', e); print(open('gpurun_out/ab_$v.err').read()[-800:])
PY
  if [ -n "$AB_CFG4" ]; then
  timeout 600 python benchmarks/run_configs.py --configs 4 > gpurun_out/ab_cfg4_$v.jsonl 2> gpurun_out/ab_cfg4_$v.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/ab_cfg4_$v.jsonl').read().strip().splitlines()[-1])
print('$v cfg4', round(d['gpu_constraints_per_s']), round(d['gpu_wall_ms'],2), round(d['gpu_device_ms'],2), d['parity_ok'], {k:v['ms'] for k,v in d['gpu_kernels_one_batch'].items()})
PY
  fi
done
