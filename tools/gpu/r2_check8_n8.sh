bash tools/gpu/manifest.sh check || exit 9
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 2> gpurun_out/bench_r2g_n8.err | grep "^{" > gpurun_out/bench_r2g_n8.json
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r2g_n8.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','n_gpus','ms_per_step','device_ms_per_step','collective_ms_per_step','constraints_per_sec')}, d['e2e'].get('value'), d['e2e'].get('ms_per_step'))
print(d['step_ms_per_rank'])
PY
tail -c 300 gpurun_out/bench_r2g_n8.err
