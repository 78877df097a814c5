bash tools/gpu/manifest.sh check || exit 9
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 2> gpurun_out/bench_r2e_n8.err | grep "^{" > gpurun_out/bench_r2e_n8.json
CSM_BENCH_NO_COLLECTIVE=1 timeout 300 $TR --master-port 29522 bench.py --gpus 8 --steps 10 --warmup 3 2>> gpurun_out/bench_r2e_n8.err | grep "^{" > gpurun_out/bench_r2e_n8_nocoll.json
python - <<PY
import json
for f in ('gpurun_out/bench_r2e_n8.json','gpurun_out/bench_r2e_n8_nocoll.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f,{k:d.get(k) for k in ('value','n_gpus','ms_per_step','device_ms_per_step','collective_ms_per_step','constraints_per_sec')}, d['e2e'].get('value'), d['e2e'].get('ms_per_step'))
        print(d['step_ms_per_rank'])
    except Exception as e:
        print(f,'ERR',e)
PY
tail -c 500 gpurun_out/bench_r2e_n8.err
