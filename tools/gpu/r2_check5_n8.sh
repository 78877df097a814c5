bash tools/gpu/manifest.sh check || exit 9
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout 600 $TR --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_r2d_n8.json 2> gpurun_out/bench_r2d_n8.err; grep -m6 "NCCL INFO.*\(comm\|ncclCommInitRank\|Connected\|NVLS\)" gpurun_out/bench_r2d_n8.err | cut -c1-200
timeout 900 $TR --master-port 29512 bench.py --gpus 8 --config 4 --steps 2 --warmup 1 > gpurun_out/bench_r2d_c4_n8.json 2> gpurun_out/bench_r2d_c4_n8.err; tail -c 600 gpurun_out/bench_r2d_c4_n8.err
timeout 1200 $TR --master-port 29513 bench.py --gpus 8 --config 5 --steps 1 --warmup 0 > gpurun_out/bench_r2d_c5_n8.json 2> gpurun_out/bench_r2d_c5_n8.err; tail -c 600 gpurun_out/bench_r2d_c5_n8.err
python - <<PY
import json
for f in ('gpurun_out/bench_r2d_n8.json','gpurun_out/bench_r2d_c4_n8.json','gpurun_out/bench_r2d_c5_n8.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f,{k:d.get(k) for k in ('value','n_gpus','ms_per_step','constraints_per_sec','parity_checked','parity_failed')}, d['e2e'].get('value'), d.get('cpu_baseline'))
        print({k:d['config'].get(k) for k in ('jobs','found','submaps','nodes','collective','stack_build_s_per_rank','matcher_build_s_per_rank','host_generation_s')})
    except Exception as e:
        print(f,'ERR',e)
PY
