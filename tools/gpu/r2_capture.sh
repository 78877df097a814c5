bash tools/gpu/manifest.sh check || exit 9
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_expand_lattice|k_score_top_tile' -c 16 -o gpurun_out/r2j_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2j.json 2> gpurun_out/bench_r2j.err; tail -c 300 gpurun_out/bench_r2j.err
