# ncu captures for the round-1 kernels (run under gpurun, 1 GPU)
set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_score_top_tile -c 1 -o gpurun_out/prof_r1d_top -f python bench.py --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/ncu_top.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_expand_lattice -s 5 -c 2 -o gpurun_out/prof_r1d_lattice -f python bench.py --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/ncu_lat.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/ncu_launch.log 2>&1
ls -la gpurun_out
