bash tools/gpu/manifest.sh check || exit 9
set -x
mkdir -p gpurun_out
timeout 170 python -m pytest tests -m gpu -q -x > gpurun_out/tests_gpu_l.log 2>&1; tail -15 gpurun_out/tests_gpu_l.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_l.log 2>&1; tail -3 gpurun_out/smoke_l.log
timeout 80 python -m benchmarks.bench_refine3d > gpurun_out/bench_refine3d.json 2> gpurun_out/bench_refine3d.err; tail -c 700 gpurun_out/bench_refine3d.json; tail -c 300 gpurun_out/bench_refine3d.err
timeout 50 python -m benchmarks.bench_refine2d > gpurun_out/bench_refine2d_b.json 2> gpurun_out/bench_refine2d_b.err; tail -c 600 gpurun_out/bench_refine2d_b.json
timeout 90 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_ceres3d.py -q -x -k "known_answers or residuals or batch_over" > gpurun_out/r2_memcheck_ceres3d.log 2>&1; tail -6 gpurun_out/r2_memcheck_ceres3d.log
