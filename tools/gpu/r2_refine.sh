bash tools/gpu/manifest.sh check || exit 9
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/tests_gpu_k.log 2>&1; tail -15 gpurun_out/tests_gpu_k.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_k.log 2>&1; tail -3 gpurun_out/smoke_k.log
timeout 300 python -m benchmarks.bench_refine2d > gpurun_out/bench_refine2d.json 2> gpurun_out/bench_refine2d.err; tail -c 600 gpurun_out/bench_refine2d.json; tail -c 300 gpurun_out/bench_refine2d.err
timeout 240 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_ceres2d.py -q -x -k "smoke or residuals or known_answers or batch_over" > gpurun_out/r2_memcheck_ceres2d.log 2>&1; tail -6 gpurun_out/r2_memcheck_ceres2d.log
