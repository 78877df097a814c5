# compute-sanitizer memcheck over the small-case GPU tests (slow: run on a subset)
mkdir -p gpurun_out
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/memcheck_2d.log \
  python -m pytest tests/test_gpu_edge_cases.py -m gpu -x -q -k "not full_size" > gpurun_out/memcheck_2d.out 2>&1
echo "memcheck 2d rc=$?"; tail -3 gpurun_out/memcheck_2d.out; grep -c "Invalid\|out of bounds" gpurun_out/memcheck_2d.log; tail -3 gpurun_out/memcheck_2d.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/memcheck_adapter.log \
  cartographer_b200/adapter/adapter_selftest > gpurun_out/memcheck_adapter.out 2>&1
echo "memcheck adapter rc=$?"; tail -4 gpurun_out/memcheck_adapter.out; tail -2 gpurun_out/memcheck_adapter.log
