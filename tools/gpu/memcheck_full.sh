mkdir -p gpurun_out
for f in test_gpu_parity_2d test_gpu_parity_3d test_gpu_constraint_builder; do
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/memcheck_$f.log \
  python -m pytest tests/$f.py -m gpu -x -q > gpurun_out/memcheck_$f.out 2>&1
echo "memcheck $f rc=$?"; tail -2 gpurun_out/memcheck_$f.out; tail -1 gpurun_out/memcheck_$f.log
done
