mkdir -p gpurun_out
for tool in racecheck synccheck; do
timeout 300 compute-sanitizer --tool $tool --error-exitcode 9 --log-file gpurun_out/${tool}_2d.log \
  python -m pytest tests/test_gpu_edge_cases.py tests/test_gpu_constraint_builder.py -m gpu -x -q > gpurun_out/${tool}_2d.out 2>&1
echo "$tool rc=$?"; tail -2 gpurun_out/${tool}_2d.out; tail -2 gpurun_out/${tool}_2d.log
done
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/racecheck_3d.log \
  python -m pytest tests/test_gpu_parity_3d.py -m gpu -x -q -k "building or batch" > gpurun_out/racecheck_3d.out 2>&1
echo "racecheck 3d rc=$?"; tail -2 gpurun_out/racecheck_3d.out; tail -2 gpurun_out/racecheck_3d.log
