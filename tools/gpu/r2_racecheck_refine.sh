bash tools/gpu/manifest.sh check || exit 9
set -x
mkdir -p gpurun_out
timeout 40 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_gpu_ceres2d.py -q -x -k "known_answers or smoke" > gpurun_out/r2_racecheck_ceres2d.log 2>&1; tail -5 gpurun_out/r2_racecheck_ceres2d.log
timeout 40 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_gpu_ceres3d.py -q -x -k "known_answers" > gpurun_out/r2_racecheck_ceres3d.log 2>&1; tail -5 gpurun_out/r2_racecheck_ceres3d.log
