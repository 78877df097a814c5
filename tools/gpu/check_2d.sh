# 2D parity tests + adapter self-test + bench/config-4 kernel breakdown (1 GPU)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_2d.py tests/test_gpu_edge_cases.py tests/test_gpu_constraint_builder.py tests/test_gpu_adapter.py -m gpu -x -q > gpurun_out/t2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t2.log
tail -5 gpurun_out/t2.log
AB_CFG4=1 bash tools/gpu/ab_variants.sh default
