#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_model_gpu.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log )
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python scripts/gpu_gemm_shapes.py > gpurun_out/gemm_shapes_cl1.log 2>&1; cat gpurun_out/gemm_shapes_cl1.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cut -c1-330 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
