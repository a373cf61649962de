#!/bin/bash
# scratch: the command of the most recent gpurun call
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | tail -4
for eg in 2 0 3; do
  FB_GEMM_EG=$eg timeout 200 python scripts/gpu_gemm_shapes.py > gpurun_out/gemm_shapes_eg$eg.log 2>&1; echo "EG=$eg"; tail -9 gpurun_out/gemm_shapes_eg$eg.log
done
for eg in 0 2; do
  FB_GEMM_EG=$eg timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_eg$eg.json 2> gpurun_out/bench_eg$eg.err
  cut -c1-330 gpurun_out/bench_eg$eg.json | cut -c100-330; tail -2 gpurun_out/bench_eg$eg.err
done
