#!/bin/bash
# scratch: the command of the most recent gpurun call
mkdir -p gpurun_out
timeout 60 python scripts/gpu_adapter_gemm_time.py
timeout 300 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | tail -3
timeout 100 python scripts/gpu_gemm_shapes.py | tail -4
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pf.json 2> gpurun_out/bench_pf.err
cut -c100-330 gpurun_out/bench_pf.json; tail -2 gpurun_out/bench_pf.err
