#!/bin/bash
# scratch: the command of the most recent gpurun call
mkdir -p gpurun_out
timeout 100 python scripts/gpu_probe_kernels.py
timeout 300 python -m pytest tests/test_attention_bwd_gpu.py -x -q 2>&1 | tail -2
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_relpf.json 2> gpurun_out/bench_relpf.err
cut -c100-330 gpurun_out/bench_relpf.json; tail -2 gpurun_out/bench_relpf.err
