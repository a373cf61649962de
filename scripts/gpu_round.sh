#!/bin/bash
# scratch: the command of the most recent gpurun call
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 100 python scripts/gpu_ln_time.py
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ln4.json 2> gpurun_out/bench_ln4.err
cut -c100-330 gpurun_out/bench_ln4.json; tail -2 gpurun_out/bench_ln4.err
