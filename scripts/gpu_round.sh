#!/bin/bash
# scratch: the command of the most recent gpurun call
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_relbatched.json 2> gpurun_out/bench_relbatched.err
cut -c1-330 gpurun_out/bench_relbatched.json; tail -3 gpurun_out/bench_relbatched.err
FB_REL_BATCHED=0 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_relbatched0.json 2> gpurun_out/bench_relbatched0.err
cut -c1-330 gpurun_out/bench_relbatched0.json
