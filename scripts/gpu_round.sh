#!/bin/bash
mkdir -p gpurun_out
timeout 120 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
( timeout 300 python -m pytest tests -m gpu -x -q --durations=3 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log )
tail -8 gpurun_out/pytest_gpu.log
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cut -c1-330 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
