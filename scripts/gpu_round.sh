#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log )
tail -4 gpurun_out/pytest_gpu.log
FB_DW_STREAM=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_dw0.json 2> gpurun_out/bench.err; echo "bench dw0 exit $?"; cut -c1-330 gpurun_out/bench_dw0.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cut -c1-330 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
