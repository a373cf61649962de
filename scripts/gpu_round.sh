#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log )
grep -E "^E  |FAILED|passed|failed|pytest exit" gpurun_out/pytest_gpu.log | cut -c1-250 | head -40
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2>gpurun_out/bench.err; cut -c1-260 gpurun_out/bench.json; tail -2 gpurun_out/bench.err
FB_PDL=0 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_nopdl.json 2>/dev/null; cut -c1-260 gpurun_out/bench_nopdl.json
timeout 300 python bench.py --no-cpu-baseline --workload mc > gpurun_out/bench_mc.json 2>/dev/null; cut -c1-260 gpurun_out/bench_mc.json
timeout 300 python bench.py --no-cpu-baseline --workload bert > gpurun_out/bench_bert.json 2>/dev/null; cut -c1-260 gpurun_out/bench_bert.json
