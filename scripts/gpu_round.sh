#!/bin/bash
# scratch: the command of the most recent gpurun call
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_attention_gpu.py tests/test_attention_bwd_gpu.py tests/test_model_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log )
grep -E "^E  |FAILED|passed|failed|pytest exit" gpurun_out/pytest_gpu.log | cut -c1-250 | head -40
FB_ATTN_FWD=1 timeout 120 python scripts/gpu_attn_time.py > gpurun_out/attn_time_v1.log 2>&1; tail -4 gpurun_out/attn_time_v1.log
timeout 120 python scripts/gpu_attn_time.py > gpurun_out/attn_time_v2.log 2>&1; tail -4 gpurun_out/attn_time_v2.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --no-cpu-baseline --workload mc > gpurun_out/bench_mc.json 2> gpurun_out/bench_mc.err; cut -c1-300 gpurun_out/bench_mc.json
