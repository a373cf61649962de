#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/gpu_attn_time.py > gpurun_out/attn_time.log 2>&1; echo "attn time exit $?"; cat gpurun_out/attn_time.log | tail -6
( timeout 900 python -m pytest tests/test_attention_bwd_gpu.py tests/test_model_gpu.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log )
tail -4 gpurun_out/pytest_gpu.log
