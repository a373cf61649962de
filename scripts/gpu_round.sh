#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/gpu_attn_bwd_diag.py > gpurun_out/attn_bwd_diag.log 2>&1; echo "diag exit $?"; grep -c "nan: 0" gpurun_out/attn_bwd_diag.log; grep -E "EXCEPTION" gpurun_out/attn_bwd_diag.log | head
timeout 300 python scripts/gpu_attn_bwd_timeline.py > gpurun_out/attn_bwd_timeline.log 2>&1; echo "timeline exit $?"
cat gpurun_out/attn_bwd_timeline.log | cut -c1-420
( timeout 900 python -m pytest tests/test_attention_bwd_gpu.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log )
tail -4 gpurun_out/pytest_gpu.log
