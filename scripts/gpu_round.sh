#!/bin/bash
# scratch: the command of the most recent gpurun call
mkdir -p gpurun_out
for ns in 0 100 400; do echo "FB_GEMM_RELAXED_NS=$ns"; FB_GEMM_RELAXED_NS=$ns timeout 100 python scripts/gpu_gemm_shapes.py | tail -8; done
