#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/gpu_umma_timing.py > gpurun_out/umma_timing.log 2>&1; echo "timing exit $?"
cat gpurun_out/umma_timing.log | cut -c1-220
