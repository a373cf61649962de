#!/bin/bash
mkdir -p gpurun_out
timeout 25 python scripts/gpu_umma_m64_probe.py > gpurun_out/umma_m64_probe.log 2>&1; echo "probe exit $?"; head -70 gpurun_out/umma_m64_probe.log | cut -c1-160
