#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/gpu_probe_kernels.py 2>&1 | tail -5
timeout 200 python scripts/gpu_sanitize.py 2>&1 | tail -3
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/gpu_sanitize.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "sanitizer $tool exit $?"; tail -4 gpurun_out/sanitizer_$tool.log
done
