#!/bin/bash
# ncu --set full of the fused attention backward + band-gradient kernel at the configs[1] shape: raw + source pages as CSV
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:attn_bwd -s 6 -c 2 -o gpurun_out/prof_attnbwd -f python scripts/gpu_attn_time.py > gpurun_out/ncu_attnbwd.log 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/prof_attnbwd.ncu-rep --page raw --csv > gpurun_out/attnbwd_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_attnbwd.ncu-rep --page source --csv --kernel-name regex:attn_bwd_kernel > gpurun_out/attnbwd_src.csv 2>/dev/null
ncu -i gpurun_out/prof_attnbwd.ncu-rep --page source --csv --kernel-name regex:attn_bwd_rel > gpurun_out/attnrel_src.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/attnbwd_raw.csv
rm -f gpurun_out/*.ncu-rep
