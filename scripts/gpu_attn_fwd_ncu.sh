#!/bin/bash
# ncu --set full of the attention forward (generation given by $1, default 2) at the configs[1] shape: raw + source pages as CSV
mkdir -p gpurun_out
for gen in ${1:-2}; do
  FB_ATTN_FWD=$gen timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
      -k regex:attn_fwd -s 4 -c 1 -o gpurun_out/prof_attnfwd_g$gen -f python scripts/gpu_attn_time.py > gpurun_out/ncu_attnfwd_g$gen.log 2>&1
  echo "ncu gen $gen exit $?"
  ncu -i gpurun_out/prof_attnfwd_g$gen.ncu-rep --page raw --csv > gpurun_out/attnfwd_g${gen}_raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_attnfwd_g$gen.ncu-rep --page source --csv > gpurun_out/attnfwd_g${gen}_src.csv 2>/dev/null
  python scripts/ncu_summary.py gpurun_out/attnfwd_g${gen}_raw.csv
done
rm -f gpurun_out/*.ncu-rep
