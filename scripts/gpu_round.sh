#!/bin/bash
# scratch: the command of the most recent gpurun call
mkdir -p gpurun_out
timeout 1200 ncu --profile-from-start off --set full --clock-control none --kernel-name-base demangled -o gpurun_out/prof_all -f \
    python bench.py --layers 2 --profile-step > gpurun_out/ncu_all.log 2>&1; echo "ncu all exit $?"
ncu -i gpurun_out/prof_all.ncu-rep --page raw --csv > gpurun_out/all_raw.csv 2>/dev/null
python scripts/ncu_summary.py --distinct --rows --trim gpurun_out/ncu_raw_all_kernels.csv gpurun_out/all_raw.csv > gpurun_out/ncu_all_kernels.md; wc -l gpurun_out/ncu_all_kernels.md
rm -f gpurun_out/*.ncu-rep gpurun_out/all_raw.csv
head -30 gpurun_out/ncu_all_kernels.md | cut -c1-170
