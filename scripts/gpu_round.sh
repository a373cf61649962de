#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 300 python scripts/gpu_attn_time.py > gpurun_out/attn_time.log 2>&1; echo "attn time exit $?"; cat gpurun_out/attn_time.log | tail -6
( timeout 900 python -m pytest tests/test_attention_bwd_gpu.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log )
tail -4 gpurun_out/pytest_gpu.log
for spec in "attnfwd:attn_fwd_kernel:1" "attnbwd:attn_bwd:2"; do
  IFS=: read tag pat cnt <<< "$spec"
  timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
      --kernel-name-base demangled -k regex:"$pat" -c $cnt -o gpurun_out/prof_$tag -f \
      python bench.py --layers 2 --profile-step > gpurun_out/ncu_full_$tag.log 2>&1; echo "ncu full $tag exit $?"
  ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
done
ncu -i gpurun_out/prof_attnfwd.ncu-rep --page source --csv > gpurun_out/src_attnfwd.csv 2>/dev/null
ncu -i gpurun_out/prof_attnbwd.ncu-rep --page source --csv --kernel-name regex:attn_bwd_kernel > gpurun_out/src_attnbwd.csv 2>/dev/null
ncu -i gpurun_out/prof_attnbwd.ncu-rep --page source --csv --kernel-name regex:attn_bwd_rel > gpurun_out/src_attnrel.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
