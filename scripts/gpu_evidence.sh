#!/bin/bash
# Round-end evidence: smoke, full GPU suite, bench (with cpu_baseline), reference arm, launch list, ncu --set full of the
# dominant kernels (raw + source pages exported to CSV on the box; gpurun_out must stay < 64 MiB).
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
( timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log )
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cut -c1-330 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"; cut -c1-300 gpurun_out/bench_ref.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python bench.py --profile-step > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches exit $?"
python scripts/summarize_launches.py gpurun_out/launches.csv 2>/dev/null | head -14
for spec in "gemm:gemm_kernel<.int.256, .bool.0, .bool.0, .bool.0, .int.2>:4" "attnfwd:attn_fwd_kernel:1" "attnbwd:attn_bwd:2"; do
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
du -sh gpurun_out
