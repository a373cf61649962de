#!/bin/bash
# One gpurun call: attention-backward diagnostics + GPU parity tests, bench line, ncu launch list of one step,
# ncu --set full of the attention kernels (reports exported to CSV on the box; gpurun_out must stay < 64 MiB).
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 300 python scripts/gpu_attn_bwd_diag.py > gpurun_out/attn_bwd_diag.log 2>&1; echo "diag exit $?"; tail -40 gpurun_out/attn_bwd_diag.log
( timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log )
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
FB_ATTN_BWD=staged timeout 600 python bench.py --no-cpu-baseline --steps 4 > gpurun_out/bench_staged.json 2> gpurun_out/bench_staged.err; cat gpurun_out/bench_staged.json
# launch list of exactly one resident training step (cudaProfilerStart/Stop around it)
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python bench.py --profile-step > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches exit $?"
# full captures on a 2-layer model (identical per-layer shapes)
for spec in "attnfwd:attn_fwd_kernel:1" "attnbwd:attn_bwd:2"; do
  IFS=: read tag pat cnt <<< "$spec"
  timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
      -k regex:"$pat" -c $cnt -o gpurun_out/prof_$tag -f \
      python bench.py --layers 2 --profile-step > gpurun_out/ncu_full_$tag.log 2>&1; echo "ncu full $tag exit $?"
  ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
done
du -sh gpurun_out; ls -la gpurun_out
