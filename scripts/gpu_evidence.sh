#!/bin/bash
# Round-end evidence (one gpurun call, ~12 GPU-minutes): smoke, full GPU suite, the three bench workloads + the
# reference arm, launch list of one step, `ncu --set full` of the roofline kernels and of EVERY kernel of a 2-layer
# step (raw pages exported to CSV on the box; gpurun_out must stay < 64 MiB), compute-sanitizer logs.
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/parity_report.json
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
( timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log )
tail -4 gpurun_out/pytest_gpu.log
for wl in mlm mc bert; do
  timeout 900 python bench.py --workload $wl > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl exit $?"; cut -c1-230 gpurun_out/bench_$wl.json
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"; cut -c1-230 gpurun_out/bench_ref.json
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python bench.py --profile-step > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches exit $?"
python scripts/summarize_launches.py gpurun_out/launches.csv > gpurun_out/step_launches.md 2>/dev/null; head -12 gpurun_out/step_launches.md
# roofline kernels, one unambiguous capture each (bench.py reads the DRAM bytes from the trimmed CSV)
PROBE_ONCE=1 timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:"gemm_kernel|attn_fwd|attn_bwd" --launch-skip 1 --launch-count 4 -o gpurun_out/prof_probes -f \
    python scripts/gpu_probe_kernels.py > gpurun_out/ncu_probes.log 2>&1; echo "ncu probes exit $?"
ncu -i gpurun_out/prof_probes.ncu-rep --page raw --csv > gpurun_out/probes_raw.csv 2>/dev/null
python scripts/ncu_summary.py --trim gpurun_out/ncu_raw_probes.csv gpurun_out/probes_raw.csv > gpurun_out/ncu_probes.md; cat gpurun_out/ncu_probes.md | cut -c1-200
# every kernel of a (2-layer) train step
timeout 1500 ncu --profile-from-start off --set full --clock-control none --kernel-name-base demangled -o gpurun_out/prof_all -f \
    python bench.py --layers 2 --profile-step > gpurun_out/ncu_all.log 2>&1; echo "ncu all exit $?"
ncu -i gpurun_out/prof_all.ncu-rep --page raw --csv > gpurun_out/all_raw.csv 2>/dev/null
python scripts/ncu_summary.py --distinct --rows --trim gpurun_out/ncu_raw_all_kernels.csv gpurun_out/all_raw.csv > gpurun_out/ncu_all_kernels.md; wc -l gpurun_out/ncu_all_kernels.md
rm -f gpurun_out/*.ncu-rep gpurun_out/all_raw.csv gpurun_out/probes_raw.csv
# compute-sanitizer on the hand-rolled synchronisation protocols
# (racecheck takes ~8 GPU-minutes; its findings are analysed in profiles/r2_sanitizer.md -- RACECHECK=1 re-runs it)
for tool in memcheck synccheck ${RACECHECK:+racecheck}; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/gpu_sanitize.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "sanitizer $tool exit $?"; tail -3 gpurun_out/sanitizer_$tool.log
done
timeout 100 python scripts/gpu_ln_time.py > gpurun_out/row_kernel_times.log 2>&1; timeout 100 python scripts/gpu_adapter_gemm_time.py >> gpurun_out/row_kernel_times.log 2>&1
timeout 200 python scripts/gpu_gemm_shapes.py >> gpurun_out/row_kernel_times.log 2>&1; tail -22 gpurun_out/row_kernel_times.log
du -sh gpurun_out
