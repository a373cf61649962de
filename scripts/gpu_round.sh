#!/bin/bash
# One gpurun call: UMMA layout probe, GPU parity tests, bench line, ncu launch list of one step.
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/src_*.csv gpurun_out/prof_*_raw.csv
for e in E0 E1 E2 E3 E4 E5; do timeout 120 python scripts/gpu_umma_probe.py $e 2>&1 | grep -v "^probe done" ; done > gpurun_out/umma_probe.log 2>&1
cat gpurun_out/umma_probe.log | cut -c1-200
( timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log )
tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python bench.py --profile-step > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches exit $?"
python scripts/summarize_launches.py gpurun_out/launches.csv | head -14
du -sh gpurun_out
