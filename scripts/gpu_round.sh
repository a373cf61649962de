#!/bin/bash
# scratch: launch list of one step
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python bench.py --profile-step > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches exit $?"
python scripts/summarize_launches.py gpurun_out/launches.csv | head -50
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2>/dev/null; cut -c1-260 gpurun_out/bench.json
