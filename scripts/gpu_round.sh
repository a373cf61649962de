#!/bin/bash
mkdir -p gpurun_out
timeout 105 python bench.py --steps 6 > gpurun_out/bench_final.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench_final.json | cut -c1-200; tail -2 gpurun_out/bench.err
