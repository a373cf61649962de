#!/bin/bash
# scratch: the command of the most recent gpurun call
mkdir -p gpurun_out
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 200 python scripts/gpu_ffn1_probe.py 2>&1 | tail -8
