#!/bin/bash
# scratch: the command of the most recent gpurun call
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
