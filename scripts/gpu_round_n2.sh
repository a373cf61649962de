#!/bin/bash
# 2-GPU call: data-parallel bench (NCCL all-reduce of the flat gradient bucket over NVLink)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n2.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 exit $?"
cat gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
