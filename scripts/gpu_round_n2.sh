#!/bin/bash
# 2-GPU call: data-parallel correctness test + bench (one NCCL all-reduce of the flat bucket at the end of backward)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n2.log 2>&1
( timeout 900 python -m pytest tests/test_distributed_gpu.py tests/test_dropin_gpu.py -m gpu -q --timeout 800 -p no:cacheprovider > gpurun_out/pytest_n2.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_n2.log )
grep -E "^E  |FAILED|passed|failed|skipped|pytest exit" gpurun_out/pytest_n2.log | cut -c1-250 | head -40
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 12 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 exit $?"
cut -c1-300 gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
