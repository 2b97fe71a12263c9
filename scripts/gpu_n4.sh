#!/bin/bash
# GPU box with 4 GPUs: N=4 bench (one process per GPU, NCCL), then the reference arm launched the same way
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi -L | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_n4.log 2>&1; echo "bench n4 exit $?"; grep '^{' gpurun_out/bench_n4.log | cut -c1-1500; tail -n 3 gpurun_out/bench_n4.log | cut -c1-300
