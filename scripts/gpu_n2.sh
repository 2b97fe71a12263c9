#!/bin/bash
# GPU box with 2 GPUs: decoder-gradient hook test + N=2 bench (one process per GPU, NCCL)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "decoder_gradient_event" --timeout 200 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_n2.log 2>&1; echo "bench n2 exit $?"; tail -n 2 gpurun_out/bench_n2.log
