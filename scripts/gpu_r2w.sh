#!/bin/bash
# round 2, call W (8 GPUs): N=8 weak-scaling bench with the final defaults
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 50 --warmup 3 > gpurun_out/bench_r2w_n8.log 2>&1; echo "bench exit $?"; tail -c 3500 gpurun_out/bench_r2w_n8.log | grep -o '"value": [0-9.]*, "unit": "steps/s", "n_gpus": 8\|"ms_per_step": [0-9.]*\|"e2e": {"value": [0-9.]*' | tr '\n' ' '
