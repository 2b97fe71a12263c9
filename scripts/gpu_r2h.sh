#!/bin/bash
# round 2, call H (2 GPUs): DP inside the boundary — tests, then bench at N=2
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_dp.py -x -q -m gpu > gpurun_out/pytest_r2h_dp.log 2>&1; echo "dp tests exit $?"; tail -12 gpurun_out/pytest_r2h_dp.log
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "adam" > gpurun_out/pytest_r2h_adam.log 2>&1; echo "adam tests exit $?"; tail -5 gpurun_out/pytest_r2h_adam.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_r2h_n2.log 2>&1; echo "bench N=2 exit $?"; tail -c 1800 gpurun_out/bench_r2h_n2.log
