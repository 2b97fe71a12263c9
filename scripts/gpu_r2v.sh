#!/bin/bash
# round 2, call V (2 GPUs): data-parallel tests + N=2 bench
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests/test_gpu_dp.py -q -m gpu > gpurun_out/pytest_r2v_dp.log 2>&1; echo "dp pytest exit $?"; tail -5 gpurun_out/pytest_r2v_dp.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 3 > gpurun_out/bench_r2v_n2.log 2>&1; echo "bench exit $?"; tail -c 2500 gpurun_out/bench_r2v_n2.log | grep -o '"value": [0-9.]*, "unit": "steps/s", "n_gpus": 2\|"ms_per_step": [0-9.]*\|"e2e": {"value": [0-9.]*'
