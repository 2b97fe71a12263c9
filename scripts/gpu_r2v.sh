#!/bin/bash
# round 2, last 2-GPU check of the final tree: sharded VAE.loss == single GPU
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 200 python -m pytest tests/test_gpu_dp.py -q -m gpu -k "sharded or bucket" > gpurun_out/pytest_r2z_dp.log 2>&1; echo "dp pytest exit $?"; tail -3 gpurun_out/pytest_r2z_dp.log
