#!/bin/bash
# round 2, call V3 (2 GPUs): data-parallel tests with the final defaults (one bucket after the backward)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests/test_gpu_dp.py -q -m gpu > gpurun_out/pytest_r2v_dp.log 2>&1; echo "dp pytest exit $?"; tail -5 gpurun_out/pytest_r2v_dp.log
LAGVAE_DP_OVERLAP=1 timeout 900 python -m pytest tests/test_gpu_dp.py -q -m gpu -k "sharded" > gpurun_out/pytest_r2v_dp_ov1.log 2>&1; echo "dp (early hand-over) pytest exit $?"; tail -2 gpurun_out/pytest_r2v_dp_ov1.log
