#!/bin/bash
# round 2, call D: probe variants + LSTM tile-image kernel (tests, trace, bench A/B of the producer-side fence)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1

timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_benchmarked_config.py -x -q -m gpu > gpurun_out/pytest_r2d_lstm.log 2>&1; echo "lstm tests exit $?"; tail -8 gpurun_out/pytest_r2d_lstm.log
timeout 300 python scripts/lstm_trace.py > gpurun_out/trace.log 2>&1; tail -22 gpurun_out/trace.log
LAGVAE_LSTM_PROD_FENCE=1 timeout 300 python scripts/lstm_trace.py > gpurun_out/trace_prodfence.log 2>&1; grep "kernel:" gpurun_out/trace_prodfence.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-image --no-cpu > gpurun_out/bench_r2d.log 2>&1; echo "bench exit $?"; tail -c 1500 gpurun_out/bench_r2d.log
