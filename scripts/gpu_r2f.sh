#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_benchmarked_config.py -x -q -m gpu -k "lstm" > gpurun_out/pytest_lstm_quick.log 2>&1; tail -2 gpurun_out/pytest_lstm_quick.log
python scripts/lstm_trace.py > gpurun_out/trace.log 2>&1; grep -E "kernel:|first stage|MMAs issued|accumulators|cell done|next step" gpurun_out/trace.log
