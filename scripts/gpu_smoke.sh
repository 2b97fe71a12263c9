#!/bin/bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_image.py -m gpu -q -k "micro_batch or all_gradients" --timeout 300 2>&1 | tail -6
timeout 600 python scripts/bench_sweep.py > gpurun_out/sweep.log 2>&1; cut -c1-200 gpurun_out/sweep.log
