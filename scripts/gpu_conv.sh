#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_image.py -m gpu -q --timeout 300 2>&1 | tail -25
timeout 300 env IMG_GRAPH=1 python scripts/bench_image.py 2>&1 | tail -n 1 | cut -c1-900
bash scripts/gpu_imgprof.sh
