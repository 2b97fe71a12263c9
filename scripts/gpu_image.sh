#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_image.py -m gpu -q --timeout 600 > gpurun_out/image.log 2>&1; echo "exit $?"; tail -40 gpurun_out/image.log
