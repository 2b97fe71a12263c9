#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_text_graph.py -q -m gpu > gpurun_out/pytest_r2s_graph.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_r2s_graph.log
