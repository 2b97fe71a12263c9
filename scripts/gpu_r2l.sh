#!/bin/bash
# round 2, call L (2 GPUs): full suite incl. 2-GPU tests, grad error report, bench
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 2400 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_r2l.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_r2l.log
timeout 300 python scripts/grad_error_report.py > gpurun_out/grad_report.log 2>&1; tail -12 gpurun_out/grad_report.log
timeout 600 python bench.py --steps 30 --warmup 3 --no-image --no-cpu > gpurun_out/bench_r2l.log 2>&1; echo "bench exit $?"; tail -c 600 gpurun_out/bench_r2l.log
