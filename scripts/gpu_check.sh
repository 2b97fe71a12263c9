#!/bin/bash
# GPU box: what the driver runs at round end (full GPU suite, smoke, default bench) without the ncu passes.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/smoke.log
( time timeout 900 python bench.py ) > gpurun_out/bench_default.log 2>&1; echo "bench exit $?"; grep '^{' gpurun_out/bench_default.log | cut -c1-4000; tail -4 gpurun_out/bench_default.log | grep real
