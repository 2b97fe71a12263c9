#!/bin/bash
# round 2, call G: whole GPU suite + smoke + default bench with the new LSTM operand path
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1800 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_r2g.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_r2g.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2g.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke_r2g.log
timeout 900 python bench.py --no-image > gpurun_out/bench_r2g.log 2>&1; echo "bench exit $?"; tail -c 4000 gpurun_out/bench_r2g.log
