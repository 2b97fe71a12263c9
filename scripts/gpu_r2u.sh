#!/bin/bash
# round 2, final validation: whole GPU suite + smoke + default bench + reference arm + LSTM trace
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 2400 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_r2u.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_r2u.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2u.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke_r2u.log
timeout 1200 python bench.py > gpurun_out/bench_r2u.log 2>&1; echo "bench exit $?"; tail -c 6000 gpurun_out/bench_r2u.log | grep -o '"value": [0-9.]*, "unit": "steps/s", "n_gpus": 1\|"ms_per_step": [0-9.]*\|"e2e": {"value": [0-9.]*\|"graph": {"value": [0-9.]*' | head -5 | tr '\n' ' '; echo
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2u_ref.log 2>&1; echo "ref arm exit $?"
python scripts/lstm_trace.py > gpurun_out/trace_r2u.log 2>&1; grep -E "kernel:|next step" gpurun_out/trace_r2u.log
