#!/bin/bash
# round 2, final validation: whole GPU suite + smoke + default bench + reference arm + large-batch sweep + LSTM trace
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 2400 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_r2u.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_r2u.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2u.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke_r2u.log
timeout 1200 python bench.py > gpurun_out/bench_r2u.log 2>&1; echo "bench exit $?"; tail -c 6000 gpurun_out/bench_r2u.log | grep -o '"value": [0-9.]*, "unit": "steps/s", "n_gpus": 1\|"ms_per_step": [0-9.]*\|"e2e": {"value": [0-9.]*\|"graph": {"value": [0-9.]*' | head -5 | tr '\n' ' '; echo
timeout 600 python scripts/bench_sweep.py > gpurun_out/sweep_r2u.jsonl 2> gpurun_out/sweep_r2u.err; echo "sweep exit $?"; grep -o '"B": [0-9]*\|"ms_per_step": [0-9.]*' gpurun_out/sweep_r2u.jsonl | tr '\n' ' '; echo
LAGVAE_GATHER_SPLIT=0 timeout 300 python bench.py --no-image --no-cpu --no-e2e --steps 45 --warmup 5 > gpurun_out/bench_r2u_nogs.log 2>&1; echo "no gather-split: $(tail -c 6000 gpurun_out/bench_r2u_nogs.log | grep -o '"ms_per_step": [0-9.]*' | head -1)"
