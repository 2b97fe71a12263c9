#!/bin/bash
# last call of the round: configs[2] (Yelp, B=32, T=100) on one GPU with the final tree
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 120 python scripts/bench_strong.py --steps 40 --warmup 5 > gpurun_out/strong_r2z_n1.log 2>&1; echo "exit $?"; tail -1 gpurun_out/strong_r2z_n1.log | cut -c1-600
