#!/bin/bash
# round 2, call B: fixed parity/driver tests + exchange-protocol probe
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 120 scripts/microbench/exchange_probe 400 0 > gpurun_out/exchange_probe_d0.txt 2>&1; echo "probe exit $?"; cat gpurun_out/exchange_probe_d0.txt

timeout 900 python -m pytest tests/test_gpu_benchmarked_config.py -q -m gpu > gpurun_out/pytest_r2c_bench_cfg.log 2>&1; echo "bench-config tests exit $?"; tail -15 gpurun_out/pytest_r2c_bench_cfg.log
timeout 1500 python -m pytest tests/test_gpu_reference_drivers.py -q -m gpu > gpurun_out/pytest_r2c_drivers.log 2>&1; echo "driver tests exit $?"; tail -30 gpurun_out/pytest_r2c_drivers.log
