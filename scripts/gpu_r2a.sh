#!/bin/bash
# round 2, call A: new parity tests first (fail fast), then the whole GPU suite, then a short bench.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_benchmarked_config.py -x -q -m gpu > gpurun_out/pytest_r2a_bench_cfg.log 2>&1; echo "bench-config tests exit $?"; tail -15 gpurun_out/pytest_r2a_bench_cfg.log
timeout 1500 python -m pytest tests/test_gpu_reference_drivers.py -q -m gpu > gpurun_out/pytest_r2a_drivers.log 2>&1; echo "driver tests exit $?"; tail -40 gpurun_out/pytest_r2a_drivers.log
timeout 1200 python -m pytest tests/ -x -q -m gpu --deselect tests/test_gpu_reference_drivers.py --deselect tests/test_gpu_benchmarked_config.py > gpurun_out/pytest_r2a_rest.log 2>&1; echo "rest exit $?"; tail -5 gpurun_out/pytest_r2a_rest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-image > gpurun_out/bench_r2a.log 2>&1; echo "bench exit $?"; tail -c 2500 gpurun_out/bench_r2a.log
