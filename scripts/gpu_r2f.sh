#!/bin/bash
# call Y: decoder z bias added inside the forward recurrence kernel
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests/test_gpu_text_graph.py tests/test_gpu_benchmarked_config.py tests/test_gpu_parity.py tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/pytest_r2y_quick.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_r2y_quick.log
timeout 300 python bench.py --no-image --no-cpu --steps 45 --warmup 5 > gpurun_out/bench_r2y.log 2>&1; echo "bench: $(tail -c 8000 gpurun_out/bench_r2y.log | grep -o '"ms_per_step": [0-9.]*\|"e2e": {"value": [0-9.]*\|"graph": {"value": [0-9.]*' | head -4 | tr '\n' ' ')"
