#!/bin/bash
# round 2, call R: text inner loop as a CUDA graph (device Philox word, per-capture decoder-weight epoch), bench graph leg,
# decoder x-projection on a side stream under the encoder recurrence
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_text_graph.py -x -q -m gpu > gpurun_out/pytest_r2r_graph.log 2>&1; echo "graph pytest exit $?"; tail -30 gpurun_out/pytest_r2r_graph.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_benchmarked_config.py tests/test_gpu_image.py -x -q -m gpu -k "fused or inner or step or drop or philox or graph or train or forward" > gpurun_out/pytest_r2r_quick.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_r2r_quick.log
LAGVAE_OVERLAP_XPROJ=0 timeout 600 python bench.py --no-image --no-cpu --no-e2e --steps 45 --warmup 5 > gpurun_out/bench_r2r_noov.log 2>&1; tail -c 6000 gpurun_out/bench_r2r_noov.log | grep -o '"value": [0-9.]*, "unit": "steps/s", "n_gpus": 1\|"ms_per_step": [0-9.]*'
timeout 600 python bench.py --no-image --no-cpu --steps 45 --warmup 5 > gpurun_out/bench_r2r.log 2>&1; tail -c 6000 gpurun_out/bench_r2r.log | grep -o '"value": [0-9.]*, "unit": "steps/s", "n_gpus": 1\|"ms_per_step": [0-9.]*\|"e2e": {[^}]*}\|"graph": {[^}]*}'
