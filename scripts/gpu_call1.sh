#!/bin/bash
# GPU box: image inner step (eager + CUDA graph) vs the torch port, large-batch sweep, reference-arm calibration.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python scripts/bench_image.py > gpurun_out/bench_image.log 2>&1; echo "image exit $?"; tail -n 3 gpurun_out/bench_image.log | cut -c1-1500
timeout 600 python scripts/bench_sweep.py > gpurun_out/sweep.log 2>&1; echo "sweep exit $?"; cut -c1-700 gpurun_out/sweep.log
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/ref_arm.log 2>&1; echo "ref exit $?"; tail -n 6 gpurun_out/ref_arm.log | cut -c1-1200
