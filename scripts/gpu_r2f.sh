#!/bin/bash
# call X: gate-only-for-v2 check + ncu full capture of the new streaming kernels
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_text_graph.py tests/test_gpu_benchmarked_config.py tests/test_gpu_parity.py -x -q -m gpu -k "fused or inner or step or graph" > gpurun_out/pytest_r2x_quick.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_r2x_quick.log
timeout 300 python bench.py --no-image --no-cpu --no-e2e --steps 45 --warmup 5 > gpurun_out/bench_r2x.log 2>&1; echo "bench: $(tail -c 6000 gpurun_out/bench_r2x.log | grep -o '"ms_per_step": [0-9.]*' | head -1)"
LAGVAE_SIDE_WGRAD=0 LAGVAE_OVERLAP_XPROJ=0 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base function -k regex:'^k_(ce_fused|embed_gather_split|clip_sgd|sumsq|add_row_periodic)' -s 5 -c 5 \
    -o gpurun_out/prof_r2x -f python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-image > gpurun_out/prof_r2x_bench.log 2>&1
echo "full capture exit $?"; ls -la gpurun_out/prof_r2x.ncu-rep
