#!/bin/bash
# round 2, call S: norm-only dW_pred on a side stream under the backward recurrences, gated on "recurrence grid resident"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_text_graph.py tests/test_gpu_parity.py tests/test_gpu_benchmarked_config.py -q -m gpu > gpurun_out/pytest_r2s_quick.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_r2s_quick.log
for cfg in "1 1"; do
  set -- $cfg
  LAGVAE_SIDE_WGRAD=$1 LAGVAE_SIDE_WGRAD_FRAC=$2 timeout 600 python bench.py --no-image --no-cpu --no-e2e --steps 45 --warmup 5 > gpurun_out/bench_r2s_$1_$2.log 2>&1
  echo "side=$1 frac=$2: $(tail -c 6000 gpurun_out/bench_r2s_$1_$2.log | grep -o '"ms_per_step": [0-9.]*' | head -1)"
done
