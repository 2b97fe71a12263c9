#!/bin/bash
# call T: 3-pass dW_pred fraction on the side stream (module path / e2e)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for f in 0 0.25 0.35 0.45; do
  LAGVAE_SIDE_WGRAD_FRAC3=$f timeout 600 python bench.py --no-image --no-cpu --steps 30 --warmup 5 > gpurun_out/bench_r2t_$f.log 2>&1
  echo "frac3=$f: $(tail -c 8000 gpurun_out/bench_r2t_$f.log | grep -o '"ms_per_step": [0-9.]*\|"e2e": {"value": [0-9.]*\|"graph": {"value": [0-9.]*' | tr '\n' ' ')"
done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_benchmarked_config.py -x -q -m gpu > gpurun_out/pytest_r2t_quick.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_r2t_quick.log
