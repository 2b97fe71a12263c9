#!/bin/bash
# GPU box: staged correctness run; every stage in its own process (a faulting kernel poisons the context).
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; tmo=$2; shift; shift; echo "=== $name"; timeout $tmo "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-12} gpurun_out/$name.log; return $rc; }
run k_simt 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "gemm_f32 or split or dropout or mi_est or clip" --timeout 120
run k_tc 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "gemm_tc" --timeout 60
run p_simt 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "simt" --timeout 300
TAILN=30 run k_lstm 240 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "lstm_persistent" --timeout 100
if [ $? -ne 0 ]; then export LAGVAE_NO_LSTM_TC=1; echo "!!! persistent LSTM failed -> LAGVAE_NO_LSTM_TC=1 for the remaining stages"; fi
TAILN=30 run p_tc 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "not simt" --timeout 400
run smoke 300 python __graft_entry__.py --smoke
run bench 900 python bench.py --steps 10 --warmup 3
