#!/bin/bash
# GPU box: fast iteration loop for the tensor-core tier (LSTM unit tests, default parity, trace, bench)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; tmo=$2; shift; shift; echo "=== $name"; timeout $tmo "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-8} gpurun_out/$name.log; return $rc; }
TAILN=25 run k_lstm 240 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "lstm_persistent or gemm_tc" --timeout 100
if [ $? -ne 0 ]; then export LAGVAE_NO_LSTM_TC=1; echo "!!! tensor-core unit tests failed -> LAGVAE_NO_LSTM_TC=1"; fi
TAILN=25 run p_tc 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "not simt" --timeout 400
TAILN=20 run trace 300 python scripts/lstm_trace.py
run bench 900 python bench.py --steps 10 --warmup 3 ${BENCH_FLAGS:---no-cpu}
