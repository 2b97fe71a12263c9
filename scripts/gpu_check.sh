#!/bin/bash
# GPU box: staged correctness run; every stage in its own process (a faulting kernel poisons the context).
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n 15 gpurun_out/$name.log; }
run k_simt python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "gemm_f32 or split or dropout or mi_est or clip" --timeout 120
run p_simt python -m pytest tests/test_gpu_parity.py -m gpu -q -k "simt" --timeout 300
run k_tc python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "gemm_tc" --timeout 60
run p_tc python -m pytest tests/test_gpu_parity.py -m gpu -q -k "not simt" --timeout 300
run smoke python __graft_entry__.py --smoke
run bench python bench.py --steps 5 --warmup 3
