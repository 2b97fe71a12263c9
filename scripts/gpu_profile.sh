#!/bin/bash
# GPU box: ncu launch list of the fused inner step + full captures of the dominant GEMM kernels (1 GPU only).
# ncu's kernel replay cannot drive the cooperative thread-block-cluster LSTM kernels (k_lstm_v2: the replayed launch
# never becomes co-resident and the tool hangs), so they are excluded by name here; their evidence is the in-kernel
# clock64 trace + CUDA-event durations (scripts/lstm_trace.py -> profiles/*_lstm_trace.txt).
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
NOLSTM='regex:^k_(gemm|split|embed|vec|head|tanh|ce_|finalize|fill|combine|time|col|dc0|reparam|sumsq|norm|clip)'
# (1) every launch with its device time (cold-cache, serialised: compare SHARES)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base function -k "$NOLSTM" -c 600 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-image > gpurun_out/launches_bench.log 2>&1
echo "launch list exit $?"
# (2) full capture of the tensor-core GEMMs of one step
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base function -k regex:'^k_gemm_tc' -s 16 -c 8 \
    -o gpurun_out/prof_r2 -f python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-image > gpurun_out/prof_bench.log 2>&1
echo "full capture exit $?"
python scripts/lstm_trace.py > gpurun_out/trace.log 2>&1; tail -14 gpurun_out/trace.log
ls -la gpurun_out/
