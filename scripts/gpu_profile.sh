#!/bin/bash
# GPU box: ncu launch list of the fused inner step + full captures of the dominant kernels (1 GPU only).
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
# (1) every launch with its device time (cold-cache, serialised: compare SHARES)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/launches_bench.log 2>&1
echo "launch list exit $?"
# (2) full capture: vocab GEMM, LSTM fwd, LSTM bwd (one launch each)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_gemm_tc|k_lstm_fwd_tc|k_lstm_bwd_tc' -s 12 -c 6 \
    -o gpurun_out/prof_r1 -f python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/prof_bench.log 2>&1
echo "full capture exit $?"
ls -la gpurun_out/
