#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for p in 0 1000 2000 3000 4000; do echo "== poll sleep ns $p"; LAGVAE_LSTM_POLL_SLEEP_NS=$p python scripts/lstm_trace.py > gpurun_out/trace_p$p.log 2>&1; grep -E "kernel:|cell done|next step" gpurun_out/trace_p$p.log; done
