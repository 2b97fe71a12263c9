#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for cs in 2 4; do for n in 2 4; do echo "== fwd cs $cs bulk $n"; LAGVAE_LSTM_FWD_CS=$cs LAGVAE_LSTM_BULK_STAGES=$n python scripts/lstm_trace.py > gpurun_out/trace_cs${cs}_b$n.log 2>&1; grep -E "forward kernel:|first stage|MMAs issued|accumulators|partials|cluster barrier|cell done|next step" gpurun_out/trace_cs${cs}_b$n.log | head -8; done; done
