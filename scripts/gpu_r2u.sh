#!/bin/bash
# round 2, call U: whole GPU suite + smoke + default bench + ncu launch list of the final step
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 2400 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_r2u.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_r2u.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2u.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke_r2u.log
timeout 1200 python bench.py > gpurun_out/bench_r2u.log 2>&1; echo "bench exit $?"; tail -c 1500 gpurun_out/bench_r2u.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2u_ref.log 2>&1; echo "ref arm exit $?"; tail -c 600 gpurun_out/bench_r2u_ref.log
# launch list (serialised under ncu: the side-stream overlaps are switched off so that every kernel is listed on its own)
NOLSTM='regex:^k_(gemm|split|embed|vec|head|tanh|ce_|finalize|fill|combine|time|col|dc0|reparam|sumsq|norm|clip|add_row|wait)'
LAGVAE_SIDE_WGRAD=0 LAGVAE_OVERLAP_XPROJ=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base function -k "$NOLSTM" -c 600 --csv \
    --log-file gpurun_out/launches_r2u.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-image > gpurun_out/launches_r2u_bench.log 2>&1
echo "launch list exit $?"
python scripts/lstm_trace.py > gpurun_out/trace_r2u.log 2>&1; grep -E "kernel:|next step" gpurun_out/trace_r2u.log
