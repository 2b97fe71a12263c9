#!/bin/bash
# round 2, call I: new tests (decode logits, outer step, clip_adam, retire) + ncu application-replay counters of the LSTM kernels
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_image.py tests/test_gpu_kernels.py -x -q -m gpu -k "decode or decoder_update or adam or retired or lstm" > gpurun_out/pytest_r2i.log 2>&1; echo "tests exit $?"; tail -6 gpurun_out/pytest_r2i.log
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__cycles_active.avg
timeout 600 ncu --replay-mode application --clock-control none --kernel-name-base function -k regex:k_lstm_v2 --metrics $M --csv --log-file gpurun_out/lstm_ncu_app.csv python scripts/lstm_only.py > gpurun_out/lstm_ncu_app.log 2>&1; echo "ncu exit $?"; tail -3 gpurun_out/lstm_ncu_app.log; cat gpurun_out/lstm_ncu_app.csv | tail -20
