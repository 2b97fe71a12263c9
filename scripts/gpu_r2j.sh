#!/bin/bash
# strong scaling (configs[2]) at the GPU counts available in this call
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
NG=$(nvidia-smi -L | wc -l)
for n in 1 2 4 8; do
  if [ $n -le $NG ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) scripts/bench_strong.py --steps 30 --warmup 3 > gpurun_out/strong_n$n.log 2>&1; echo "strong N=$n exit $?"; grep '^{' gpurun_out/strong_n$n.log | tail -1 | cut -c1-420
  fi
done
