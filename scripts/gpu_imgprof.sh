#!/bin/bash
# GPU box: kernel-time breakdown of the image inner step: torch.profiler table + ncu launch list (names, grids, durations)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python scripts/profile_image.py > gpurun_out/image_profile.txt 2>&1; tail -42 gpurun_out/image_profile.txt | cut -c1-160
K=$(head -3 gpurun_out/image_profile.txt | grep -o 'over [0-9]* kernels' | grep -o '[0-9]*')
echo "kernels per step: $K"
IMG_NCU=1 IMG_WARM=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip $((K + 200)) -c $((K - 150)) --csv \
   --log-file gpurun_out/image_launches.csv python scripts/profile_image.py > gpurun_out/image_ncu.log 2>&1; echo "ncu exit $?"
python scripts/summarize_launches.py gpurun_out/image_launches.csv gpurun_out/image_launches.md 2>&1 | head -45 | cut -c1-200
