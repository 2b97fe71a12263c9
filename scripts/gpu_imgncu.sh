#!/bin/bash
# GPU box: ncu --set full captures of the image-path kernels (forward slice and backward slice of the second step)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 IMG_NCU=1 IMG_WARM=1
K='regex:^k_(conv_tc|conv_wgrad_tc|bnact)'
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base function -k "$K" --launch-skip 447 -c 9 \
   -o gpurun_out/prof_img_fwd -f python scripts/profile_image.py > gpurun_out/prof_img_fwd.log 2>&1; echo "fwd capture exit $?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base function -k "$K" --launch-skip 591 -c 14 \
   -o gpurun_out/prof_img_bwd -f python scripts/profile_image.py > gpurun_out/prof_img_bwd.log 2>&1; echo "bwd capture exit $?"
ls -la gpurun_out/*.ncu-rep
