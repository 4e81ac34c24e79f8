#!/bin/bash
# ncu --set full captures of the fused step at the small per-GPU batches of the 8-GPU configurations (C4: 1024, C3/C2: 1024),
# CTA-per-environment form (placement warp isolated) and warp-per-environment form.
TAG=${1:-rXX}
mkdir -p gpurun_out
for wl in c4 c3; do
  SWEEP_FORMS=1 ncu --set full --clock-control none --import-source on -k regex:step_split_kernel -s 6 -c 1 -f -o gpurun_out/${TAG}_split_${wl}_b1024 python scripts/sweep_step.py $wl 1024 > gpurun_out/${TAG}_ncu_${wl}.log 2>&1
  SWEEP_FORMS=0 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 6 -c 1 -f -o gpurun_out/${TAG}_warp_${wl}_b1024 python scripts/sweep_step.py $wl 1024 >> gpurun_out/${TAG}_ncu_${wl}.log 2>&1
done
ls -la gpurun_out | grep ${TAG}
