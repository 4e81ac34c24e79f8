#!/bin/bash
# ncu --set full capture of the voxel-state placements alone (tapenv_add_blocks through scripts/probe_place.py: step 6 of a
# 10-step episode), reduced on the box: per-launch summary + executed warp-instructions / stall samples per source line.
# Usage: bash scripts/gpu_prof_voxel.sh TAG [B]
TAG=${1:-rXX}; B=${2:-1024}
mkdir -p gpurun_out
declare -A KERN=( [c3macs]="add_blocks_kernelILi4E" [c3lb]="add_blocks_kernelILi3E" )
for wl in c3macs c3lb; do
  rep=gpurun_out/${TAG}_place_${wl}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:add_blocks_kernel -s 6 -c 1 -f -o $rep python scripts/probe_place.py $wl $B > ${rep}_ncu.log 2>&1
  python scripts/ncu_summary.py $rep.ncu-rep ${rep}_summary.csv
  ncu -i $rep.ncu-rep --page source --csv > /tmp/src.csv 2>/dev/null
  python scripts/src_lines.py /tmp/src.csv tap-net_b200/lib/libtapenv.so "${KERN[$wl]}" $B 60 > ${rep}_lines.txt 2>&1
  rm -f $rep.ncu-rep
done
ls -la gpurun_out | grep ${TAG}
