#!/bin/bash
# ncu evidence for profiles/: launch list of the default bench command + one `--set full` capture of the dominant kernel per
# workload.  Numbers printed by runs under ncu are never bench values.  Usage: bash scripts/gpu_profiles.sh TAG
TAG=${1:-rXX}
mkdir -p gpurun_out
Q="--no-cpu --no-configs --no-model"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 $Q --no-graph > gpurun_out/${TAG}_ncu_bench.log 2>&1
for spec in c2:4096 c3:4096 c4:8192 c4:1024; do
  wl=${spec%%:*}; B=${spec##*:}
  ncu --set full --clock-control none --import-source on -k regex:step_ -s 25 -c 1 -f -o gpurun_out/${TAG}_prof_step_${wl}_b${B} python bench.py --workload $wl --batch $B --steps 3 --warmup 3 $Q --no-graph > gpurun_out/${TAG}_ncu_full_${wl}_${B}.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:window_kernel -s 30 -c 1 -f -o gpurun_out/${TAG}_prof_window_c5_b8192 python bench.py --workload c5 --batch 8192 --steps 2 --warmup 3 $Q --no-graph > gpurun_out/${TAG}_ncu_full_c5.log 2>&1
# the reports are ~17 MB each and gpurun brings back at most 64 MiB: reduce them on the box, against the SAME library build --
# per-launch raw metrics, the details page, and executed warp-instructions / stall samples per source line
declare -A KERN=( [step_c2_b4096]="step_kernelILi0ELb1ELi10ELi2ELb1:4096" [step_c3_b4096]="step_kernelILi1ELb1ELi10ELi6ELb1ELi7:4096" \
                  [step_c4_b8192]="step_kernelILi2ELb1ELi20ELi2ELb0:8192" [step_c4_b1024]="step_split_kernelILi2ELi20ELi2ELi4:1024" \
                  [window_c5_b8192]="window_kernelILi1ELb1ELi10ELi6:8192" )
for rep in gpurun_out/${TAG}_prof_*.ncu-rep; do
  base=${rep%.ncu-rep}; key=${base#gpurun_out/${TAG}_prof_}
  python scripts/ncu_summary.py $rep ${base}_ncu_full_summary.csv
  ncu -i $rep --page details --csv > ${base}_details.csv 2>/dev/null
  ncu -i $rep --page source --csv > /tmp/src.csv 2>/dev/null
  spec=${KERN[$key]}
  python scripts/src_lines.py /tmp/src.csv tap-net_b200/lib/libtapenv.so "${spec%%:*}" ${spec##*:} 45 > ${base}_lines.txt 2>&1
  rm -f $rep
done
# SASS of the headline kernel (C2 fused step): the 128-bit streaming loads / stores and the warp-level reductions
cuobjdump -sass -fun '_ZN6tapenv11step_kernelILi0ELb1ELi10ELi2ELb1ELi0EEEvNS_6DevCfgENS_9StatePtrsEPKlPKfS6_S6_PfS7_S7_S7_S7_S7_' tap-net_b200/lib/libtapenv.so > /tmp/sass.txt 2>/dev/null
{ echo "# step_kernel<LBG2D, FAST, 10, 2, PLACE_FIRST> (BASELINE C2): SASS mnemonic histogram, then the listing"; grep -oE "^\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9T] )?[A-Z0-9_.]+" /tmp/sass.txt | awk '{print $NF}' | sort | uniq -c | sort -rn | head -40; echo; grep -E "^\s+/\*[0-9a-f]{4}\*/" /tmp/sass.txt | sed -E 's/\s+\/\* 0x[0-9a-f]+ \*\/$//' ; } > gpurun_out/${TAG}_step_c2_sass.txt
ls -la gpurun_out | grep ${TAG}
