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
# the reports are ~17 MB each and gpurun brings back at most 64 MiB: reduce them on the box (per-launch raw metrics + the
# per-instruction source page), keep only the CSVs
for rep in gpurun_out/${TAG}_prof_*.ncu-rep; do
  base=${rep%.ncu-rep}
  python scripts/ncu_summary.py $rep ${base}_ncu_full_summary.csv
  ncu -i $rep --page source --csv > ${base}_source.csv 2>/dev/null
  ncu -i $rep --page details --csv > ${base}_details.csv 2>/dev/null
  rm -f $rep
done
ls -la gpurun_out | grep ${TAG}
