#!/bin/bash
# A/B of two library builds on the fused step: launch time (graph replays) + ncu instruction counts.  Usage: ab_ncu.sh TAG WL B
TAG=$1; WL=$2; B=$3
for lib in libtapenv_old.so libtapenv.so; do
  export TAPENV_LIB=$PWD/tap-net_b200/lib/$lib
  SWEEP_FORMS=0 python scripts/sweep_step.py $WL $B 2>&1 | grep '^{'
  SWEEP_FORMS=0 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:step_kernel -s 6 -c 1 python scripts/sweep_step.py $WL $B 2>&1 | grep -E "step_kernel|gpu__time|inst_executed|registers|warps_active|issue_active|dram__bytes"
done
