#!/bin/bash
# A/B of library builds on the rolling-window kernel (C5): bench line value + roofline launch time per build and batch.
# Usage: ab_window.sh TAG "lib1.so lib2.so ..." "B1 B2 ..."
TAG=$1; LIBS=$2; BS=$3
OUT=gpurun_out/${TAG}_window_ab.txt; : > $OUT
for B in $BS; do for lib in $LIBS; do
  export TAPENV_LIB=$PWD/tap-net_b200/lib/$lib
  timeout 300 python bench.py --workload c5 --batch $B --steps 10 --warmup 3 --no-cpu --no-configs --no-model 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('$lib B=$B value=%.4g ms_per_step=%.4f launch_us=%.2f frac=%.3f' % (d['value'], d['ms_per_step'], r['launch_us'], r['frac']))
" | tee -a $OUT
done; done
