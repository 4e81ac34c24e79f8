#!/bin/bash
# step-kernel register-cap variants for the 3D / MACS placements
for v in "" _s6 _s7 _s8; do
  if [ -n "$v" ]; then export TAPENV_LIB=$PWD/tap-net_b200/lib/libtapenv$v.so; else unset TAPENV_LIB; fi
  for w in c3 c4; do python scripts/sweep_step.py $w 1024 4096 16384 2>&1 | grep workload | python -c "
import sys,json
for l in sys.stdin:
    j=json.loads(l); print('variant[$v]', j['workload'], j['B'], j['launch_us'], j['frac_of_6540'])"; done
done
