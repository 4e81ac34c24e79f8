#!/bin/bash
# window-kernel occupancy variants + racecheck of the warp-synchronous shared-memory code
mkdir -p gpurun_out
for v in "" _w10 _w12; do
  if [ -n "$v" ]; then export TAPENV_LIB=$PWD/tap-net_b200/lib/libtapenv$v.so; else unset TAPENV_LIB; fi
  python bench.py --workload c5 --steps 20 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('variant[$v]', 'value %.4g'%j['value'], 'launch_us %.2f'%j['roofline']['launch_us'])"
done
unset TAPENV_LIB
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 10 python -m pytest tests/test_gpu_rolling.py -q -x -k "trajectory or shapes" 2>&1 | tail -8 | tee gpurun_out/r01n_racecheck.log
python bench.py --steps 100 --warmup 5 2>&1 | tail -1 | tee gpurun_out/r01n_bench_c2.json | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('c2', j['value'], j['roofline'])"
