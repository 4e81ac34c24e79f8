#!/bin/bash
# bench lines only (no ncu): all four workloads + the reference arm.  Usage: bash scripts/gpu_bench_all.sh TAG
TAG=${1:-rXX}
mkdir -p gpurun_out
for wl in c2 c3 c4 c5; do
  python bench.py --workload $wl --steps 200 --warmup 5 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_${wl}.json | python -c "
import json,sys; j=json.loads(sys.stdin.read()); e=j['e2e']; p=j.get('e2e_packed') or {}
print('$wl value %.4g (%.4f ms) e2e %.4g packed %s frac %.3f cpu %.4g'%(j['value'],j['ms_per_step'],e['value'],p.get('value'),j['roofline']['frac'],j['cpu_baseline']['value']))"
done
python bench.py --impl reference --steps 20 --warmup 2 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_ref.json | cut -c1-120
