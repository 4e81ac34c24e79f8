#!/bin/bash
# One GPU-box session: parity tests, smoke, bench lines and ncu captures.  Usage: bash scripts/gpu_round.sh TAG
TAG=${1:-rXX}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/${TAG}_smoke.log
for wl in c2 c3 c4 c5; do
  python bench.py --workload $wl --steps 100 --warmup 5 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_${wl}.json | cut -c1-160
done
python bench.py --impl reference --steps 20 --warmup 2 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_ref.json | cut -c1-160
# launch list of the default bench command (cold-cache, serialised per-launch times: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-graph > gpurun_out/${TAG}_ncu_bench.log 2>&1
# one full capture of the dominant kernel per workload
for wl in c2 c3 c4; do
  ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 20 -c 1 -o gpurun_out/${TAG}_prof_step_${wl} python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu --no-graph > gpurun_out/${TAG}_ncu_full_${wl}.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:window_kernel -s 30 -c 1 -o gpurun_out/${TAG}_prof_window_c5 python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu --no-graph > gpurun_out/${TAG}_ncu_full_c5.log 2>&1
ls gpurun_out | grep ${TAG}
