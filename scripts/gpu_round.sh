#!/bin/bash
# One GPU-box session: parity tests, smoke, bench lines and ncu captures.  Usage: bash scripts/gpu_round.sh TAG
TAG=${1:-rXX}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for wl in c2 c3 c4; do
  python bench.py --workload $wl --steps 100 --warmup 5 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_${wl}.json
done
python bench.py --impl reference --steps 20 --warmup 2 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-graph > gpurun_out/${TAG}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 20 -c 2 -o gpurun_out/${TAG}_prof_step python bench.py --steps 3 --warmup 3 --no-cpu --no-graph > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_full.log
