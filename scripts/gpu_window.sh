#!/bin/bash
# window kernel iteration: rolling parity tests, c5 bench line, ncu full capture.  Usage: bash scripts/gpu_window.sh TAG
TAG=${1:-rXX}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_rolling.py -x -q 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest_rolling.log
python bench.py --workload c5 --steps 30 --warmup 3 --no-cpu 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_c5.json | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:window_kernel -s 30 -c 2 -o gpurun_out/${TAG}_prof_window python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu --no-graph > gpurun_out/${TAG}_c5_ncu_full.log 2>&1
tail -1 gpurun_out/${TAG}_c5_ncu_full.log | cut -c1-100
