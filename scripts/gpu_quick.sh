#!/bin/bash
# Quick GPU-box session: all GPU tests, C2 + C5 bench lines.  Usage: bash scripts/gpu_quick.sh TAG
TAG=${1:-rXX}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest.log
python bench.py --workload c2 --steps 100 --warmup 5 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_c2.json | cut -c1-200
python bench.py --workload c5 --steps 30 --warmup 3 --no-cpu 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_c5.json | cut -c1-200
