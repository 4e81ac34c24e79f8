#!/bin/bash
# Quick GPU-box session: parity tests + the two bench arms exactly as the driver runs them.  Usage: bash scripts/gpu_check.sh TAG
TAG=${1:-rXX}
mkdir -p gpurun_out
nproc > gpurun_out/${TAG}_host.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/${TAG}_host.txt; nvidia-smi -L >> gpurun_out/${TAG}_host.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/${TAG}_smoke.log
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
tail -c 1500 gpurun_out/${TAG}_bench_ref.json; tail -5 gpurun_out/${TAG}_bench_ref.err
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err
tail -c 3000 gpurun_out/${TAG}_bench_c2.json; tail -5 gpurun_out/${TAG}_bench_c2.err
