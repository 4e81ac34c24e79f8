#!/bin/bash
# Multi-GPU session (gpurun --gpus N): exchange tests + the bench under torchrun, side-stream vs in-graph exchange.
TAG=${1:-rXX}; N=${2:-2}
mkdir -p gpurun_out
nproc > gpurun_out/${TAG}_host_n${N}.txt; nvidia-smi topo -m >> gpurun_out/${TAG}_host_n${N}.txt 2>&1
timeout 600 python -m pytest tests/test_multigpu.py -q 2>&1 | tail -8 | tee gpurun_out/${TAG}_multigpu_pytest_n${N}.log
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" 2>gpurun_out/${TAG}_err.log | tail -1; }
run --steps 50 --warmup 5 > gpurun_out/${TAG}_scale_n${N}.json; cut -c1-400 gpurun_out/${TAG}_scale_n${N}.json; tail -3 gpurun_out/${TAG}_err.log
run --steps 50 --warmup 5 --inline-exchange --no-configs > gpurun_out/${TAG}_scale_n${N}_inline.json; cut -c1-300 gpurun_out/${TAG}_scale_n${N}_inline.json
run --steps 50 --warmup 5 --no-reduce --no-configs > gpurun_out/${TAG}_scale_n${N}_noreduce.json; cut -c1-300 gpurun_out/${TAG}_scale_n${N}_noreduce.json
timeout 600 python bench.py --gpus 1 --steps 50 --warmup 5 --no-configs --no-model --no-cpu 2>/dev/null | tail -1 > gpurun_out/${TAG}_scale_n1.json; cut -c1-300 gpurun_out/${TAG}_scale_n1.json
