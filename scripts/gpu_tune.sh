#!/bin/bash
# Kernel-tuning session: the copy-instruction probe (C2 floor experiment) and the fused step in both launch forms vs batch.
TAG=${1:-rXX}
mkdir -p gpurun_out
for B in 1024 4096 16384; do ./build/copy_probe $B; done 2>&1 | tee gpurun_out/${TAG}_copy_probe.jsonl
python scripts/sweep_step.py c4 512 1024 2048 4096 8192 2>&1 | grep '^{' | tee gpurun_out/${TAG}_sweep_c4.jsonl
python scripts/sweep_step.py c3 512 1024 2048 4096 8192 2>&1 | grep '^{' | tee gpurun_out/${TAG}_sweep_c3.jsonl
python scripts/sweep_step.py c2 512 1024 2048 4096 8192 2>&1 | grep '^{' | tee gpurun_out/${TAG}_sweep_c2.jsonl
