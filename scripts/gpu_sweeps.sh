#!/bin/bash
# Fused step vs batch size, both launch forms.  Usage: bash scripts/gpu_sweeps.sh TAG
TAG=${1:-rXX}
mkdir -p gpurun_out
for wl in c2 c3 c4; do python scripts/sweep_step.py $wl 1024 4096 8192 65536 2>&1 | grep '^{' | tee gpurun_out/${TAG}_sweep_${wl}.jsonl; done
python scripts/probe_parts.py c4 1024 c3 1024 c2 1024 c2 4096 c3 4096 c4 8192 2>&1 | grep '^{' | tee gpurun_out/${TAG}_parts.jsonl
