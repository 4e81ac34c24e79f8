#!/bin/bash
# compute-sanitizer over the kernels that changed in round 2 (memcheck: all of them; racecheck: the CTA-per-environment step
# with its shared-memory hand-off, the warp-synchronous window kernel, the lane-0 voxel walk).  Usage: bash scripts/gpu_sanitize.sh TAG
TAG=${1:-rXX}
mkdir -p gpurun_out
SEL='fused_step_matches_oracle or macs_long or whole_episode_kernel or two_container_inputs_voxel or ragged or unplaceable'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "$SEL" 2>&1 | tail -6 | tee gpurun_out/${TAG}_memcheck.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_rolling.py -q -x -k "rolling_batch_vs_oracle or reference_recorded" 2>&1 | tail -4 | tee -a gpurun_out/${TAG}_memcheck.log
TAPENV_SPLIT=1 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "ragged or unplaceable or known_answer" 2>&1 | tail -4 | tee gpurun_out/${TAG}_racecheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "two_container_inputs_voxel" 2>&1 | tail -4 | tee -a gpurun_out/${TAG}_racecheck.log
