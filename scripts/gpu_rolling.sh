#!/bin/bash
# GPU-box session for the rolling workload (c5): parity tests, bench line, ncu launch list + full capture of window_kernel.
TAG=${1:-rXX}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/${TAG}_pytest.log
python bench.py --workload c5 --steps 30 --warmup 3 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_c5.json | cut -c1-300
python bench.py --workload c5 --impl reference --steps 5 --warmup 1 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_c5_ref.json | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_c5_launches.csv python bench.py --workload c5 --steps 1 --warmup 3 --no-cpu --no-graph > gpurun_out/${TAG}_c5_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:window_kernel -s 30 -c 2 -o gpurun_out/${TAG}_prof_window python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu --no-graph > gpurun_out/${TAG}_c5_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_c5_ncu_full.log | cut -c1-200
