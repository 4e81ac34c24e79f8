#!/bin/bash
# cost of the end-of-episode reward reduction at N GPUs: fused peer exchange vs NCCL all-gather vs none
N=$1
for flag in "" "--nccl-reduce" "--no-reduce"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 300 --warmup 5 --no-cpu $flag 2>&1 | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('N=%d [$flag]'%j['n_gpus'], 'value %.4g ms/ep %.4f'%(j['value'],j['ms_per_step']), j['config']['reward_reduction'][:50])"
done
