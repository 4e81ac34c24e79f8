#!/bin/bash
N=$1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 300 --warmup 5 2>&1 | tail -1 | tee gpurun_out/r01z_scale_c2_n$N.json | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('N=%d'%j['n_gpus'], 'value %.4g ms/ep %.4f e2e %.4g packed %.4g'%(j['value'],j['ms_per_step'],j['e2e']['value'],j['e2e_packed']['value']))"
