N=$1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 5 2>&1 | tail -1 | tee gpurun_out/r01m_scale_n$N.json | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print(j['n_gpus'], 'value %.4g ms/ep %.4f e2e %.3g (%.3f ms) roof frac %.3f'%(j['value'],j['ms_per_step'],j['e2e']['value'],j['e2e']['ms_per_step'],j['roofline']['frac']), j['config']['reward_reduction'][:40], j['clocks'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 30 --workload c5 2>&1 | tail -1 | tee gpurun_out/r01m_scale_c5_n$N.json | cut -c1-200
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 50 --workload c4 2>&1 | tail -1 | tee gpurun_out/r01m_scale_c4_n$N.json | cut -c1-200
