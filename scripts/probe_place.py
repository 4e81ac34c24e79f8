#!/usr/bin/env python
"""Launch the placement kernel alone (tapenv_add_blocks) a few times for ncu: python scripts/probe_place.py c3 1024"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tap-net_b200"))
import torch, tapenv, bench
wl, B = sys.argv[1], int(sys.argv[2])
fixture, size, rt, hm, strat, desc = bench.WORKLOADS[wl]
static_h, dynamic_h, pool = bench.load_workload(wl, B, 0)
dim = len(size); R = 2 if dim == 2 else 6; S = static_h.shape[2]; n = S // R
dev = torch.device("cuda:0")
env = tapenv.BatchedContainers(size, n, rt, hm, packing_strategy=strat, batch_size=B, device=dev)
st0, dyn0 = torch.from_numpy(static_h).to(dev), torch.from_numpy(dynamic_h).to(dev)
cur, mask = env.reset(dyn0)
dyn = dyn0
g = torch.Generator(device=dev).manual_seed(1)
for t in range(n):                                   # a real episode through the UNFUSED placement: kernel n/2.. sees half-full containers
    ptr = torch.multinomial(cur, 1, generator=g).squeeze(1)
    dyn = tapenv.update_dynamic(dyn, st0, ptr, "bot", True)
    cur, mask = tapenv.update_mask(mask, dyn, st0, ptr, "bot", True)
    blocks = torch.gather(st0[:, 1:1 + dim], 2, ptr.view(-1, 1, 1).expand(-1, dim, 1)).squeeze(2).contiguous()
    env.add_new_blocks(blocks)
torch.cuda.synchronize()
print("ok", float(env.calc_ratio().mean()))
