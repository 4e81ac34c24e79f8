#!/usr/bin/env python
"""torch.profiler over tapenv.DecodeLoop driving the reference's own network modules (tapenv.adapters.drl_actor_step) at the C2
batch: which GPU kernels make up the ~2 ms per decode step?   python scripts/profile_decode_loop.py [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tap-net_b200"))
import torch, tapenv, bench
from tests import ref_model
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda:0")
static_h, dynamic_h, _ = bench.load_workload("c2", B, 0)
st, dy = torch.from_numpy(static_h).to(dev), torch.from_numpy(dynamic_h).to(dev)
with torch.no_grad():
    actor = ref_model.make_actor(2, True).eval()
    env = tapenv.BatchedContainers([5, 50], 10, "C+P+S-lb-soft", "diff", batch_size=B, device=dev)
    loop = tapenv.DecodeLoop(env, tapenv.adapters.drl_actor_step(actor), greedy=True, use_graph=False)
    loop.run(st, dy); torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        loop.run(st, dy); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70))
