#!/usr/bin/env python
"""Time the fused step kernel alone (CUDA-graph replays over a ring of cold batches) for several batch sizes, in both launch
forms (TAPENV_SPLIT=0: one warp per environment, =1: one CTA per environment).
    python scripts/sweep_step.py [c2|c3|c4] [B ...]
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tap-net_b200"))
import numpy as np, torch
import tapenv
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
sizes = [int(v) for v in sys.argv[2:]] or [1024, 4096, 16384, 65536]
forms = os.environ.get("SWEEP_FORMS", "0,1").split(",")
dev = torch.device("cuda:0")
fixture, size, rt, hm, strat, desc = bench.WORKLOADS[wl]
for B in sizes:
    static_h, dynamic_h, pool = bench.load_workload(wl, B, 0)
    dim = len(size); R = 2 if dim == 2 else 6; S = static_h.shape[2]; n = S // R
    bytes_step = bench.algorithmic_bytes_per_env_step(n, R, dim, size[0], size[1] if dim == 3 else 1, strat == "MACS" or "mcs" in rt)
    per_set = static_h.nbytes + dynamic_h.nbytes
    RING = max(2, min(10, int(np.ceil(400e6 / (3 * per_set)))))
    env = tapenv.BatchedContainers(size, n, rt, hm, packing_strategy=strat, batch_size=B, device=dev)
    st0, dyn0 = torch.from_numpy(static_h).to(dev), torch.from_numpy(dynamic_h).to(dev)
    cur, mask = env.reset(dyn0)
    ptr = torch.multinomial(cur, 1).squeeze(1)
    sets = [(torch.roll(st0, i * 131, 0).contiguous(), torch.roll(dyn0, i * 131, 0).contiguous(), torch.roll(ptr, i * 131, 0).contiguous()) for i in range(RING)]
    outs = [(torch.empty_like(dyn0), torch.empty(B, S, device=dev), torch.empty(B, S, device=dev), torch.empty(B, dim, device=dev), torch.empty(B, env.enc_len, device=dev)) for _ in range(RING)]
    mask1 = torch.ones(B, S, device=dev)
    nl = min(RING, n)
    for form in forms:
        os.environ["TAPENV_SPLIT"] = form
        env.clear_container()
        for i in range(nl): env.step(sets[i][2], sets[i][0], sets[i][1], mask1, out=outs[i])
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(nl): env.step(sets[i][2], sets[i][0], sets[i][1], mask1, out=outs[i])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tot = 0.0; cnt = 0
        for rep in range(12):
            env.clear_container(); torch.cuda.synchronize()
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            if rep >= 2: tot += e0.elapsed_time(e1); cnt += nl
        us = 1e3 * tot / cnt
        gbs = B * bytes_step / (us * 1e-6) / 1e9
        print(json.dumps({"workload": wl, "B": B, "form": "cta" if form == "1" else "warp", "launch_us": round(us, 2), "GBps": round(gbs, 1),
                          "frac_of_6538": round(gbs / 6538.3, 3), "env_steps_per_s": B / (us * 1e-6), "lib": os.environ.get("TAPENV_LIB", "default")}), flush=True)
    del env, sets, outs
    torch.cuda.empty_cache()
