#!/usr/bin/env python
"""Where does a fused step's time go at small batches?  Times, with CUDA-graph replays of 10 back-to-back launches on cold ring
slots: the unfused pieces (update_dynamic = the masked copy alone, update_mask, add_new_blocks = placement + state only)
and the fused step in both launch forms.
    python scripts/probe_parts.py c4 1024 [c3 1024 ...]
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tap-net_b200"))
import numpy as np, torch
import tapenv
import bench

dev = torch.device("cuda:0")
args = sys.argv[1:]
for wl, B in zip(args[0::2], [int(v) for v in args[1::2]]):
    fixture, size, rt, hm, strat, desc = bench.WORKLOADS[wl]
    static_h, dynamic_h, pool = bench.load_workload(wl, B, 0)
    dim = len(size); R = 2 if dim == 2 else 6; S = static_h.shape[2]; n = S // R
    env = tapenv.BatchedContainers(size, n, rt, hm, packing_strategy=strat, batch_size=B, device=dev)
    st0, dyn0 = torch.from_numpy(static_h).to(dev), torch.from_numpy(dynamic_h).to(dev)
    cur, mask = env.reset(dyn0)
    ptr = torch.multinomial(cur, 1).squeeze(1)
    RING = max(2, min(10, int(np.ceil(400e6 / (3 * (static_h.nbytes + dynamic_h.nbytes))))))
    nl = min(RING, n)
    sets = [(torch.roll(st0, i * 131, 0).contiguous(), torch.roll(dyn0, i * 131, 0).contiguous(), torch.roll(ptr, i * 131, 0).contiguous()) for i in range(RING)]
    blocks = [torch.gather(s[0][:, 1:1 + dim], 2, s[2].view(-1, 1, 1).expand(-1, dim, 1)).squeeze(2).contiguous() for s in sets]
    outs = [(torch.empty_like(dyn0), torch.empty(B, S, device=dev), torch.empty(B, S, device=dev), torch.empty(B, dim, device=dev), torch.empty(B, env.enc_len, device=dev)) for _ in range(RING)]
    mask1 = torch.ones(B, S, device=dev)

    def timed(fn):
        env.clear_container(); fn(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tot, cnt = 0.0, 0
        for rep in range(12):
            env.clear_container(); torch.cuda.synchronize()
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            if rep >= 2: tot += e0.elapsed_time(e1); cnt += nl
        return round(1e3 * tot / cnt, 2)

    res = {"workload": wl, "B": B}
    res["update_dynamic_us"] = timed(lambda: [tapenv.update_dynamic(sets[i][1], sets[i][0], sets[i][2], "bot", True) for i in range(nl)])
    res["update_mask_us"] = timed(lambda: [tapenv.update_mask(mask1, sets[i][1], sets[i][0], sets[i][2], "bot", True) for i in range(nl)])
    res["add_new_blocks_us"] = timed(lambda: [env.add_new_blocks(blocks[i]) for i in range(nl)])
    res["torch_copy_us"] = timed(lambda: [outs[i][0].copy_(sets[i][1]) for i in range(nl)])
    for form in ("0", "1"):
        os.environ["TAPENV_SPLIT"] = form
        res["step_%s_us" % ("cta" if form == "1" else "warp")] = timed(lambda: [env.step(sets[i][2], sets[i][0], sets[i][1], mask1, out=outs[i]) for i in range(nl)])
    os.environ.pop("TAPENV_SPLIT")
    print(json.dumps(res), flush=True)
    del env, sets, outs
