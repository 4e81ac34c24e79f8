#!/usr/bin/env python
"""cProfile of the UNMODIFIED model.DRL.forward after tapenv.install() at the C2 batch: where does the drop-in path spend its
host time (model.py's own Python loops vs the tapenv proxy / operators)?   python scripts/profile_install.py [B]"""
import cProfile, os, pstats, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tap-net_b200"))
import torch, tapenv, bench
from tests import ref_model
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda:0")
static_h, dynamic_h, _ = bench.load_workload("c2", B, 0)
st, dy = torch.from_numpy(static_h).to(dev), torch.from_numpy(dynamic_h).to(dev)
mods = ref_model.reference_modules()
tapenv.install(mods["pack"], mods["tools"])
with torch.no_grad():
    actor = ref_model.make_actor(2, True).eval()
    ref_model.forward(actor, st, dy); torch.cuda.synchronize()
    pr = cProfile.Profile(); pr.enable()
    for _ in range(3):
        ref_model.forward(actor, st, dy)
    torch.cuda.synchronize(); pr.disable()
ps = pstats.Stats(pr); ps.sort_stats("tottime").print_stats(22)
