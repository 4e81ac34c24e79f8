#!/usr/bin/env python
"""PPSG dataset generation (generate.generate_blocks_with_GT, BASELINE C4's generator: 20 blocks, initial container 7 wide):
the unmodified reference vs tapenv.install(pack, tools, generate) on the same seeds -- seconds per sample, identical outputs.
    python scripts/ppsg_speed.py [samples] [n]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tap-net_b200"))
import numpy as np
import tapenv
from tests import ref_model
S = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
mods = ref_model.reference_modules()
pack, tools, generate = mods["pack"], mods["tools"], mods["generate"]
heights = [10, 14, 18, 12, 16, 20, 8, 22]
def run():
    np.random.seed(4321)
    t0 = time.perf_counter()
    out = [generate.generate_blocks_with_GT(n, [7, heights[i % len(heights)]], [7, 100], 1, [1, 5], "bot", i) for i in range(S)]
    return out, time.perf_counter() - t0
tapenv.install(pack, tools, generate)
try:
    run()                                              # warm the kernels up
    ours, t_ours = run()
finally:
    tapenv.uninstall()
ref, t_ref = run()
same = all(np.array_equal(np.asarray(x), np.asarray(y)) for a, b in zip(ours, ref) for x, y in zip(a, b))
print({"samples": S, "blocks": n, "reference_s_per_sample": round(t_ref / S, 2), "tapenv_s_per_sample": round(t_ours / S, 2),
       "speedup": round(t_ref / t_ours, 1), "identical": bool(same)})
