#!/usr/bin/env python
"""TEST INFRASTRUCTURE (build container only: needs /root/reference).  Long differential fuzz of the rolling-window oracle
(oracle/win_oracle.c) and of tapenv.rolling.calc_dependent against generate.InitialContainer on fresh
generate.generate_blocks instances: random dimension, total, window and policy (including inaccessible pointers).

    python scripts/long_fuzz_window.py SEED SECONDS

r01: 4 seeds x 300 s = 7 043 instances, 175 k window calls (2D and 3D, totals 2..64, windows 1..32): 0 mismatches."""
import os, sys, time, io, contextlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tap-net_b200"))
import numpy as np
from oracle import oracle, refshim
from tests.golden.make_golden import ic_adjacency
from tapenv.rolling import calc_dependent

generate = refshim.load(("tools", "generate"))["generate"]
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 120
np.random.seed(seed)
rng = np.random.RandomState(seed + 1000)
t0 = time.time(); inst = calls = 0; bad = []
while time.time() - t0 < budget and not bad:
    dim = int(rng.randint(2, 4))
    T = int(rng.randint(2, 65))
    R = 2 if dim == 2 else 6
    n = int(rng.randint(1, min(T, 32 if dim == 2 else 10) + 1))
    Wc = int(rng.randint(4, 9))
    ics = [Wc, 400] if dim == 2 else [Wc, int(rng.randint(4, 9)), 400]
    with contextlib.redirect_stdout(io.StringIO()):
        rot_blocks, positions, _, _, _ = generate.generate_blocks(T, ics, 1, [1, 5])
    blocks = np.asarray(rot_blocks).reshape(R, dim, T).transpose(0, 2, 1).reshape(R * T, dim)
    pos = np.asarray(positions).reshape(dim, T).transpose(1, 0)
    ic = generate.InitialContainer(blocks, pos, T, ics, True, n, "bot")
    adj = ic_adjacency(ic)
    if not np.array_equal(calc_dependent(blocks[:T], pos, ics).astype(np.uint8), adj):
        bad.append(("calc_dependent", dim, T, ics)); break
    oc = oracle.InitialContainer(adj, blocks, T, n, dim)
    inst += 1
    while True:
        s_ref, d_ref = ic.convert_to_input(); s, d = oc.convert_to_input(); calls += 1
        if not (np.array_equal(s_ref, s) and np.array_equal(d_ref, d) and [int(v) for v in ic.sub_graph_nodes] == oc.sub_graph_nodes
                and ic.is_last_graph() == oc.is_last_graph()):
            bad.append(("window", dim, T, n, calls)); break
        if ic.is_last_graph():
            break
        ptr = int(rng.randint(n * R))
        bid = int(ic.sub_graph_nodes[ptr % n])
        ic.remove_block(bid); oc.remove_block(bid)
print("instances", inst, "window calls", calls, "bad", bad)
