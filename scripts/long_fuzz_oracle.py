#!/usr/bin/env python
"""TEST INFRASTRUCTURE (build container only: needs /root/reference).  Long differential fuzz of the CPU oracle against the
live Python reference: random strategy (LB_GREEDY / MACS / LB) x dimension, container shape, blocks_num, reward type and
heightmap encoding; after EVERY add_new_block the returned encoding, heightmap, positions, valid / empty sizes, stability
flags (and the voxel grid for MACS / LB) must agree, and calc_ratio at the end.

    python scripts/long_fuzz_oracle.py SEED SECONDS

r01: 4 seeds x 600 s = 2.4 M steps (LB_GREEDY 2D 492 k / 3D 311 k, MACS 2D 485 k / 3D 309 k, LB 2D 488 k / 3D 309 k): 0 mismatches."""
import sys, numpy as np, io, contextlib, collections, time
sys.path.insert(0,'/root/repo')
from oracle import refshim, oracle
tools = refshim.load(("tools",))["tools"]
rng=np.random.RandomState(int(sys.argv[1]) if len(sys.argv)>1 else 0)
t0=time.time(); budget=float(sys.argv[2]) if len(sys.argv)>2 else 300
stats=collections.Counter(); bad=[]
cfgs=[]
for strat in ("LB_GREEDY","MACS","LB"):
    for dim in (2,3):
        cfgs.append((strat,dim))
while time.time()-t0 < budget and len(bad)<5:
    strat,dim=cfgs[rng.randint(len(cfgs))]
    if dim==2:
        W=int(rng.randint(2,13)); size=[W,int(rng.randint(60,200))]
    else:
        W=int(rng.randint(2,8)); L=int(rng.randint(2,8)); size=[W,L,int(rng.randint(60,200))]
    n=int(rng.randint(2,26 if dim==2 else 16))
    if strat=="MACS":
        rt=["C+P+S-mcs-soft","C+P+S-mcs-hard","C+P-mcs-soft","C+P-mcs-hard","mcs-soft","mcs-hard","C+P+S-mul-soft"][rng.randint(7)]
    else:
        rt=["C+P+S-lb-soft","C+P+S-lb-hard","C+P-lb-soft","C+P-lb-hard"][rng.randint(4)]
    hm=["full","zero","diff"][rng.randint(3)]
    mx=min(min(size[:-1]),6) if (strat=="MACS" and dim==3) else min(max(size[:-1]),6)
    maxh = size[-1]
    c=oracle.Container(size,n,rt,hm,packing_strategy=strat); r=tools.Container(size,n,rt,hm,packing_strategy=strat)
    for t in range(n):
        b=rng.randint(1,mx+1,size=dim).astype(np.float32)
        if np.asarray(r.heightmap).max() + b[-1] + 8 >= maxh: break
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                e=r.add_new_block(b.copy())
        except Exception as ex:
            stats['ref_exc_'+type(ex).__name__]+=1; break
        a=c.add_new_block(b.copy())
        stats[(strat,dim)]+=1
        ok=np.array_equal(np.asarray(a),np.asarray(e)) and np.array_equal(c.heightmap,np.asarray(r.heightmap)) and np.array_equal(c.positions,np.asarray(r.positions)) and c.valid_size==r.valid_size and c.empty_size==r.empty_size and list(c.stable)==[bool(x) for x in r.stable]
        if ok and strat!="LB_GREEDY": ok = np.array_equal(c.container,np.asarray(r.container))
        if not ok:
            bad.append((strat,dim,size,n,rt,hm,t)); print("MISMATCH",bad[-1],flush=True); break
    else:
        if abs(c.calc_ratio()-r.calc_ratio())>0 and not (np.isnan(c.calc_ratio()) and np.isnan(r.calc_ratio())):
            bad.append(("ratio",strat,dim,size,n,rt)); print("RATIO MISMATCH",bad[-1],flush=True)
print("steps",dict(stats)); print("bad",bad)
