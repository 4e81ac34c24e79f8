"""Generate the committed golden fixtures from the UNMODIFIED Python reference.

TEST INFRASTRUCTURE.  Runs only in the build container (needs /root/reference);
the fixtures it writes travel to the GPU box, the reference does not.

    python tests/golden/make_golden.py [--only rand2d,rand3d,ppsg,traj,kat]

Writes next to this file:
  rand2d_n10.npz   RAND 2D n=10, 4096 samples (BASELINE configs 1/2)       inputs only
  rand3d_n10.npz   RAND 3D n=10, 2048 samples (config 3)                   inputs only
  ppsg2d_n20.npz   PPSG 2D n=20 W=7 pool (config 4)                        inputs only
  traj_*.npz       per-step trajectories of the live reference env path
                   (pack.update_dynamic + pack.update_mask + tools.Container) under a
                   seeded random-valid policy: ptr, heightmap, encoded heightmap, masks,
                   positions, stable, valid/empty, calc_ratio
  kat.npz          the known-answer vectors G1-G5 of SURVEY.md section 4 re-derived from the reference
  rolling{2,3}d_t50.npz   rolling.get_dataset(50, ...) + rolling.RollingDataset: per instance the five precedence
                   graphs of generate.InitialContainer (bit-packed adjacency) + rotation-major blocks + positions
  traj_*_mul*.npz  the two-container env section of model.DRL.forward (input_type 'mul' / 'mul-with', model.py:396-447,
                   :499-507) on the live reference: per-step heightmaps of both containers, cat(A,B) decoder input, fp32 scores
  traj_rolling_*.npz      rolling.validate's loop (rolling.py:575-640) on the live reference, network replaced by a
                   seeded random-valid policy: every window's static/dynamic/sub_graph_nodes, every step's
                   ptr / heightmap / returned encoding, final positions / stable / calc_ratio

Input tensors are stored compactly: `static` as uint8 [B,1+dim,S] (values are small
integers), `dynamic` (0/1 entries) bit-packed along the flattened [3n*S] axis.
tests/golden_io.py restores the float32 tensors.
"""
import argparse
import os
import shutil
import signal
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402


def pack_inputs(static, dynamic):
    static = np.asarray(static)
    dynamic = np.asarray(dynamic)
    assert (static == np.round(static)).all() and static.min() >= 0 and static.max() < 256
    assert np.isin(dynamic, (0.0, 1.0)).all()
    B = dynamic.shape[0]
    bits = np.packbits(dynamic.reshape(B, -1).astype(np.uint8), axis=1)
    return dict(static_u8=static.astype(np.uint8), dynamic_bits=bits, dynamic_shape=np.array(dynamic.shape))


def in_scratch(fn):
    def wrapped(*a, **k):
        cwd = os.getcwd()
        d = tempfile.mkdtemp(prefix="tapgold_")
        os.chdir(d)
        try:
            return fn(*a, **k)
        finally:
            os.chdir(cwd)
            shutil.rmtree(d, ignore_errors=True)
    return wrapped


@in_scratch
def make_rand(obj_dim, num, out_name, seed=12345):
    """pack.create_dataset(10, num, 16, obj_dim, 7, 50, 1, [1,5], seed) + PACKDataset('bot','diff',True,5)
    -- the scripts/train.sh defaults."""
    mods = refshim.load(("tools", "generate", "pack"))
    pack = mods["pack"]
    t = time.time()
    train_dir, _ = pack.create_dataset(10, num, 16, obj_dim, 7, 50, 1, [1, 5], seed=seed)
    ds = pack.PACKDataset(train_dir, 10, num, seed, "bot", "diff", True, 5)
    static, dynamic = ds.static.numpy(), ds.dynamic.numpy()
    np.savez_compressed(os.path.join(HERE, out_name), seed=seed, obj_dim=obj_dim, blocks_num=10,
                        how="pack.create_dataset(10,%d,16,%d,7,50,1,[1,5],seed=%d)" % (num, obj_dim, seed),
                        **pack_inputs(static, dynamic))
    print(out_name, static.shape, dynamic.shape, "%.1fs" % (time.time() - t))


class _Timeout(Exception):
    pass


def _alarm(signum, frame):
    raise _Timeout()


def _ppsg_worker(args):
    """One PPSG sample through generate.generate_blocks_with_GT (generate.py:977), with a
    wall-clock timeout + reseed because the generator's rejection loops are unbounded
    (SURVEY.md section 6).  DEVIATION from pack.create_dataset_gt: per-sample seeds."""
    idx, blocks_num, W, H, size_range, per_sample_timeout, prob, key = args
    mods = refshim.load(("tools", "generate"))
    generate = mods["generate"]
    import io
    import contextlib
    attempt = 0
    while True:
        seed = 12345 + idx * 1000 + attempt
        np.random.seed(seed)
        target = [W, int(np.random.choice(key, p=prob))]
        signal.signal(signal.SIGALRM, _alarm)
        signal.alarm(per_sample_timeout)
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                r = generate.generate_blocks_with_GT(blocks_num, target, [W, H], 1, size_range, "bot", idx,
                                                     allow_rot=True)
            signal.alarm(0)
            return idx, seed, [np.asarray(x) for x in r]
        except _Timeout:
            attempt += 1
        except Exception:
            signal.alarm(0)
            attempt += 1


@in_scratch
def make_ppsg(num, out_name, blocks_num=20, W=7, H=50, procs=8):
    import multiprocessing as mp
    mods = refshim.load(("tools", "generate", "pack"))
    pack, generate = mods["pack"], mods["generate"]
    t = time.time()
    # generate_height_prob (generate.py:977) reads the RAND valid set of 10-block instances
    pack.create_dataset(10, 16, 10000, 2, W, H, 1, [1, 5], seed=12345)
    prob, key = generate.generate_height_prob(2, blocks_num, [1, 5], W, W)
    with mp.Pool(procs) as pool:
        res = pool.map(_ppsg_worker, [(i, blocks_num, W, H, [1, 5], 20, prob, key) for i in range(num)], chunksize=1)
    res.sort(key=lambda r: r[0])
    d = "./data/gt_2d/pool/"
    os.makedirs(d)
    files = {k: open(d + k + ".txt", "w") for k in ("blocks", "pos", "container", "dep_move", "dep_small", "dep_large")}

    def w(f, arr):
        f.write(" ".join(str(v) for v in np.asarray(arr).reshape(-1)) + "\n")
    for _, _, (rot_blocks, positions, deps_move, small, large) in res:
        for bi in range(len(rot_blocks)):
            w(files["blocks"], rot_blocks[bi]); w(files["dep_small"], small[bi]); w(files["dep_large"], large[bi])
        w(files["pos"], positions); w(files["dep_move"], deps_move)
        w(files["container"], np.zeros(blocks_num, dtype=int))
    for f in files.values():
        f.close()
    ds = pack.PACKDataset(d, blocks_num, num, 12345, "bot", "diff", True, W)
    static, dynamic = ds.static.numpy(), ds.dynamic.numpy()
    np.savez_compressed(os.path.join(HERE, out_name), seeds=np.array([r[1] for r in res]), obj_dim=2,
                        blocks_num=blocks_num,
                        how="generate.generate_blocks_with_GT(20,[7,h~height_prob],[7,50],1,[1,5],'bot') per sample, "
                            "20 s timeout + reseed", **pack_inputs(static, dynamic))
    print(out_name, static.shape, dynamic.shape, "%.1fs" % (time.time() - t))


def ref_trajectory(static, dynamic, container_size, reward_type, heightmap_type, packing_strategy, seed,
                   input_type="bot"):
    """Drive the reference env path exactly as model.py:294-307, :376-458, :509-510 does, with
    ptr ~ multinomial(current_mask) from a seeded generator."""
    import torch
    mods = refshim.load(("tools", "pack"))
    tools, pack = mods["tools"], mods["pack"]
    static_t = torch.from_numpy(static)
    dyn = torch.from_numpy(dynamic)
    B, _, S = static.shape
    dim = len(container_size)
    R = 2 if dim == 2 else 6
    n = S // R
    g = torch.Generator().manual_seed(seed)
    containers = [tools.Container(container_size, n, reward_type, heightmap_type, packing_strategy=packing_strategy)
                  for _ in range(B)]
    mask = torch.ones(B, S)
    move = dyn[:, :n].sum(1); small = dyn[:, n:2 * n].sum(1); large = dyn[:, 2 * n:].sum(1)     # model.py:297-307
    dm = small * large + move
    cur = mask.clone(); cur[dm.ne(0)] = 0.0
    out = dict(ptr=[], heightmap=[], dec_dyn=[], cur_mask=[cur.numpy().copy()], mask=[], valid=[], empty=[])
    for t in range(n):
        ptr = torch.multinomial(cur, 1, generator=g).squeeze(1)
        dyn = pack.update_dynamic(dyn, static_t, ptr, input_type, True)
        cur, mask = pack.update_mask(mask, dyn, static_t, ptr, input_type, True)
        blocks = torch.gather(static_t[:, 1:1 + dim], 2, ptr.view(-1, 1, 1).expand(-1, dim, 1)).squeeze(2).numpy()
        hms = [np.asarray(containers[b].add_new_block(blocks[b], bool(ptr[b] < n))).copy() for b in range(B)]
        out["ptr"].append(ptr.numpy().copy())
        out["heightmap"].append(np.stack([np.asarray(c.heightmap).copy() for c in containers]))
        out["dec_dyn"].append(np.stack(hms).reshape(B, -1))
        out["cur_mask"].append(cur.numpy().copy()); out["mask"].append(mask.numpy().copy())
        out["valid"].append(np.array([c.valid_size for c in containers]))
        out["empty"].append(np.array([c.empty_size for c in containers]))
    res = {k: np.stack(v) for k, v in out.items()}
    res["positions"] = np.stack([np.asarray(c.positions) for c in containers])
    res["stable"] = np.stack([np.asarray(c.stable, dtype=np.uint8) for c in containers])
    res["ratio"] = np.array([c.calc_ratio() for c in containers], dtype=np.float64)
    res["dynamic_final"] = np.packbits(dyn.numpy().reshape(B, -1).astype(np.uint8), axis=1)
    return res


def make_traj():
    from tests.golden_io import load_inputs
    cases = [
        ("traj_2d_lbg_soft", "rand2d_n10.npz", [5, 50], "C+P+S-lb-soft", "diff", "LB_GREEDY", 96),
        ("traj_2d_lbg_hard", "rand2d_n10.npz", [5, 50], "C+P+S-lb-hard", "zero", "LB_GREEDY", 96),
        ("traj_2d_lbg_w7_full", "rand2d_n10.npz", [7, 50], "C+P-lb-soft", "full", "LB_GREEDY", 48),
        ("traj_2d_macs_rand", "rand2d_n10.npz", [5, 50], "C+P+S-mcs-soft", "diff", "MACS", 64),
        ("traj_2d_macs_ppsg", "ppsg2d_n20.npz", [7, 50], "C+P+S-mcs-hard", "diff", "MACS", 48),
        ("traj_3d_lbg_soft", "rand3d_n10.npz", [5, 5, 50], "C+P+S-lb-soft", "diff", "LB_GREEDY", 64),
        ("traj_3d_lbg_hard", "rand3d_n10.npz", [5, 5, 50], "C+P+S-lb-hard", "full", "LB_GREEDY", 48),
        ("traj_2d_lb_soft", "rand2d_n10.npz", [5, 50], "C+P+S-lb-soft", "diff", "LB", 64),
        ("traj_2d_lb_hard", "rand2d_n10.npz", [5, 50], "C+P+S-lb-hard", "full", "LB", 48),
        ("traj_3d_lb_soft", "rand3d_n10.npz", [5, 5, 50], "C+P+S-lb-soft", "diff", "LB", 48),
        ("traj_3d_macs_soft", "rand3d_n10.npz", [5, 5, 50], "C+P+S-mcs-soft", "diff", "MACS", 40),
        ("traj_3d_macs_hard", "rand3d_n10.npz", [5, 5, 50], "C+P+S-mcs-hard", "zero", "MACS", 40),
    ]
    only = os.environ.get("TRAJ_ONLY")
    if only:
        cases = [c for c in cases if only in c[0]]
    for name, src, size, rt, hm, strat, num in cases:
        p = os.path.join(HERE, src)
        if not os.path.exists(p):
            print("skip", name, "(no", src, ")")
            continue
        t = time.time()
        static, dynamic = load_inputs(p, num)
        res = ref_trajectory(static, dynamic, size, rt, hm, strat, seed=2024)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), source=src, container_size=np.array(size),
                            reward_type=rt, heightmap_type=hm, packing_strategy=strat, num=num, **res)
        print(name, "%.1fs" % (time.time() - t), "mean ratio %.4f" % res["ratio"].mean())


def ic_adjacency(ic):
    """[5,T,T] 0/1: adj[g,u,v] = edge u -> v of generate.InitialContainer's G_move/G_left/G_right/G_forward/G_backward."""
    T = ic.blocks_num
    adj = np.zeros((5, T, T), np.uint8)
    for g, G in enumerate((ic.G_move, ic.G_left, ic.G_right, ic.G_forward, ic.G_backward)):
        for u, v in G.edges():
            adj[g, int(u), int(v)] = 1
    return adj


@in_scratch
def make_rolling(obj_dim, num, out_name, T=50, n=10, seed=12345):
    """rolling.get_dataset(T, num, 4, obj_dim, 7, 250, 1, [1,5], seed) (rolling.py:20-84, the rolling.py CLI defaults)
    + rolling.RollingDataset (rolling.py:462-534) -> the InitialContainers the rolling driver consumes."""
    mods = refshim.load(("tools", "generate", "rolling"))
    rolling = mods["rolling"]
    t = time.time()
    os.makedirs("./data/rand_%dd" % obj_dim)
    train_dir, _ = rolling.get_dataset(T, num, 4, obj_dim, 7, 250, 1, [1, 5], seed=seed)
    ds = rolling.RollingDataset(train_dir, T, n, num, obj_dim, seed, "bot", "diff", True, 5, 7, 250)
    adj = np.stack([ic_adjacency(ic) for ic in ds.initial_containers])
    blocks = np.stack([np.asarray(ic.blocks) for ic in ds.initial_containers])
    positions = np.stack([np.asarray(ic.positions) for ic in ds.initial_containers])
    assert blocks.min() >= 0 and blocks.max() < 256
    np.savez_compressed(os.path.join(HERE, out_name), seed=seed, obj_dim=obj_dim, total_blocks_num=T,
                        how="rolling.get_dataset(%d,%d,4,%d,7,250,1,[1,5],seed=%d) + rolling.RollingDataset" % (T, num, obj_dim, seed),
                        adj_bits=np.packbits(adj.reshape(num, -1), axis=1), adj_shape=np.array(adj.shape),
                        blocks_u8=blocks.astype(np.uint8), positions=positions.astype(np.int16))
    print(out_name, adj.shape, blocks.shape, "%.1fs" % (time.time() - t))


def ref_rolling_trajectory(adj_src, num, container_size, n, reward_type, heightmap_type, packing_strategy, seed):
    """rolling.validate's loop (rolling.py:575-640) + the env section of rolling.DRL.forward (:325-335, :404-436)."""
    import torch
    from tests.golden_io import load_rolling
    mods = refshim.load(("tools", "generate", "pack"))
    tools, generate, pack = mods["tools"], mods["generate"], mods["pack"]
    z = load_rolling(adj_src, num)
    T, dim = z["T"], z["dim"]
    ics = [7, 250] if dim == 2 else [7, 7, 250]
    g = torch.Generator().manual_seed(seed)
    out = dict(static=[], dynamic_bits=[], nodes=[], ptr=[], heightmap=[], dec_dyn=[], positions=[], stable=[], ratio=[])
    for b in range(num):
        ic = generate.InitialContainer(z["blocks"][b], z["positions"][b], T, ics, True, n, "bot")
        assert np.array_equal(ic_adjacency(ic), z["adj"][b])
        cont = tools.Container(container_size, T, reward_type, heightmap_type, ics, packing_strategy=packing_strategy)
        rec = dict(static=[], dynamic_bits=[], nodes=[], ptr=[], heightmap=[], dec_dyn=[])
        one_step = True
        while one_step:
            static, dynamic = ic.convert_to_input()
            rec["static"].append(static.astype(np.uint8)); rec["nodes"].append(np.array(ic.sub_graph_nodes))
            rec["dynamic_bits"].append(np.packbits(dynamic.reshape(-1).astype(np.uint8)))
            if ic.is_last_graph():
                one_step = False
            st = torch.FloatTensor(static).unsqueeze(0); dyn = torch.FloatTensor(dynamic).unsqueeze(0)
            S = st.shape[2]
            mask = torch.ones(1, S)
            dm = dyn[:, n:2 * n].sum(1) * dyn[:, 2 * n:].sum(1) + dyn[:, :n].sum(1)
            cur = mask.clone(); cur[dm.ne(0)] = 0.0
            for _ in range(1 if one_step else n):
                ptr = torch.multinomial(cur, 1, generator=g).squeeze(1)
                dyn = pack.update_dynamic(dyn, st, ptr, "bot", True)
                cur, mask = pack.update_mask(mask, dyn, st, ptr, "bot", True)
                blk = st[0, 1:, int(ptr)].numpy()
                enc = np.asarray(cont.add_new_block(blk, bool(ptr < n))).reshape(-1).copy()
                rec["ptr"].append(int(ptr)); rec["heightmap"].append(np.asarray(cont.heightmap).reshape(-1).copy())
                rec["dec_dyn"].append(enc)
            p = int(ptr)
            while p >= n:
                p -= n
            ic.remove_block(ic.sub_graph_nodes[p])
        for k in rec:
            out[k].append(np.stack(rec[k]))
        out["positions"].append(np.asarray(cont.positions)); out["stable"].append(np.asarray(cont.stable, dtype=np.uint8))
        out["ratio"].append(cont.calc_ratio())
    return {k: np.stack(v) for k, v in out.items()}


def make_rolling_traj():
    cases = [("traj_rolling_3d", "rolling3d_t50.npz", 16, [5, 5, 250], "C+P+S-lb-soft", "diff", "LB_GREEDY"),
             ("traj_rolling_2d", "rolling2d_t50.npz", 24, [5, 250], "C+P+S-lb-hard", "diff", "LB_GREEDY")]
    for name, src, num, size, rt, hm, strat in cases:
        t = time.time()
        res = ref_rolling_trajectory(src, num, size, 10, rt, hm, strat, seed=2025)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), source=src, container_size=np.array(size), window=10,
                            reward_type=rt, heightmap_type=hm, packing_strategy=strat, num=num, **res)
        print(name, "%.1fs" % (time.time() - t), "mean ratio %.4f" % res["ratio"].mean())


def ref_trajectory_mul(static, dynamic, container_size, reward_type, heightmap_type, packing_strategy, seed, input_type):
    """The two-container env section of model.DRL.forward (model.py:286-292, :376-447, :499-507) on the live reference:
    static carries the target-container row, the chosen block goes into containers_a / containers_b."""
    import torch
    mods = refshim.load(("tools", "pack"))
    tools, pack = mods["tools"], mods["pack"]
    static_t = torch.from_numpy(static)
    dyn = torch.from_numpy(dynamic)
    B, rows, S = static.shape
    dim = rows - 2
    R = 2 if dim == 2 else 6
    n = S // R
    g = torch.Generator().manual_seed(seed)
    mk = lambda: [tools.Container(container_size, n, reward_type, heightmap_type, packing_strategy=packing_strategy) for _ in range(B)]
    ca, cb = mk(), mk()
    mask = torch.ones(B, S)
    dm = dyn[:, n:2 * n].sum(1) * dyn[:, 2 * n:].sum(1) + dyn[:, :n].sum(1)
    cur = mask.clone(); cur[dm.ne(0)] = 0.0
    out = dict(ptr=[], hm_a=[], hm_b=[], dec_dyn=[], dec_static=[], cur_mask=[cur.numpy().copy()], mask=[])
    static_part = static_t[:, 1:-1, :] if input_type == "mul" else static_t[:, 1:, :]                 # model.py:388-394
    ssz = static_part.shape[1]
    for t in range(n):
        ptr = torch.multinomial(cur, 1, generator=g).squeeze(1)
        dyn = pack.update_dynamic(dyn, static_t, ptr, input_type, True)
        cur, mask = pack.update_mask(mask, dyn, static_t, ptr, input_type, True)
        target_ids = torch.gather(static_t[:, -1, :], 1, ptr.view(-1, 1))                               # model.py:396-401
        decoder_static = torch.gather(static_part, 2, ptr.view(-1, 1, 1).expand(-1, ssz, 1))
        blocks = decoder_static.transpose(2, 1).squeeze(1).numpy()[:, :dim]
        ha, hb = [], []
        for b in range(B):
            if target_ids[b] == 0:
                ha.append(np.asarray(ca[b].add_new_block(blocks[b], bool(ptr[b] < n))).copy()); hb.append(np.asarray(cb[b].get_heightmap()).copy())
            elif target_ids[b] == 1:
                ha.append(np.asarray(ca[b].get_heightmap()).copy()); hb.append(np.asarray(cb[b].add_new_block(blocks[b], bool(ptr[b] < n))).copy())
        ha, hb = torch.FloatTensor(np.array(ha)), torch.FloatTensor(np.array(hb))
        if dim == 2:
            dd = torch.cat((ha.unsqueeze(2), hb.unsqueeze(2)), 1)                                       # model.py:424-430
        else:
            if heightmap_type != "diff":
                ha, hb = ha.unsqueeze(1), hb.unsqueeze(1)
            dd = torch.cat((ha, hb), 1)                                                                # model.py:431-441
        out["ptr"].append(ptr.numpy().copy()); out["dec_dyn"].append(dd.numpy().reshape(B, -1).copy())
        out["dec_static"].append(decoder_static.squeeze(2).numpy().copy())
        out["hm_a"].append(np.stack([np.asarray(c.heightmap).reshape(-1).copy() for c in ca]))
        out["hm_b"].append(np.stack([np.asarray(c.heightmap).reshape(-1).copy() for c in cb]))
        out["cur_mask"].append(cur.numpy().copy()); out["mask"].append(mask.numpy().copy())
    scores = torch.zeros(B)                                                                            # model.py:499-507
    for b in range(B):
        scores[b] += ca[b].calc_ratio()
        scores[b] += cb[b].calc_ratio()
        scores[b] /= 2.0
    res = {k: np.stack(v) for k, v in out.items()}
    res["scores"] = scores.numpy().copy()
    res["positions_a"] = np.stack([np.asarray(c.positions) for c in ca]); res["positions_b"] = np.stack([np.asarray(c.positions) for c in cb])
    res["static"] = static.astype(np.uint8)
    res["dynamic_final"] = np.packbits(dyn.numpy().reshape(B, -1).astype(np.uint8), axis=1)
    return res


def make_traj_mul():
    from tests.golden_io import load_inputs
    cases = [("traj_2d_mulwith_lbg", "rand2d_n10.npz", [5, 50], "C+P+S-lb-soft", "diff", "LB_GREEDY", 64, "mul-with"),
             ("traj_2d_mul_macs", "rand2d_n10.npz", [5, 50], "C+P+S-mul-hard", "zero", "MUL", 48, "mul"),
             ("traj_3d_mulwith_lbg", "rand3d_n10.npz", [5, 5, 50], "C+P+S-lb-hard", "diff", "LB_GREEDY", 32, "mul-with"),
             ("traj_3d_mul_lbg_full", "rand3d_n10.npz", [5, 5, 50], "C+P+S-lb-soft", "full", "LB_GREEDY", 24, "mul")]
    for name, src, size, rt, hm, strat, num, it in cases:
        t = time.time()
        static, dynamic = load_inputs(os.path.join(HERE, src), num)
        B, rows, S = static.shape
        R = 2 if len(size) == 2 else 6
        n = S // R
        ids = np.random.RandomState(31).randint(0, 2, size=(B, 1, n)).astype(np.float32)             # container.txt row, pack.py:212-216
        static = np.concatenate([static, np.tile(ids, (1, 1, R))], 1)
        res = ref_trajectory_mul(static, dynamic, size, rt, hm, strat, seed=77, input_type=it)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), source=src, container_size=np.array(size), reward_type=rt,
                            heightmap_type=hm, packing_strategy=strat, num=num, input_type=it, **res)
        print(name, "%.1fs" % (time.time() - t), "mean score %.4f" % res["scores"].mean())


def make_reward_kat():
    """pack.reward (pack.py:378-473, the deprecated re-packing reward) on the live reference: single-container 'bot' inputs and
    the two-container 'mul-with' split, for seeded random tours."""
    import torch
    from tests.golden_io import load_inputs
    pack = refshim.load(("tools", "pack"))["pack"]
    out = {}
    rng = np.random.RandomState(17)
    for name, src, num, it, rt, W, H in [("bot2d", "rand2d_n10.npz", 48, "bot", "C+P+S-lb-soft", 5, 50),
                                         ("mul2d", "rand2d_n10.npz", 48, "mul-with", "C+P+S-lb-hard", 5, 50),
                                         ("mul3d", "rand3d_n10.npz", 24, "mul", "C+P+S-lb-soft", 5, 50)]:
        static, _ = load_inputs(os.path.join(HERE, src), num)
        B, rows, S = static.shape
        dim = rows - 1
        R = 2 if dim == 2 else 6
        n = S // R
        if it != "bot":
            ids = rng.randint(0, 2, size=(B, 1, n)).astype(np.float32)
            ids[0] = 0                                                  # one environment with an empty container B
            static = np.concatenate([static, np.tile(ids, (1, 1, R))], 1)
        tour = np.stack([rng.permutation(n) + n * rng.randint(0, R, size=n) for _ in range(B)]).astype(np.int64)
        r = pack.reward(torch.from_numpy(static), torch.from_numpy(tour), rt, it, True, W, H)
        out[name + "_static"] = static.astype(np.uint8); out[name + "_tour"] = tour; out[name + "_reward"] = r.numpy()
        out[name + "_args"] = np.array([rt, it, str(W), str(H)])
    np.savez_compressed(os.path.join(HERE, "reward_kat.npz"), **out)
    print("reward_kat.npz", {k: float(v.mean()) for k, v in out.items() if k.endswith("_reward")})


def make_g6():
    """G6 of SURVEY.md section 4: the UNMODIFIED reference model.DRL with the shipped pretrained actors
    (pretrain_model/{2d,3d}-bot-C+P+S-lb-soft-width-5-note-sh-R-diff/actor.pt), eval() = greedy decode, on fixture inputs;
    the environment inside is the reference's tools.Container.  Recorded: the tours the network chose and the rewards
    DRL.forward returned (-calc_ratio).  A network-in-the-loop pin of the env path: replaying the tours must reproduce the
    rewards (the tours themselves are argmax outputs of an fp32 CPU network and are data here, not something to re-derive)."""
    import io
    import contextlib
    import torch
    from tests.golden_io import load_inputs
    mods = refshim.load(("tools", "generate", "pack", "model"))
    pack, model = mods["pack"], mods["model"]
    out = {}
    for dim, src, num, ck in ((2, "rand2d_n10.npz", 64, "2d-bot-C+P+S-lb-soft-width-5-note-sh-R-diff"),
                              (3, "rand3d_n10.npz", 32, "3d-bot-C+P+S-lb-soft-width-5-note-sh-R-diff")):
        static, dynamic = load_inputs(os.path.join(HERE, src), num)
        with contextlib.redirect_stdout(io.StringIO()):
            actor = model.DRL(dim, 30, 128, 256, False, "bot", True, 5, 50, dim, "C+P+S-lb-soft", "shape_heightmap", "diff",
                              "LB_GREEDY", pack.update_dynamic, pack.update_mask, 1, 0.1, 1.0)
        sd = torch.load(os.path.join(refshim.REFERENCE_DIR, "pretrain_model", ck, "actor.pt"), map_location="cpu")
        missing = actor.load_state_dict(sd)
        actor.eval()
        st, dy = torch.from_numpy(static), torch.from_numpy(dynamic)
        dec_static = torch.zeros(num, dim, 1)
        dec_dyn = torch.zeros(num, 4, 1) if dim == 2 else torch.zeros(num, 2, 5, 5)
        with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
            tour_idx, tour_logp, _, R = actor(st, dy, [dec_static, dec_dyn])
        out["g6_%dd_tour" % dim] = tour_idx.numpy().astype(np.int64)
        out["g6_%dd_reward" % dim] = R.numpy().astype(np.float32)
        out["g6_%dd_num" % dim] = num
        print("G6 %dD:" % dim, str(missing), "first tour", tour_idx[0].tolist(), "mean reward %.5f" % float(R.mean()))
    np.savez_compressed(os.path.join(HERE, "g6_tours.npz"), **out)


def make_kat():
    """Known-answer vectors (G1-G4 sequences, doc/data.md, visual/draw_result.py),
    outputs re-derived here from the live reference."""
    tools = refshim.load(("tools",))["tools"]
    seqs = {
        "G1": ([5, 50], "C+P+S-lb-soft", "LB_GREEDY",
               [[3, 3], [3, 3], [2, 3], [2, 4], [4, 2], [3, 3], [3, 4], [1, 1], [1, 3], [3, 4]]),
        "G2": ([5, 50], "C+P+S-lb-hard", "LB_GREEDY",
               [[4, 1], [2, 3], [4, 2], [1, 1], [3, 3], [4, 4], [2, 2], [1, 4], [3, 1], [2, 1]]),
        "G3": ([7, 100], "C+P+S-mcs-hard", "MACS",
               [[3, 2], [1, 4], [4, 3], [2, 1], [3, 2], [1, 2], [1, 4], [2, 3], [2, 3], [3, 2], [4, 3], [4, 4], [3, 4],
                [1, 2], [1, 2], [2, 2], [3, 2], [2, 3], [1, 3], [1, 4]]),
        "G4": ([5, 5, 50], "C+P+S-lb-soft", "LB_GREEDY",
               [[2, 4, 3], [3, 2, 2], [1, 4, 3], [3, 1, 4], [3, 2, 2], [2, 2, 3], [2, 2, 3], [1, 2, 2], [2, 3, 2],
                [3, 3, 1]]),
        "DRAW": ([4, 6], "C+P+S-lb-hard", "LB_GREEDY", [[3, 2], [1, 1], [1, 2]]),   # visual/draw_result.py:14-34
    }
    out = {}
    for name, (size, rt, strat, blocks) in seqs.items():
        c = tools.Container(size, len(blocks), rt, "diff", packing_strategy=strat)
        hms, encs = [], []
        for b in blocks:
            encs.append(np.asarray(c.add_new_block(np.array(b, dtype=np.float32))).reshape(-1).copy())
            hms.append(np.asarray(c.heightmap).reshape(-1).copy())
        out[name + "_size"] = np.array(size); out[name + "_reward_type"] = rt; out[name + "_strategy"] = strat
        out[name + "_blocks"] = np.array(blocks); out[name + "_heightmaps"] = np.stack(hms)
        out[name + "_enc"] = np.stack(encs); out[name + "_positions"] = np.asarray(c.positions)
        out[name + "_stable"] = np.asarray(c.stable, dtype=np.uint8)
        out[name + "_valid"] = c.valid_size; out[name + "_empty"] = c.empty_size; out[name + "_ratio"] = c.calc_ratio()
    np.savez_compressed(os.path.join(HERE, "kat.npz"), **out)
    print("kat.npz", {k: float(out[k + "_ratio"]) for k in seqs})


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="rand2d,rand3d,ppsg,traj,kat,rolling,rolltraj,mul,reward,g6")
    ap.add_argument("--ppsg-num", type=int, default=512)
    a = ap.parse_args()
    only = a.only.split(",")
    if "kat" in only: make_kat()
    if "rand2d" in only: make_rand(2, 4096, "rand2d_n10.npz")
    if "rand3d" in only: make_rand(3, 2048, "rand3d_n10.npz")
    if "ppsg" in only: make_ppsg(a.ppsg_num, "ppsg2d_n20.npz")
    if "traj" in only: make_traj()
    if "rolling" in only:
        make_rolling(3, 512, "rolling3d_t50.npz")
        make_rolling(2, 256, "rolling2d_t50.npz")
    if "rolltraj" in only: make_rolling_traj()
    if "mul" in only: make_traj_mul()
    if "reward" in only: make_reward_kat()
    if "g6" in only: make_g6()
