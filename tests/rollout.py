"""Shared test helpers: seeded random-valid policies and the CPU-oracle rollout the CUDA path is checked against."""
import numpy as np

from oracle import oracle


def dims_of(static, container_size):
    dim = len(container_size)
    R = 2 if dim == 2 else 6
    S = static.shape[2]
    return dim, R, S // R, S


def oracle_rollout(static, dynamic, container_size, reward_type, heightmap_type, packing_strategy, ptr_seq=None,
                   seed=0, steps=None):
    """Run the oracle exactly as model.py drives the reference (initial mask, then per step update_dynamic,
    update_mask, add_new_block).  ptr_seq [steps,B] is replayed if given, otherwise drawn uniformly from the
    accessible candidates with np.random.RandomState(seed) (and returned)."""
    B = static.shape[0]
    dim, R, n, S = dims_of(static, container_size)
    steps = n if steps is None else steps
    conts = [oracle.Container(container_size, n, reward_type, heightmap_type, packing_strategy=packing_strategy)
             for _ in range(B)]
    rng = np.random.RandomState(seed)
    mask = np.ones((B, S), np.float32)
    cur = oracle.initial_mask(dynamic, n, R)
    dyn = dynamic
    out = dict(ptr=[], heightmap=[], dec_dyn=[], cur_mask=[cur.copy()], mask=[], valid=[], empty=[], dynamic=[])
    for t in range(steps):
        if ptr_seq is not None:
            ptr = np.asarray(ptr_seq[t], dtype=np.int64)
        else:
            u = rng.random_sample((B, S)) * (cur > 0)
            ptr = np.argmax(u, axis=1).astype(np.int64)       # uniform over accessible candidates
        dyn = oracle.update_dynamic(dyn, static, ptr)
        cur, mask = oracle.update_mask(mask, dyn, static, ptr)
        blocks = static[np.arange(B)[:, None], 1 + np.arange(dim)[None, :], ptr[:, None]]
        enc = [np.asarray(conts[b].add_new_block(blocks[b])).reshape(-1) for b in range(B)]
        out["ptr"].append(ptr); out["dec_dyn"].append(np.stack(enc))
        out["heightmap"].append(np.stack([c.heightmap.reshape(-1) for c in conts]))
        out["cur_mask"].append(cur.copy()); out["mask"].append(mask.copy()); out["dynamic"].append(dyn)
        out["valid"].append(np.array([c.valid_size for c in conts])); out["empty"].append(np.array([c.empty_size for c in conts]))
    res = {k: np.stack(v) for k, v in out.items()}
    res["positions"] = np.stack([c.positions for c in conts])
    res["stable"] = np.stack([np.array(c.stable, dtype=np.uint8) for c in conts])
    res["ratio"] = np.array([c.calc_ratio() for c in conts], dtype=np.float64)
    res["k"] = np.array([c.current_blocks_num for c in conts])
    return res


def random_valid_ptrs(static, dynamic, container_size, seed=0):
    """A [n,B] pointer sequence of accessible candidates (independent of the container)."""
    dim, R, n, S = dims_of(static, container_size)
    B = static.shape[0]
    rng = np.random.RandomState(seed)
    mask = np.ones((B, S), np.float32)
    cur = oracle.initial_mask(dynamic, n, R)
    dyn = dynamic
    seq = []
    for t in range(n):
        u = rng.random_sample((B, S)) * (cur > 0)
        ptr = np.argmax(u, axis=1).astype(np.int64)
        dyn = oracle.update_dynamic(dyn, static, ptr)
        cur, mask = oracle.update_mask(mask, dyn, static, ptr)
        seq.append(ptr)
    return np.stack(seq)


def oracle_rollout_mul(static, dynamic, ptr_seq, container_size, reward_type, heightmap_type, packing_strategy, input_type):
    """The two-container env section of model.DRL.forward (model.py:286-292, :376-447, :499-507) on the oracle: static has
    the target-container row; per step update_dynamic, update_mask, the chosen block into container A or B, the other
    container's get_heightmap(); scores accumulated in fp32."""
    B, rows, S = static.shape
    dim = rows - 2
    R = 2 if dim == 2 else 6
    n = S // R
    mk = lambda: [oracle.Container(container_size, n, reward_type, heightmap_type, packing_strategy=packing_strategy) for _ in range(B)]
    ca, cb = mk(), mk()
    mask = np.ones((B, S), np.float32)
    cur = oracle.initial_mask(dynamic, n, R)
    dyn = dynamic
    out = dict(hm_a=[], hm_b=[], dec_dyn=[], dec_static=[], cur_mask=[cur.copy()], mask=[], dynamic=[])
    part = static[:, 1:-1] if input_type == "mul" else static[:, 1:]
    for t in range(ptr_seq.shape[0]):
        ptr = np.asarray(ptr_seq[t], dtype=np.int64)
        dyn = oracle.update_dynamic(dyn, static, ptr, input_type)
        cur, mask = oracle.update_mask(mask, dyn, static, ptr, input_type)
        tgt = static[np.arange(B), -1, ptr]
        dec_static = part[np.arange(B)[:, None], np.arange(part.shape[1])[None, :], ptr[:, None]]
        enc = []
        for b in range(B):
            blk = dec_static[b, :dim]
            if tgt[b] == 0:
                a = np.asarray(ca[b].add_new_block(blk)).reshape(-1); bb = np.asarray(cb[b].get_heightmap()).reshape(-1)
            else:
                a = np.asarray(ca[b].get_heightmap()).reshape(-1); bb = np.asarray(cb[b].add_new_block(blk)).reshape(-1)
            enc.append(np.concatenate([a, bb]))
        out["dec_dyn"].append(np.stack(enc)); out["dec_static"].append(dec_static)
        out["hm_a"].append(np.stack([c.heightmap.reshape(-1) for c in ca])); out["hm_b"].append(np.stack([c.heightmap.reshape(-1) for c in cb]))
        out["cur_mask"].append(cur.copy()); out["mask"].append(mask.copy()); out["dynamic"].append(dyn)
    res = {k: np.stack(v) for k, v in out.items()}
    s = np.zeros(B, np.float32)
    for b in range(B):                                   # model.py:503-507 on an fp32 tensor
        s[b] = np.float32(s[b] + np.float32(ca[b].calc_ratio()))
        s[b] = np.float32(s[b] + np.float32(cb[b].calc_ratio()))
        s[b] = np.float32(s[b] / np.float32(2.0))
    res["scores"] = s
    res["positions_a"] = np.stack([c.positions for c in ca]); res["positions_b"] = np.stack([c.positions for c in cb])
    return res
