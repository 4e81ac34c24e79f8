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
