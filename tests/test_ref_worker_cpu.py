"""CPU: the reference-arm harness (oracle/ref_worker.py -- the UNMODIFIED Python reference's env path in worker processes)
against the C oracle under the same recorded policy: heightmaps, masks, pointers, positions bit-equal, reward <= 1e-6."""
import numpy as np
import pytest

from oracle import oracle, ref_worker, refshim
from tests.golden_io import load_inputs

pytestmark = pytest.mark.skipif(not refshim.available(), reason="reference tree not present")


def _oracle_policy_rollout(static, dynamic, u, size, rt, hm, strat):
    B, rows, S = static.shape
    dim = rows - 1
    R = 2 if dim == 2 else 6
    n = S // R
    conts = [oracle.Container(size, n, rt, hm, packing_strategy=strat) for _ in range(B)]
    mask = np.ones((B, S), np.float32)
    cur = oracle.initial_mask(dynamic, n, R)
    dyn = dynamic
    out = dict(ptr=[], heightmap=[], cur_mask=[cur.copy()])
    for t in range(n):
        ptr = ref_worker.policy_pick(cur, u[t])
        dyn = oracle.update_dynamic(dyn, static, ptr)
        cur, mask = oracle.update_mask(mask, dyn, static, ptr)
        blocks = static[np.arange(B)[:, None], 1 + np.arange(dim)[None, :], ptr[:, None]]
        for b in range(B):
            conts[b].add_new_block(blocks[b])
        out["ptr"].append(ptr); out["cur_mask"].append(cur.copy())
        out["heightmap"].append(np.stack([c.heightmap.reshape(-1) for c in conts]))
    res = {k: np.stack(v) for k, v in out.items()}
    res["positions"] = np.stack([c.positions for c in conts])
    res["reward"] = np.array([c.calc_ratio() for c in conts])
    return res


def test_policy_pick_is_uniform_over_accessible():
    cur = np.array([[0, 1, 0, 1, 1], [1, 0, 0, 0, 0]], np.float32)
    assert ref_worker.policy_pick(cur, np.array([0.0, 0.99])).tolist() == [1, 0]
    assert ref_worker.policy_pick(cur, np.array([0.34, 0.5])).tolist() == [3, 0]
    assert ref_worker.policy_pick(cur, np.array([0.99, 0.0])).tolist() == [4, 0]


@pytest.mark.parametrize("fixture,size,rt,strat", [("rand2d_n10.npz", [5, 50], "C+P+S-lb-soft", "LB_GREEDY"),
                                                   ("rand3d_n10.npz", [5, 5, 50], "C+P+S-lb-soft", "LB_GREEDY")])
def test_reference_pool_matches_oracle_under_recorded_policy(fixture, size, rt, strat):
    workers, per = 2, 12
    static, dynamic = load_inputs(fixture, workers * per)
    n = static.shape[2] // (2 if len(size) == 2 else 6)
    u = np.random.RandomState(5).random_sample((n, workers * per))
    shards, ranges = ref_worker.make_shards(static, dynamic, u, size, rt, "diff", strat, workers, per)
    pool = ref_worker.ReferencePool(workers)
    try:
        pool.load(shards)
        res = pool.run(trace=True)
        again = pool.run(trace=False)
    finally:
        pool.close()
    want = _oracle_policy_rollout(static, dynamic, u, size, rt, "diff", strat)
    for w, idx in enumerate(ranges):
        assert np.array_equal(res[w]["ptr"], want["ptr"][:, idx])
        assert np.array_equal(res[w]["heightmap"], want["heightmap"][:, idx])
        assert np.array_equal(res[w]["cur_mask"], want["cur_mask"][:, idx])
        assert np.array_equal(res[w]["positions"], want["positions"][idx])
        assert np.abs(res[w]["reward"].astype(np.float64) - want["reward"][idx]).max() <= 1e-6
        assert np.array_equal(again[w]["reward"], res[w]["reward"]) and again[w]["seconds"] > 0
