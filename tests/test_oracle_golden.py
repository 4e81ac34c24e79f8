"""CPU: pin the oracle (oracle/tap_oracle.c) against the golden vectors taken from the Python reference."""
import numpy as np
import pytest

from oracle import oracle
from tests.golden_io import load_inputs, load_traj, golden_path, unpack_dynamic
from tests.rollout import oracle_rollout

KAT = np.load(golden_path("kat.npz"))

# SURVEY.md section 4 (values re-derived by make_golden.py from the live reference; asserted literally here too)
LITERAL = {
    "G1": dict(ratio=0.8775, valid=77, empty=3, final=[20, 20, 20, 9, 11],
               positions=[[0, 0], [0, 3], [3, 0], [3, 3], [0, 7], [0, 9], [0, 12], [4, 7], [4, 8], [0, 16]],
               stable=[1, 1, 1, 1, 0, 1, 1, 1, 1, 1]),
    "G2": dict(ratio=0.96, valid=49, empty=0, final=[10, 10, 10, 10, 9],
               positions=[[0, 0], [0, 1], [0, 0], [4, 0], [2, 1], [0, 4], [0, 8], [4, 4], [2, 8], [2, 9]]),
    "G3": dict(ratio=0.9415343915343916, valid=121, empty=5, final=[20, 20, 17, 18, 18, 18, 15]),
    "G4": dict(ratio=0.8405555555555555, valid=121, empty=11,
               final=[6, 6, 6, 6, 0, 3, 3, 3, 3, 0, 8, 8, 8, 5, 5, 8, 8, 8, 5, 5, 8, 8, 8, 2, 2]),
    "DRAW": dict(positions=[[0, 0], [3, 0], [3, 1]], valid=9, empty=0),
}


@pytest.mark.parametrize("name", ["G1", "G2", "G3", "G4", "DRAW"])
def test_known_answer_sequences(name):
    size = KAT[name + "_size"].tolist()
    blocks = KAT[name + "_blocks"]
    c = oracle.Container(size, len(blocks), str(KAT[name + "_reward_type"]), "diff",
                         packing_strategy=str(KAT[name + "_strategy"]))
    for t, b in enumerate(blocks):
        enc = c.add_new_block(b.astype(np.float32))
        assert np.array_equal(np.asarray(enc).reshape(-1), KAT[name + "_enc"][t]), (name, t)
        assert np.array_equal(c.heightmap.reshape(-1), KAT[name + "_heightmaps"][t]), (name, t)
    assert np.array_equal(c.positions, KAT[name + "_positions"])
    assert np.array_equal(np.array(c.stable, dtype=np.uint8), KAT[name + "_stable"])
    assert c.valid_size == int(KAT[name + "_valid"]) and c.empty_size == int(KAT[name + "_empty"])
    assert c.calc_ratio() == float(KAT[name + "_ratio"])
    lit = LITERAL[name]
    if "ratio" in lit: assert c.calc_ratio() == lit["ratio"]
    if "final" in lit: assert c.heightmap.reshape(-1).tolist() == lit["final"]
    if "positions" in lit: assert c.positions.tolist() == lit["positions"]
    if "stable" in lit: assert [int(s) for s in c.stable] == lit["stable"]
    assert c.valid_size == lit["valid"] and c.empty_size == lit["empty"]


def test_g1_diff_heightmaps_literal():
    """the returned 'diff' encodings."""
    want = [[0, 0, -3, 0], [0, 0, -6, 0], [0, 0, -3, 0], [0, 0, 1, 0], [0, 0, 0, -2], [0, 0, -3, -2], [0, 0, -7, -2],
            [0, 0, -7, -1], [0, 0, -7, 2], [0, 0, -11, 2]]
    assert KAT["G1_enc"].tolist() == want


def test_g5_mask_dynamic_known_answer():
    """SURVEY.md G5 (doc/data.md example 1, n=3, R=2)."""
    static = np.array([[[0, 1, 2, 0, 1, 2], [3, 1, 1, 2, 1, 2], [2, 1, 2, 3, 1, 1]]], np.float32)
    move = np.array([[0, 0, 0], [0, 0, 0], [0, 1, 0]], np.float32)
    left = np.array([[1, 1, 1], [0, 0, 0], [0, 0, 0]], np.float32)
    right = np.array([[0, 0, 0], [1, 1, 0], [1, 0, 1]], np.float32)
    z = np.zeros_like(move)
    dynamic = np.concatenate([np.hstack([move, move]), np.hstack([z, left]), np.hstack([z, right])])[None]
    cur = oracle.initial_mask(dynamic, 3, 2)
    assert cur.tolist() == [[1, 0, 1, 0, 0, 0]]
    mask = np.ones((1, 6), np.float32)
    want = [(2, [1, 1, 0, 0, 0, 0], [1, 1, 0, 1, 1, 0]), (4, [1, 0, 0, 1, 0, 0], [1, 0, 0, 1, 0, 0]),
            (0, [0, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0])]
    for ptr, wcur, wmask in want:
        p = np.array([ptr], np.int64)
        dynamic = oracle.update_dynamic(dynamic, static, p)
        cur, mask = oracle.update_mask(mask, dynamic, static, p)
        assert cur.tolist() == [wcur] and mask.tolist() == [wmask]


TRAJ = ["traj_2d_lbg_soft", "traj_2d_lbg_hard", "traj_2d_lbg_w7_full", "traj_2d_macs_rand", "traj_2d_macs_ppsg",
        "traj_3d_lbg_soft", "traj_3d_lbg_hard", "traj_2d_lb_soft", "traj_2d_lb_hard", "traj_3d_lb_soft",
        "traj_3d_macs_soft", "traj_3d_macs_hard"]


@pytest.mark.parametrize("name", TRAJ)
def test_oracle_replays_reference_trajectories(name):
    """Per-step equality with trajectories recorded from the live reference env path."""
    import os
    if not os.path.exists(golden_path(name + ".npz")):
        pytest.skip("fixture not generated")
    t = load_traj(name)
    static, dynamic = load_inputs(str(t["source"]), int(t["num"]))
    r = oracle_rollout(static, dynamic, t["container_size"].tolist(), str(t["reward_type"]), str(t["heightmap_type"]),
                       str(t["packing_strategy"]), ptr_seq=t["ptr"])
    B = static.shape[0]
    for k in ("cur_mask", "mask", "valid", "empty", "positions", "stable"):
        assert np.array_equal(r[k], t[k]), k
    assert np.array_equal(r["heightmap"], t["heightmap"].reshape(r["heightmap"].shape))
    assert np.array_equal(r["dec_dyn"], t["dec_dyn"])
    assert np.array_equal(r["ratio"], t["ratio"])          # bit-exact fp64
    dyn_final = unpack_dynamic(t["dynamic_final"], dynamic.shape)
    assert np.array_equal(r["dynamic"][-1], dyn_final)
    # the threaded whole-episode driver (the timed CPU baseline) agrees with the step-by-step objects
    eb = oracle.episode_batch(static, dynamic, t["ptr"], t["container_size"].tolist(), str(t["reward_type"]),
                              str(t["heightmap_type"]), str(t["packing_strategy"]), nthreads=3)
    assert eb["status"] == 0
    assert np.array_equal(eb["heightmap"], r["heightmap"][-1])
    assert np.array_equal(eb["positions"], r["positions"]) and np.array_equal(eb["stable"], r["stable"])
    assert np.array_equal(eb["reward"], r["ratio"].astype(np.float32))
    assert np.array_equal(eb["cur_mask"], r["cur_mask"][-1]) and np.array_equal(eb["mask"], r["mask"][-1])
    assert np.array_equal(eb["dynamic"], r["dynamic"][-1]) and np.array_equal(eb["dec_dyn"], r["dec_dyn"][-1])
    assert B == int(t["num"])


MUL_TRAJ = ["traj_2d_mulwith_lbg", "traj_2d_mul_macs", "traj_3d_mulwith_lbg", "traj_3d_mul_lbg_full"]


@pytest.mark.parametrize("name", MUL_TRAJ)
def test_two_container_inputs_golden(name):
    """input_type 'mul' / 'mul-with' (model.py:396-447, :499-507): the oracle's composition of two containers per
    environment reproduces the live-reference recording -- both heightmaps per step, cat(A,B) decoder input, masks, fp32 scores."""
    from tests.rollout import oracle_rollout_mul
    t = load_traj(name)
    num = int(t["num"])
    static = t["static"].astype(np.float32)
    _, dynamic = load_inputs(str(t["source"]), num)
    o = oracle_rollout_mul(static, dynamic, t["ptr"], t["container_size"].tolist(), str(t["reward_type"]),
                           str(t["heightmap_type"]), str(t["packing_strategy"]), str(t["input_type"]))
    for k in ("hm_a", "hm_b", "cur_mask", "mask", "dec_static", "positions_a", "positions_b"):
        assert np.array_equal(o[k], t[k]), k
    assert np.array_equal(o["dec_dyn"].astype(np.float32), t["dec_dyn"])
    assert np.array_equal(o["scores"], t["scores"])
    fin = np.unpackbits(t["dynamic_final"], axis=1)[:, :dynamic[0].size].reshape(dynamic.shape).astype(np.float32)
    assert np.array_equal(o["dynamic"][-1], fin)


@pytest.mark.parametrize("dim,fixture,size", [(2, "rand2d_n10.npz", [5, 50]), (3, "rand3d_n10.npz", [5, 5, 50])])
def test_g6_pretrained_network_tours(dim, fixture, size):
    """G6: tours chosen by the UNMODIFIED reference model.DRL + shipped pretrained actor (greedy), with
    the rewards DRL.forward returned; the oracle replaying those tours must reproduce the rewards (network-in-the-loop pin)."""
    z = np.load(golden_path("g6_tours.npz"))
    num = int(z["g6_%dd_num" % dim])
    static, dynamic = load_inputs(fixture, num)
    tour = z["g6_%dd_tour" % dim]
    o = oracle.episode_batch(static, dynamic, tour.T.copy(), size, "C+P+S-lb-soft", "diff", "LB_GREEDY", want=("reward", "mask"))
    assert o["status"] == 0 and float(o["mask"].sum()) == 0.0            # the network's tours visit every block exactly once
    assert np.abs(o["reward"].astype(np.float64) + z["g6_%dd_reward" % dim].astype(np.float64)).max() <= 1e-6
