"""CPU: pin the rolling-window oracle (oracle/win_oracle.c: generate.InitialContainer, generate.py:1589-1825) against
the trajectories recorded from the live reference (tests/golden/traj_rolling_*.npz), against the running interpreter's
own `set` (the CPython iteration order the reference inherits through networkx) and -- in the build container -- against
generate.InitialContainer itself."""
import numpy as np
import pytest

from oracle import oracle, refshim
from tests.golden_io import load_rolling, load_traj


def test_pyset_order_matches_this_interpreter():
    rng = np.random.RandomState(0)
    for trial in range(20000):
        k = int(rng.randint(1, 40))
        hi = (64, 50, 200, 17)[trial % 4]
        keys = [int(x) for x in rng.choice(hi, size=min(k, hi), replace=False)]
        assert oracle.pyset_order(keys) == list(set(keys)), keys
    # the window shapes the rolling driver produces: 9 sorted survivors + 1 new node, total 50
    for trial in range(5000):
        keys = sorted(int(x) for x in rng.choice(50, size=9, replace=False))
        new = int(rng.choice([v for v in range(50) if v not in keys]))
        assert oracle.pyset_order(keys + [new]) == list(set(keys + [new]))


def replay_oracle_windows(traj, data, order=oracle.WINDOW_ORDER_REFERENCE):
    """Drive oracle.InitialContainer with the recorded pointers; yields (b, call, static, dynamic, nodes)."""
    n = int(traj["window"])
    T, dim = data["T"], data["dim"]
    for b in range(int(traj["num"])):
        ic = oracle.InitialContainer(data["adj"][b], data["blocks"][b], T, n, dim, order=order)
        ptrs = traj["ptr"][b]
        calls = traj["static"][b].shape[0]
        t = 0
        for c in range(calls):
            static, dynamic = ic.convert_to_input()
            yield b, c, static, dynamic, ic.sub_graph_nodes
            last = ic.is_last_graph()
            assert last == (c == calls - 1)
            t += n if last else 1
            ic.remove_block(ic.sub_graph_nodes[int(ptrs[t - 1]) % n])
        assert t == T and ic.error == 0


@pytest.mark.parametrize("name", ["traj_rolling_3d", "traj_rolling_2d"])
def test_window_sequence_golden(name):
    traj = load_traj(name)
    data = load_rolling(str(traj["source"]), int(traj["num"]))
    n = int(traj["window"])
    S = traj["static"].shape[-1]
    differs = 0
    for b, c, static, dynamic, nodes in replay_oracle_windows(traj, data):
        assert np.array_equal(static, traj["static"][b, c].astype(np.float32)), (b, c)
        ref_dyn = np.unpackbits(traj["dynamic_bits"][b, c])[:3 * n * S].reshape(3 * n, S).astype(np.float32)
        assert np.array_equal(dynamic, ref_dyn), (b, c)
        assert nodes == traj["nodes"][b, c].tolist()
        differs += oracle.pyset_order(nodes) != nodes
    assert differs > 0        # the set-order quirk is actually exercised by the fixture


def test_sorted_order_differs_only_in_dependency_rows():
    traj = load_traj("traj_rolling_2d")
    data = load_rolling(str(traj["source"]), 4)
    traj = dict(traj); traj["num"] = 4
    a = list(replay_oracle_windows(traj, data, oracle.WINDOW_ORDER_REFERENCE))
    b = list(replay_oracle_windows(traj, data, oracle.WINDOW_ORDER_SORTED))
    neq = 0
    for (_, _, sa, da, na), (_, _, sb, db, nb) in zip(a, b):
        assert np.array_equal(sa, sb) and na == nb
        neq += not np.array_equal(da, db)
    assert neq > 0


@pytest.mark.parametrize("name", ["traj_rolling_3d", "traj_rolling_2d"])
def test_rolling_batch_driver_golden(name):
    """The threaded CPU driver (bench.py's baseline for the rolling workload) reproduces the recorded episodes."""
    traj = load_traj(name)
    num = int(traj["num"])
    data = load_rolling(str(traj["source"]), num)
    out = oracle.rolling_batch(data["adj"], data["blocks"], traj["ptr"].T.copy(), traj["container_size"].tolist(),
                               int(traj["window"]), str(traj["reward_type"]), str(traj["heightmap_type"]),
                               str(traj["packing_strategy"]), nthreads=3)
    assert out["status"] == 0
    assert np.array_equal(out["heightmap"], traj["heightmap"][:, -1])
    assert np.abs(out["reward"].astype(np.float64) - traj["ratio"]).max() <= 1e-6


@pytest.mark.skipif(not refshim.available(), reason="reference tree not present")
@pytest.mark.parametrize("dim,T,n,trials", [(2, 50, 10, 12), (3, 50, 10, 8), (2, 16, 10, 8), (3, 30, 12, 5), (2, 64, 8, 4)])
def test_window_differential_live(dim, T, n, trials):
    """generate.InitialContainer vs the oracle on fresh generate.generate_blocks instances under a random-valid policy
    (T=16/n=10 takes networkx's ascending-order branch, the others the set-order branch)."""
    from tests.golden.make_golden import ic_adjacency
    generate = refshim.load(("tools", "generate"))["generate"]
    np.random.seed(100 + dim * 7 + T)
    rng = np.random.RandomState(5)
    ics = [7, 250] if dim == 2 else [7, 7, 250]
    R = 2 if dim == 2 else 6
    for tr in range(trials):
        rot_blocks, positions, _, _, _ = generate.generate_blocks(T, ics, 1, [1, 5])
        blocks = np.asarray(rot_blocks).reshape(R, dim, T).transpose(0, 2, 1).reshape(R * T, dim)   # rolling.py:483-485
        pos = np.asarray(positions).reshape(dim, T).transpose(1, 0)                                  # rolling.py:490-491
        ic = generate.InitialContainer(blocks, pos, T, ics, True, n, "bot")
        oc = oracle.InitialContainer(ic_adjacency(ic), blocks, T, n, dim)
        while True:
            s_ref, d_ref = ic.convert_to_input()
            s, d = oc.convert_to_input()
            assert np.array_equal(s_ref, s) and np.array_equal(d_ref, d)
            assert [int(v) for v in ic.sub_graph_nodes] == oc.sub_graph_nodes
            assert ic.is_last_graph() == oc.is_last_graph()
            if ic.is_last_graph():
                break
            ok = np.nonzero((d[:n].sum(0) + d[n:2 * n].sum(0) * d[2 * n:].sum(0)) == 0)[0]
            ptr = int(rng.choice(ok)) if len(ok) else int(rng.randint(n * R))
            bid = int(ic.sub_graph_nodes[ptr % n])
            ic.remove_block(bid); oc.remove_block(bid)


def _product_calc_dependent():
    # host-side numpy code of the product package; imported lazily (the package dlopens libtapenv.so, no GPU needed)
    from tapenv.rolling import calc_dependent, pack_graphs
    return calc_dependent, pack_graphs


@pytest.mark.parametrize("name", ["rolling3d_t50.npz", "rolling2d_t50.npz"])
def test_calc_dependent_reproduces_the_fixture_graphs(name):
    """tapenv.rolling.calc_dependent (interval arithmetic) vs the graphs generate.InitialContainer built by voxel
    scanning, recorded in the fixtures: every instance, all five relations."""
    calc_dependent, pack_graphs = _product_calc_dependent()
    z = load_rolling(name)
    T, dim = z["T"], z["dim"]
    ics = [7, 250] if dim == 2 else [7, 7, 250]
    for b in range(z["adj"].shape[0]):
        got = calc_dependent(z["blocks"][b, :T], z["positions"][b], ics)
        assert np.array_equal(got.astype(np.uint8), z["adj"][b]), b
    pred = pack_graphs(z["adj"][:3])
    for g, u, v in [(0, 3, 5), (1, 0, 0), (2, 7, 7)]:
        assert ((int(pred[1, g, v]) >> u) & 1) == int(z["adj"][1, g, u, v])


@pytest.mark.skipif(not refshim.available(), reason="reference tree not present")
@pytest.mark.parametrize("dim,T,arm", [(2, 30, 1), (2, 20, 2), (3, 30, 1), (3, 12, 1)])
def test_calc_dependent_live(dim, T, arm):
    calc_dependent, _ = _product_calc_dependent()
    generate = refshim.load(("tools", "generate"))["generate"]
    np.random.seed(77 + dim + T)
    ics = [7, 250] if dim == 2 else [6, 8, 250]
    R = 2 if dim == 2 else 6
    for tr in range(12):
        rot_blocks, positions, _, _, _ = generate.generate_blocks(T, ics, arm, [1, 5])
        blocks = np.asarray(rot_blocks).reshape(R, dim, T).transpose(0, 2, 1).reshape(R * T, dim)[:T]
        pos = np.asarray(positions).reshape(dim, T).transpose(1, 0)
        cont = np.zeros(ics, dtype=int)
        for i in range(T):
            sl = tuple(slice(int(pos[i, d]), int(pos[i, d] + blocks[i, d])) for d in range(dim))
            cont[sl] = i + 1
        ref = generate.calc_dependent(blocks, pos, cont, arm)[:5]
        got = calc_dependent(blocks, pos, ics, arm)
        for g in range(5):
            assert np.array_equal(got[g], np.asarray(ref[g]).astype(bool)), (tr, g)
