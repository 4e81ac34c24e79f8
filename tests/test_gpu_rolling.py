"""GPU parity of the rolling window (tapenv_window_next / tapenv_rolling_step, generate.InitialContainer +
rolling.validate's loop) through the C ABI: bit-exact against the trajectories recorded from the live reference and
against the CPU oracle on larger seeded batches."""
import numpy as np
import pytest
import torch

from oracle import oracle
from tests.golden_io import load_rolling, load_traj

pytestmark = pytest.mark.gpu


def _tapenv():
    import tapenv
    return tapenv


def _dyn_from_bits(bits, n, S):
    return np.unpackbits(bits)[:3 * n * S].reshape(3 * n, S).astype(np.float32)


@pytest.mark.parametrize("name", ["traj_rolling_3d", "traj_rolling_2d"])
def test_window_sequence_vs_reference_trajectory(name):
    """convert_to_input / remove_block / is_last_graph replayed with the reference's own pointers."""
    tapenv = _tapenv()
    traj = load_traj(name)
    num, n = int(traj["num"]), int(traj["window"])
    data = load_rolling(str(traj["source"]), num)
    T, dim = data["T"], data["dim"]
    win = tapenv.BatchedInitialContainers(data["adj"], data["blocks"], T, n, dim)
    S = win.S
    calls = traj["static"].shape[1]
    ptrs = traj["ptr"]                                    # [num, T]
    t = 0
    for c in range(calls):
        cur = torch.empty(num, S, device="cuda"); mask = torch.empty(num, S, device="cuda")
        static, dynamic = win.convert_to_input(masks=(cur, mask))
        assert np.array_equal(static.cpu().numpy(), traj["static"][:, c].astype(np.float32)), c
        ref_dyn = np.stack([_dyn_from_bits(traj["dynamic_bits"][b, c], n, S) for b in range(num)])
        assert np.array_equal(dynamic.cpu().numpy(), ref_dyn), c
        assert np.array_equal(win.sub_graph_nodes.cpu().numpy(), traj["nodes"][:, c]), c
        assert np.array_equal(cur.cpu().numpy(), oracle.initial_mask(ref_dyn, n, win.rotate_types))
        assert bool((mask == 1).all())
        last = bool(win.is_last_graph().all())
        assert last == (c == calls - 1) and bool(win.is_last_graph().any()) == last
        t += n if last else 1
        win.remove_block(torch.from_numpy(ptrs[:, t - 1].astype(np.int64)).cuda())
    win.check_flags()


@pytest.mark.parametrize("name", ["traj_rolling_3d", "traj_rolling_2d"])
def test_rolling_runner_vs_reference_trajectory(name):
    """The fused rolling decode step: heightmap and returned encoding after EVERY step, final positions / stable /
    calc_ratio, against the live-reference recording."""
    tapenv = _tapenv()
    traj = load_traj(name)
    num, n = int(traj["num"]), int(traj["window"])
    data = load_rolling(str(traj["source"]), num)
    T, dim = data["T"], data["dim"]
    size = traj["container_size"].tolist()
    env = tapenv.BatchedContainers(size, T, str(traj["reward_type"]), str(traj["heightmap_type"]),
                                   packing_strategy=str(traj["packing_strategy"]), batch_size=num, window=n)
    win = tapenv.BatchedInitialContainers(data["adj"], data["blocks"], T, n, dim)
    run = tapenv.RollingRunner(env, win)
    S = win.S
    static, dynamic, cur = run.begin()
    call = 0
    for t in range(T):
        if t <= T - n:                                    # a fresh window is visible at the start of these steps
            assert np.array_equal(static.cpu().numpy(), traj["static"][:, call].astype(np.float32)), t
            ref_dyn = np.stack([_dyn_from_bits(traj["dynamic_bits"][b, call], n, S) for b in range(num)])
            assert np.array_equal(dynamic.cpu().numpy(), ref_dyn), t
            call += 1
        ptr = torch.from_numpy(traj["ptr"][:, t].astype(np.int64)).cuda()
        assert bool(torch.gather(cur, 1, ptr[:, None]).eq(1).all()), "recorded pointer must be accessible"
        static, dynamic, cur, dec_static, dec_dyn = run.step(ptr)
        assert np.array_equal(env.heightmap.cpu().numpy().reshape(num, -1), traj["heightmap"][:, t]), t
        assert np.array_equal(dec_dyn.cpu().numpy().reshape(num, -1), traj["dec_dyn"][:, t].astype(np.float32)), t
    assert np.array_equal(env.positions.cpu().numpy(), traj["positions"])
    assert np.array_equal(env.stable.cpu().numpy(), traj["stable"])
    r = env.calc_ratio().cpu().numpy().astype(np.float64)
    assert np.abs(r - traj["ratio"]).max() <= 1e-6
    env.check_flags(); win.check_flags()


def _record_policy(run, T, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    static, dynamic, cur = run.begin()
    ptrs = []
    for t in range(T):
        ptr = torch.multinomial(cur, 1, generator=g).squeeze(1)
        ptrs.append(ptr)
        static, dynamic, cur, _, _ = run.step(ptr)
    return torch.stack(ptrs)


@pytest.mark.parametrize("src,B,size,rt,strat", [
    ("rolling3d_t50.npz", 2048, [5, 5, 250], "C+P+S-lb-soft", "LB_GREEDY"),
    ("rolling2d_t50.npz", 1024, [5, 250], "C+P+S-lb-hard", "LB_GREEDY"),
    ("rolling2d_t50.npz", 512, [7, 250], "C+P+S-mcs-soft", "MACS"),
    ("rolling3d_t50.npz", 256, [5, 5, 250], "C+P+S-mcs-soft", "MACS"),     # voxel-state strategies: unfused rolling step
    ("rolling2d_t50.npz", 256, [5, 250], "C+P+S-lb-soft", "LB"),
    ("rolling3d_t50.npz", 128, [5, 5, 250], "C+P+S-lb-hard", "LB"),        # LB 3D, 250 levels: the warp form on level masks
    ("rolling3d_t50.npz", 64, [5, 5, 400], "C+P+S-lb-soft", "LB"),         # above 256 levels: the one-thread walk
    ("rolling3d_t50.npz", 128, [5, 5, 250], "C+P+S-mcs-hard", "MACS"),
])
def test_rolling_batch_vs_oracle(src, B, size, rt, strat):
    """Large tiled batches under an on-device random-valid policy, against the threaded CPU oracle driver."""
    tapenv = _tapenv()
    data = load_rolling(src)
    pool, T, dim = data["adj"].shape[0], data["T"], data["dim"]
    idx = np.arange(B) % pool
    adj, blocks = data["adj"][idx], data["blocks"][idx]
    n = 10
    env = tapenv.BatchedContainers(size, T, rt, "diff", packing_strategy=strat, batch_size=B, window=n)
    win = tapenv.BatchedInitialContainers(adj, blocks, T, n, dim)
    run = tapenv.RollingRunner(env, win)
    ptr_seq = _record_policy(run, T, seed=7)
    r = env.calc_ratio().cpu().numpy()
    hm = env.heightmap.cpu().numpy().reshape(B, -1)
    env.check_flags(); win.check_flags()
    o = oracle.rolling_batch(adj, blocks, ptr_seq.cpu().numpy(), size, n, rt, "diff", strat, nthreads=8)
    assert o["status"] == 0
    assert np.array_equal(hm, o["heightmap"])
    assert np.abs(r.astype(np.float64) - o["reward"].astype(np.float64)).max() <= 1e-6
    # replay reproduces itself (state fully reset by begin())
    r2 = run.run(ptr_seq).cpu().numpy()
    assert np.array_equal(r, r2)


@pytest.mark.parametrize("src,T,n,order", [
    ("rolling2d_t50.npz", 50, 20, 0),     # 20-node windows: CPython's set table grows to 128 slots (sequential path)
    ("rolling2d_t50.npz", 50, 9, 0),      # S = 18: scalar emission path
    ("rolling2d_t50.npz", 50, 25, 0),     # 2n >= T: networkx enumerates in ascending order
    ("rolling2d_t50.npz", 50, 10, 1),     # TAPENV_WINDOW_ORDER_SORTED
    ("rolling3d_t50.npz", 50, 7, 0),      # S = 42: scalar path in 3D
    ("rolling3d_t50.npz", 50, 10, 1),
    ("rolling2d_t50.npz", 50, 50, 0),     # window == total: not representable (S > 64) -> error, see below
])
def test_window_shapes_vs_oracle(src, T, n, order):
    tapenv = _tapenv()
    num = 48
    data = load_rolling(src, num)
    dim = data["dim"]
    R = 2 if dim == 2 else 6
    if n * R > 64:
        with pytest.raises(ValueError):
            tapenv.BatchedInitialContainers(data["adj"], data["blocks"], T, n, dim, node_order=order)
        return
    win = tapenv.BatchedInitialContainers(data["adj"], data["blocks"], T, n, dim, node_order=order)
    ocs = [oracle.InitialContainer(data["adj"][b], data["blocks"][b], T, n, dim, order=order) for b in range(num)]
    rng = np.random.RandomState(3)
    S = n * R
    for call in range(T - n + 1):
        static, dynamic = win.convert_to_input()
        outs = [oc.convert_to_input() for oc in ocs]
        assert np.array_equal(static.cpu().numpy(), np.stack([o[0] for o in outs])), call
        dref = np.stack([o[1] for o in outs])
        assert np.array_equal(dynamic.cpu().numpy(), dref), call
        assert np.array_equal(win.sub_graph_nodes.cpu().numpy(), np.array([oc.sub_graph_nodes for oc in ocs])), call
        last = all(oc.is_last_graph() for oc in ocs)
        assert bool(win.is_last_graph().all()) == last
        if last:
            break
        cur = oracle.initial_mask(dref, n, R)
        u = rng.random_sample((num, S)) * (cur > 0) + 1e-9 * rng.random_sample((num, S))
        ptr = np.argmax(u, axis=1).astype(np.int64)
        for b, oc in enumerate(ocs):
            oc.remove_block(oc.sub_graph_nodes[int(ptr[b]) % n])
        win.remove_block(torch.from_numpy(ptr).cuda())
    win.check_flags()


def test_window_flags_and_errors():
    tapenv = _tapenv()
    data = load_rolling("rolling2d_t50.npz", 4)
    win = tapenv.BatchedInitialContainers(data["adj"], data["blocks"], 50, 10, 2)
    win.convert_to_input()
    win.remove_block(torch.tensor([0, 19, 20, -1], dtype=torch.int64, device="cuda"))      # 20 and -1 are outside [0,S)
    win.convert_to_input()
    assert win.flags.cpu().tolist() == [0, 0, 4, 4]
    with pytest.raises(IndexError):
        win.check_flags()
    # a cyclic movement graph: the reference would never return
    adj = data["adj"].copy()
    adj[:, 0] = 0
    for u in range(50):
        adj[0, 0, u, (u + 1) % 50] = 1
    cyc = tapenv.BatchedInitialContainers(adj, data["blocks"], 50, 10, 2)
    cyc.convert_to_input()
    f = cyc.flags.cpu().tolist()
    assert f[0] & 1 and f[1] == 0
    with pytest.raises(ValueError):
        tapenv.BatchedInitialContainers(data["adj"], data["blocks"], 50, 10, 2, input_type="rot")
    empty = tapenv.BatchedInitialContainers(data["adj"][:0], data["blocks"][:0], 50, 10, 2)
    s, d = empty.convert_to_input()
    assert s.shape == (0, 3, 20) and d.shape == (0, 30, 20)


def test_blocks_without_rotation_structure_take_the_generic_path():
    """blocks_are_rotations is only an optimisation: arbitrary [R*T,dim] block tables give the same static rows as the
    oracle (every rotation row read from the table), and the flag is detected from the data."""
    tapenv = _tapenv()
    data = load_rolling("rolling3d_t50.npz", 16)
    T, n, dim = 50, 10, 3
    blocks = data["blocks"].copy()
    rng = np.random.RandomState(0)
    blocks[:, T:] = rng.randint(1, 6, size=blocks[:, T:].shape)          # rotations no longer permutations of row i
    win = tapenv.BatchedInitialContainers(data["adj"], blocks, T, n, dim)
    assert win.wcfg.blocks_are_rotations == 0
    ok = tapenv.BatchedInitialContainers(data["adj"], data["blocks"], T, n, dim)
    assert ok.wcfg.blocks_are_rotations == 1
    ocs = [oracle.InitialContainer(data["adj"][b], blocks[b], T, n, dim) for b in range(16)]
    for call in range(5):
        static, dynamic = win.convert_to_input()
        outs = [oc.convert_to_input() for oc in ocs]
        assert np.array_equal(static.cpu().numpy(), np.stack([o[0] for o in outs]))
        assert np.array_equal(dynamic.cpu().numpy(), np.stack([o[1] for o in outs]))
        ptr = np.zeros(16, np.int64)
        for b, oc in enumerate(ocs):
            oc.remove_block(oc.sub_graph_nodes[0])
        win.remove_block(torch.from_numpy(ptr).cuda())
    # the fused step gathers the chosen block from the same table
    env = tapenv.BatchedContainers([5, 5, 250], T, "C+P+S-lb-soft", "diff", batch_size=16, window=n)
    win2 = tapenv.BatchedInitialContainers(data["adj"], blocks, T, n, dim)
    run = tapenv.RollingRunner(env, win2)
    static, dynamic, cur = run.begin()
    ptr = torch.argmax(cur, dim=1)
    want = torch.gather(static[:, 1:], 2, ptr.view(-1, 1, 1).expand(-1, dim, 1)).squeeze(2).clone()
    _, _, _, dec_static, _ = run.step(ptr)
    assert torch.equal(dec_static, want)


def test_per_instance_initial_container_protocol():
    """tapenv.rolling.InitialContainer: the reference's constructor (blocks, positions, ...) and the three calls of
    rolling.validate (rolling.py:593-640), NumPy in / NumPy out, against the oracle driven the same way; the graphs
    come from tapenv.rolling.calc_dependent."""
    from tapenv.rolling import InitialContainer
    data = load_rolling("rolling3d_t50.npz", 3)
    T, n, dim = 50, 10, 3
    rng = np.random.RandomState(2)
    for b in range(3):
        ic = InitialContainer(data["blocks"][b], data["positions"][b], T, [7, 7, 250], True, n, "bot")
        assert np.array_equal(ic.deps.astype(np.uint8), data["adj"][b])
        oc = oracle.InitialContainer(data["adj"][b], data["blocks"][b], T, n, dim)
        steps = 0
        while True:
            static, dynamic = ic.convert_to_input()
            s_ref, d_ref = oc.convert_to_input()
            assert static.dtype == np.int64 and dynamic.dtype == np.float64
            assert np.array_equal(static, s_ref) and np.array_equal(dynamic, d_ref)
            assert ic.sub_graph_nodes == oc.sub_graph_nodes and ic.is_last_graph() == oc.is_last_graph()
            if ic.is_last_graph():
                break
            ptr = int(rng.randint(n * 6))
            while ptr >= n:
                ptr -= n
            bid = ic.sub_graph_nodes[ptr]
            ic.remove_block(bid); oc.remove_block(bid)
            ic.remove_block(999)                                # unknown id: ignored like the reference's try/except
            steps += 1
        assert steps == T - n


def _sub_instance(data, T):
    """The sub-instance on the first T nodes (edges among them, their block rows in every rotation)."""
    T0 = data["T"]
    R = data["blocks"].shape[1] // T0
    adj = data["adj"][:, :, :T, :T].copy()
    rows = np.concatenate([np.arange(T) + r * T0 for r in range(R)])
    return adj, data["blocks"][:, rows].copy()


@pytest.mark.parametrize("src,T,n", [("rolling2d_t50.npz", 10, 10), ("rolling2d_t50.npz", 12, 5), ("rolling3d_t50.npz", 10, 10),
                                     ("rolling3d_t50.npz", 21, 10), ("rolling2d_t50.npz", 33, 16), ("rolling2d_t50.npz", 3, 1)])
def test_small_totals_and_single_window(src, T, n):
    """window == total (a single, fully decoded window: no rolling step at all), totals around 2*window (the branch
    point of networkx's enumeration), a window of one node."""
    tapenv = _tapenv()
    data = load_rolling(src, 24)
    dim = data["dim"]
    adj, blocks = _sub_instance(data, T)
    size = [5, 250] if dim == 2 else [5, 5, 250]
    env = tapenv.BatchedContainers(size, T, "C+P+S-lb-soft", "diff", batch_size=24, window=n)
    win = tapenv.BatchedInitialContainers(adj, blocks, T, n, dim)
    run = tapenv.RollingRunner(env, win)
    ptr_seq = _record_policy(run, T, seed=3)
    o = oracle.rolling_batch(adj, blocks, ptr_seq.cpu().numpy(), size, n, "C+P+S-lb-soft", "diff", "LB_GREEDY", nthreads=4)
    assert o["status"] == 0
    assert np.array_equal(env.heightmap.cpu().numpy().reshape(24, -1), o["heightmap"])
    assert np.abs(env.calc_ratio().cpu().numpy().astype(np.float64) - o["reward"]).max() <= 1e-6
    assert int(env.current_blocks_num.min()) == T
    env.check_flags(); win.check_flags()


def test_full_size_rolling_properties():
    """BASELINE config 5 at its FULL size (65 536 instances x 50 blocks on one GPU), through size-independent properties:
    (a) every instance packs each of its 50 blocks exactly once and the window tensors stay well-formed,
    (b) volume conservation: sum(heightmap) == valid_size + empty_size and valid_size == sum of the placed volumes,
    (c) instances are independent: the batch is 512 distinct instances tiled 128 times under the SAME pointers, so all
        copies must agree bit for bit -- and with the oracle on the 512 originals."""
    tapenv = _tapenv()
    data = load_rolling("rolling3d_t50.npz")
    pool, T, n, dim = data["adj"].shape[0], 50, 10, 3
    B = 65536
    idx = np.arange(B) % pool
    size = [5, 5, 250]
    env = tapenv.BatchedContainers(size, T, "C+P+S-lb-soft", "diff", batch_size=B, window=n)
    win = tapenv.BatchedInitialContainers(data["adj"][idx], data["blocks"][idx], T, n, dim)
    run = tapenv.RollingRunner(env, win)
    static, dynamic, cur = run.begin()
    ptrs, nodes_seen = [], torch.zeros(B, T, dtype=torch.int32, device="cuda")
    for t in range(T):
        ptr = torch.argmax(cur[:pool] * torch.arange(cur.shape[1], 0, -1, device="cuda"), dim=1).repeat(B // pool)   # first accessible
        assert bool(torch.gather(cur, 1, ptr[:, None]).eq(1).all())
        if t <= T - n:
            chosen = torch.gather(win.sub_graph_nodes, 1, (ptr % n)[:, None].to(torch.int64)).squeeze(1)
        else:                                             # inside the last window the node list no longer changes
            chosen = torch.gather(last_nodes, 1, (ptr % n)[:, None].to(torch.int64)).squeeze(1)
        if t == T - n:
            last_nodes = win.sub_graph_nodes.clone()
        nodes_seen.scatter_add_(1, chosen[:, None].to(torch.int64), torch.ones(B, 1, dtype=torch.int32, device="cuda"))
        ptrs.append(ptr)
        static, dynamic, cur, dec_static, _ = run.step(ptr)
        assert bool((dec_static >= 1).all())
    assert bool((nodes_seen == 1).all())                                                        # (a)
    hm = env.heightmap.reshape(B, -1).to(torch.int64)
    sc = env.scalars.to(torch.int64)
    assert bool((hm.sum(1) == sc[:, 0] + sc[:, 1]).all())                                       # (b)
    vol = (env.blocks.to(torch.int64).prod(2) * 1).sum(1)
    assert bool((sc[:, 0] == vol).all()) and int(sc[:, 3].min()) == T                           # soft reward: every block is placed
    r = env.calc_ratio()
    assert bool(torch.equal(hm.view(B // pool, pool, -1), hm[:pool].expand(B // pool, pool, -1)))   # (c)
    assert bool(torch.equal(r.view(B // pool, pool), r[:pool].expand(B // pool, pool)))
    o = oracle.rolling_batch(data["adj"], data["blocks"], torch.stack(ptrs)[:, :pool].cpu().numpy(), size, n,
                             "C+P+S-lb-soft", "diff", "LB_GREEDY", nthreads=8)
    assert o["status"] == 0 and np.array_equal(hm[:pool].cpu().numpy(), o["heightmap"])
    assert np.abs(r[:pool].cpu().numpy().astype(np.float64) - o["reward"]).max() <= 1e-6
    env.check_flags(); win.check_flags()


@pytest.mark.parametrize("seed", range(int(__import__("os").environ.get("TAPENV_FUZZ_SEEDS", "16"))))
def test_window_random_graphs_fuzz(seed):
    """Random DAGs (any density, multi-wave admissions), random rotation graphs with wall self-loops, arbitrary block tables,
    totals 1..64, windows 1..32 (2D) / 1..10 (3D), both node orders, arbitrary (not necessarily accessible) pointers:
    every window bit-exact against the oracle."""
    tapenv = _tapenv()
    rng = np.random.RandomState(5000 + seed)
    dim = 2 if seed % 2 else 3
    T = int(rng.randint(1, 65))
    n = int(rng.randint(1, min(T, 32 if dim == 2 else 10) + 1))
    R = 2 if dim == 2 else 6
    B = int(rng.randint(1, 40))
    order = int(rng.randint(2))
    dens = [0.02, 0.08, 0.3][int(rng.randint(3))]
    adj = np.zeros((B, 5, T, T), np.uint8)
    perm = np.stack([rng.permutation(T) for _ in range(B)])
    for b in range(B):
        rank = np.empty(T, int); rank[perm[b]] = np.arange(T)
        m = (rng.random_sample((T, T)) < dens) & (rank[:, None] > rank[None, :])       # edge u -> v only if u is later in a random order: a DAG
        adj[b, 0] = m
        for g in range(1, 5 if dim == 3 else 3):
            adj[b, g] = rng.random_sample((T, T)) < dens * 0.5
            adj[b, g][np.arange(T), np.arange(T)] = rng.random_sample(T) < 0.2            # against the wall
    blocks = rng.randint(1, 6, size=(B, R * T, dim)).astype(np.int32)
    win = tapenv.BatchedInitialContainers(adj, blocks, T, n, dim, node_order=order)
    ocs = [oracle.InitialContainer(adj[b], blocks[b], T, n, dim, order=order) for b in range(B)]
    S = n * R
    for call in range(T - n + 1):
        static, dynamic = win.convert_to_input()
        outs = [oc.convert_to_input() for oc in ocs]
        assert np.array_equal(static.cpu().numpy(), np.stack([o[0] for o in outs])), (call, T, n)
        assert np.array_equal(dynamic.cpu().numpy(), np.stack([o[1] for o in outs])), (call, T, n)
        assert np.array_equal(win.sub_graph_nodes.cpu().numpy(), np.array([oc.sub_graph_nodes for oc in ocs])), call
        assert bool(win.is_last_graph().all()) == all(oc.is_last_graph() for oc in ocs)
        ptr = rng.randint(0, S, size=B).astype(np.int64)
        for b, oc in enumerate(ocs):
            oc.remove_block(oc.sub_graph_nodes[int(ptr[b]) % n])
        win.remove_block(torch.from_numpy(ptr).cuda())
    win.check_flags()
