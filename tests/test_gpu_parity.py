"""GPU (B200): the CUDA path, called through the C ABI, against the CPU oracle and the reference-derived
golden fixtures.  Integer / index / mask outputs must be bit-exact; the reward is compared as float32
within 1e-6 (north_star tolerance) and is in practice bit-equal."""
import os

import numpy as np
import pytest

from tests.golden_io import load_inputs, load_traj, golden_path
from tests.rollout import oracle_rollout, random_valid_ptrs

pytestmark = pytest.mark.gpu
REWARD_TOL = 1e-6


def _torch():
    import torch
    return torch


def gpu_rollout(static, dynamic, ptr_seq, container_size, reward_type, heightmap_type, packing_strategy, fused=True):
    """Drive BatchedContainers like model.py drives the reference; returns per-step numpy arrays."""
    torch = _torch()
    import tapenv
    dev = torch.device("cuda:0")
    B, _, S = static.shape
    dim = len(container_size)
    n = S // (2 if dim == 2 else 6)
    st = torch.from_numpy(static).to(dev)
    dyn = torch.from_numpy(dynamic).to(dev)
    dyn0 = dyn.clone()
    env = tapenv.BatchedContainers(container_size, n, reward_type, heightmap_type, packing_strategy=packing_strategy,
                                   batch_size=B, device=dev)
    cur, mask = env.reset(dyn)
    import ctypes
    if tapenv._capi.lib.tapenv_add_blocks(ctypes.byref(env.cfg), None, None, None, None) == tapenv._capi.EUNSUPPORTED:
        pytest.skip("strategy not built yet")
    out = dict(heightmap=[], dec_dyn=[], cur_mask=[cur.cpu().numpy()], mask=[], valid=[], empty=[], dynamic=[], dec_static=[])
    for t in range(ptr_seq.shape[0]):
        ptr = torch.from_numpy(ptr_seq[t]).to(dev)
        if fused:
            dyn_in = dyn
            dyn, cur, mask, dec_static, dec_dyn = env.step(ptr, st, dyn_in, mask)
        else:
            dyn = tapenv.update_dynamic(dyn, st, ptr, "bot", True)
            cur, mask = tapenv.update_mask(mask, dyn, st, ptr, "bot", True)
            dec_static = torch.gather(st[:, 1:1 + dim], 2, ptr.view(-1, 1, 1).expand(-1, dim, 1)).squeeze(2)
            dec_dyn = env.add_new_blocks(dec_static)
        out["heightmap"].append(env.heightmap.cpu().numpy().reshape(B, -1))
        out["dec_dyn"].append(dec_dyn.cpu().numpy().reshape(B, -1)); out["dec_static"].append(dec_static.cpu().numpy())
        out["cur_mask"].append(cur.cpu().numpy()); out["mask"].append(mask.cpu().numpy())
        out["valid"].append(env.valid_size.cpu().numpy()); out["empty"].append(env.empty_size.cpu().numpy())
        out["dynamic"].append(dyn.cpu().numpy())
    assert torch.equal(dyn0, torch.from_numpy(dynamic).to(dev)), "inputs must never be modified (pack.py:370 clone)"
    res = {k: np.stack(v) for k, v in out.items()}
    res["positions"] = env.positions.cpu().numpy()
    res["stable"] = env.stable.cpu().numpy()
    res["reward"] = env.calc_ratio().cpu().numpy()
    res["k"] = env.current_blocks_num.cpu().numpy()
    res["flags"] = env.flags.cpu().numpy()
    return res


def assert_same(g, r, ptr_seq, static, dim):
    B = static.shape[0]
    for k in ("heightmap", "cur_mask", "mask", "valid", "empty", "positions", "stable", "dynamic"):
        assert np.array_equal(g[k], r[k]), k
    assert np.array_equal(g["dec_dyn"], r["dec_dyn"].astype(np.float32)), "encoded heightmap"
    want_static = np.stack([static[np.arange(B)[:, None], 1 + np.arange(dim)[None, :], p[:, None]] for p in ptr_seq])
    assert np.array_equal(g["dec_static"], want_static)
    gr, rr = g["reward"].astype(np.float64), np.asarray(r["ratio"], dtype=np.float64)
    # an environment in which nothing could be placed has height 0: calc_CPS divides 0 by numpy's 0 -> nan on both sides
    assert np.array_equal(np.isnan(gr), np.isnan(rr))
    ok = ~np.isnan(rr)
    assert not ok.any() or np.abs(gr[ok] - rr[ok]).max() <= REWARD_TOL
    assert np.array_equal(g["reward"], rr.astype(np.float32), equal_nan=True)     # and in fact bit-equal after fp64->fp32
    assert (g["flags"] == 0).all()


TRAJ = ["traj_2d_lbg_soft", "traj_2d_lbg_hard", "traj_2d_lbg_w7_full", "traj_2d_macs_rand", "traj_2d_macs_ppsg",
        "traj_3d_lbg_soft", "traj_3d_lbg_hard", "traj_2d_lb_soft", "traj_2d_lb_hard", "traj_3d_lb_soft",
        "traj_3d_macs_soft", "traj_3d_macs_hard"]


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("name", TRAJ)
def test_reference_trajectories(name, fused):
    """CUDA vs trajectories recorded from the live Python reference (independent of the C oracle)."""
    if not os.path.exists(golden_path(name + ".npz")):
        pytest.skip("fixture not generated")
    t = load_traj(name)
    static, dynamic = load_inputs(str(t["source"]), int(t["num"]))
    size = t["container_size"].tolist()
    g = gpu_rollout(static, dynamic, t["ptr"], size, str(t["reward_type"]), str(t["heightmap_type"]),
                    str(t["packing_strategy"]), fused=fused)
    B = static.shape[0]
    for k in ("cur_mask", "mask", "valid", "empty", "positions", "stable"):
        assert np.array_equal(g[k], t[k]), k
    assert np.array_equal(g["heightmap"], t["heightmap"].reshape(g["heightmap"].shape))
    assert np.array_equal(g["dec_dyn"], t["dec_dyn"].astype(np.float32))
    assert np.abs(g["reward"].astype(np.float64) - t["ratio"]).max() <= REWARD_TOL
    assert np.array_equal(g["reward"], t["ratio"].astype(np.float32))


CASES = [
    # fixture, num, container, reward_type, heightmap_type, strategy
    ("rand2d_n10.npz", 4096, [5, 50], "C+P+S-lb-soft", "diff", "LB_GREEDY"),     # BASELINE config 2
    ("rand2d_n10.npz", 1024, [5, 50], "C+P+S-lb-hard", "zero", "LB_GREEDY"),
    ("rand2d_n10.npz", 515, [4, 50], "C+P+S-lb-hard", "full", "LB_GREEDY"),      # width 4: frequent unplaced blocks
    ("rand2d_n10.npz", 257, [9, 50], "C+P-lb-soft", "diff", "LB_GREEDY"),
    ("rand2d_n10.npz", 300, [32, 50], "C+P+S-lb-soft", "diff", "LB_GREEDY"),     # maximum compiled width
    ("rand2d_n10.npz", 1024, [5, 50], "C+P+S-mcs-soft", "diff", "MACS"),
    ("rand2d_n10.npz", 512, [5, 50], "C+P+S-mcs-hard", "full", "MACS"),
    ("ppsg2d_n20.npz", 512, [7, 50], "C+P+S-mcs-hard", "diff", "MACS"),          # BASELINE config 4 (pool)
    ("rand3d_n10.npz", 2048, [5, 5, 50], "C+P+S-lb-soft", "diff", "LB_GREEDY"),  # BASELINE config 3
    ("rand3d_n10.npz", 512, [5, 5, 50], "C+P+S-lb-hard", "zero", "LB_GREEDY"),
    ("rand3d_n10.npz", 256, [4, 6, 50], "C+P+S-lb-soft", "full", "LB_GREEDY"),   # W != L
    ("rand2d_n10.npz", 1024, [5, 50], "C+P+S-lb-soft", "diff", "LB"),            # the older corner-list strategy (a11)
    ("rand2d_n10.npz", 512, [6, 50], "C+P+S-lb-hard", "zero", "LB"),
    ("rand3d_n10.npz", 512, [5, 5, 50], "C+P+S-lb-soft", "diff", "LB"),
    ("rand3d_n10.npz", 256, [4, 6, 50], "C+P+S-lb-hard", "full", "LB"),
    ("rand3d_n10.npz", 384, [5, 5, 50], "C+P+S-mcs-soft", "diff", "MACS"),       # calc_one_position_mcs_3d
    ("rand3d_n10.npz", 256, [5, 5, 50], "C+P+S-mcs-hard", "zero", "MACS"),
    ("rand3d_n10.npz", 128, [4, 6, 60], "mcs-hard", "full", "MACS"),             # every score 0.0: pure usable-space choice
    ("rand3d_n10.npz", 128, [6, 4, 60], "C+P-mcs-soft", "diff", "MUL"),
]


@pytest.fixture
def step_form(request):
    """The fused step exists in two launch forms -- one warp per environment (step_kernel) and one CTA per environment
    (step_split_kernel: copy warps + a placement warp); tapenv_step picks by batch size.  TAPENV_SPLIT=0/1 forces one."""
    old = os.environ.get("TAPENV_SPLIT")
    os.environ["TAPENV_SPLIT"] = {"warp": "0", "cta": "1"}[request.param]
    yield request.param
    if old is None:
        os.environ.pop("TAPENV_SPLIT", None)
    else:
        os.environ["TAPENV_SPLIT"] = old


@pytest.mark.parametrize("step_form", ["warp", "cta"], indirect=True)
@pytest.mark.parametrize("case", CASES, ids=lambda c: "%s-%s-%s-%s" % (c[0].split("_")[0], "x".join(map(str, c[2])), c[3], c[4]))
def test_fused_step_matches_oracle_every_step(case, step_form):
    src, num, size, rt, hm, strat = case
    if not os.path.exists(golden_path(src)):
        pytest.skip("fixture not generated")
    if step_form == "cta" and (strat in ("LB", "MUL") or (strat == "MACS" and len(size) == 3)):
        pytest.skip("voxel-state strategies have no CTA-per-environment form")
    static, dynamic = load_inputs(src, num)
    r = oracle_rollout(static, dynamic, size, rt, hm, strat, seed=7)
    g = gpu_rollout(static, dynamic, r["ptr"], size, rt, hm, strat, fused=True)
    assert_same(g, r, r["ptr"], static, len(size))


@pytest.mark.parametrize("case", [CASES[0], CASES[2], CASES[6], CASES[9], CASES[11], CASES[13]],
                         ids=["2d", "2d-w4-hard", "macs", "3d", "lb-2d", "lb-3d"])
def test_unfused_ops_match_oracle_every_step(case):
    src, num, size, rt, hm, strat = case
    if not os.path.exists(golden_path(src)):
        pytest.skip("fixture not generated")
    static, dynamic = load_inputs(src, min(num, 384))
    r = oracle_rollout(static, dynamic, size, rt, hm, strat, seed=11)
    g = gpu_rollout(static, dynamic, r["ptr"], size, rt, hm, strat, fused=False)
    assert_same(g, r, r["ptr"], static, len(size))


@pytest.mark.parametrize("B", [0, 1, 3, 33])
def test_ragged_batch_sizes(B):
    torch = _torch()
    import tapenv
    static, dynamic = load_inputs("rand2d_n10.npz", max(B, 1))
    static, dynamic = static[:B], dynamic[:B]
    if B == 0:
        env = tapenv.BatchedContainers([5, 50], 10, "C+P+S-lb-soft", "diff", batch_size=0)
        cur, mask = env.reset(torch.zeros(0, 30, 20, device="cuda"))
        assert cur.shape == (0, 20) and env.calc_ratio().shape == (0,)
        return
    r = oracle_rollout(static, dynamic, [5, 50], "C+P+S-lb-soft", "diff", "LB_GREEDY", seed=3)
    g = gpu_rollout(static, dynamic, r["ptr"], [5, 50], "C+P+S-lb-soft", "diff", "LB_GREEDY")
    assert_same(g, r, r["ptr"], static, 2)


def test_known_answer_sequences_on_gpu():
    torch = _torch()
    import tapenv
    kat = np.load(golden_path("kat.npz"))
    for name in ("G1", "G2", "DRAW", "G3", "G4"):
        size = kat[name + "_size"].tolist()
        blocks = kat[name + "_blocks"]
        try:
            env = tapenv.BatchedContainers(size, len(blocks), str(kat[name + "_reward_type"]), "diff",
                                           packing_strategy=str(kat[name + "_strategy"]), batch_size=2)
        except tapenv.TapEnvError as e:
            if e.code == -6:
                continue
            raise
        for t, b in enumerate(blocks):
            enc = env.add_new_blocks(torch.tensor(np.stack([b, b]), dtype=torch.float32, device="cuda"))
            assert np.array_equal(enc.cpu().numpy().reshape(2, -1)[1], kat[name + "_enc"][t].astype(np.float32)), (name, t)
            assert np.array_equal(env.heightmap.cpu().numpy().reshape(2, -1)[0], kat[name + "_heightmaps"][t]), (name, t)
        assert np.array_equal(env.positions.cpu().numpy()[1], kat[name + "_positions"])
        assert np.array_equal(env.stable.cpu().numpy()[0], kat[name + "_stable"])
        assert abs(float(env.calc_ratio()[0]) - float(kat[name + "_ratio"])) <= REWARD_TOL


def test_unplaceable_and_overflow_edge_cases():
    """Q1: a block wider than the container is not placed, state unchanged, k still advances; more than
    blocks_num blocks / stacks above container height set the sticky flags instead of raising IndexError."""
    torch = _torch()
    import tapenv
    from oracle import oracle
    env = tapenv.BatchedContainers([5, 12], 4, "C+P+S-lb-soft", "full", batch_size=1)
    ref = oracle.Container([5, 12], 4, "C+P+S-lb-soft", "full")
    for b in ([3, 2], [6, 1], [2, 4]):
        enc = env.add_new_blocks(torch.tensor([b], dtype=torch.float32, device="cuda"))
        want = ref.add_new_block(np.array(b, np.float32))
        assert np.array_equal(enc.cpu().numpy()[0], want.astype(np.float32))
    assert env.stable.cpu().numpy()[0].tolist() == [int(s) for s in ref.stable]
    assert env.positions.cpu().numpy()[0].tolist() == ref.positions.tolist()
    assert int(env.current_blocks_num[0]) == 3 and abs(float(env.calc_ratio()[0]) - ref.calc_ratio()) <= REWARD_TOL
    assert int(env.flags[0]) == 0
    env.add_new_blocks(torch.tensor([[5, 4]], dtype=torch.float32, device="cuda"))
    env.add_new_blocks(torch.tensor([[1, 1]], dtype=torch.float32, device="cuda"))      # 5th block of 4
    assert int(env.flags[0]) & 2
    with pytest.raises(IndexError):
        env.check_flags()
    env.clear_container()
    assert int(env.flags[0]) == 0 and int(env.heightmap.sum()) == 0
    for _ in range(4):
        env.add_new_blocks(torch.tensor([[5, 4]], dtype=torch.float32, device="cuda"))  # 16 > height 12
    assert int(env.flags[0]) & 1


def test_container_proxy_drop_in_for_unmodified_model_loop():
    """model.py:294, :452-453, :509-510 verbatim usage pattern against the per-environment views."""
    torch = _torch()
    import tapenv
    static, dynamic = load_inputs("rand2d_n10.npz", 64)
    B, n, dim = 64, 10, 2
    r = oracle_rollout(static, dynamic, [5, 50], "C+P+S-lb-soft", "diff", "LB_GREEDY", seed=5)
    containers = tapenv.BatchedContainers([5, 50], n, "C+P+S-lb-soft", "diff", batch_size=B)
    st = torch.from_numpy(static).cuda()
    for t in range(n):
        ptr = torch.from_numpy(r["ptr"][t]).cuda()
        decoder_static = torch.gather(st[:, 1:1 + dim], 2, ptr.view(-1, 1, 1).expand(-1, dim, 1))
        blocks = decoder_static.squeeze(2).cpu().numpy()                       # model.py:412
        is_rotate = (ptr < n).cpu().numpy().astype("bool")
        heightmaps = []
        for batch_index in range(B):                                           # model.py:452-453
            heightmaps.append(containers[batch_index].add_new_block(blocks[batch_index], is_rotate[batch_index]))
        assert np.array_equal(np.stack(heightmaps), r["dec_dyn"][t])
    scores = [containers[b].calc_ratio() for b in range(B)]                    # model.py:509-510
    assert np.abs(np.array(scores) - r["ratio"]).max() <= REWARD_TOL
    assert containers[3].positions.tolist() == r["positions"][3].tolist()
    assert containers[5].stable == [bool(v) for v in r["stable"][5]]


def test_reward_partial_sums_are_deterministic():
    torch = _torch()
    import tapenv
    static, dynamic = load_inputs("rand2d_n10.npz", 1000)
    ptrs = random_valid_ptrs(static, dynamic, [5, 50], seed=1)
    env = tapenv.BatchedContainers([5, 50], 10, "C+P+S-lb-soft", "diff", batch_size=1000)
    st, dyn = torch.from_numpy(static).cuda(), torch.from_numpy(dynamic).cuda()
    cur, mask = env.reset(dyn)
    for t in range(10):
        dyn, cur, mask, _, _ = env.step(torch.from_numpy(ptrs[t]).cuda(), st, dyn, mask)
    r, s = env.calc_ratio(partial_sums=True)
    r2, s2 = env.calc_ratio(partial_sums=True)
    assert torch.equal(s, s2) and float(s[2]) == 1000.0
    r64 = r.double().cpu().numpy()
    assert abs(float(s[0]) - r64.sum()) < 1e-9 and abs(float(s[1]) - (r64 ** 2).sum()) < 1e-9


@pytest.mark.parametrize("case", [CASES[0], CASES[1], CASES[3], CASES[5], CASES[7], CASES[8], CASES[9], CASES[11], CASES[13], CASES[15], CASES[16]],
                         ids=["2d-soft", "2d-hard", "2d-w9", "macs-rand", "macs-ppsg", "3d-soft", "3d-hard", "lb-2d", "lb-3d", "macs3d-soft", "macs3d-hard"])
def test_whole_episode_kernel_matches_stepwise_oracle(case):
    """K7 (tapenv_episode): one launch per episode == reset + n steps + calc_ratio of the oracle."""
    torch = _torch()
    import tapenv
    src, num, size, rt, hm, strat = case
    if not os.path.exists(golden_path(src)):
        pytest.skip("fixture not generated")
    static, dynamic = load_inputs(src, min(num, 1024))
    B = static.shape[0]
    r = oracle_rollout(static, dynamic, size, rt, hm, strat, seed=21)
    n = r["ptr"].shape[0]
    env = tapenv.BatchedContainers(size, n, rt, hm, packing_strategy=strat, batch_size=B)
    st, dyn = torch.from_numpy(static).cuda(), torch.from_numpy(dynamic).cuda()
    for steps in (n, 3, 0):
        reward, cur, mask, dec_dyn = env.episode(st, dyn, torch.from_numpy(r["ptr"][:steps]).cuda().reshape(steps, B))
        assert np.array_equal(cur.cpu().numpy(), r["cur_mask"][steps])
        if steps:
            assert np.array_equal(mask.cpu().numpy(), r["mask"][steps - 1])
            assert np.array_equal(env.heightmap.cpu().numpy().reshape(B, -1), r["heightmap"][steps - 1])
            assert np.array_equal(dec_dyn.cpu().numpy().reshape(B, -1), r["dec_dyn"][steps - 1].astype(np.float32))
            assert np.array_equal(env.valid_size.cpu().numpy(), r["valid"][steps - 1])
            assert np.array_equal(env.empty_size.cpu().numpy(), r["empty"][steps - 1])
        else:
            assert int(env.heightmap.sum()) == 0 and float(mask.min()) == 1.0
        if steps == n:
            assert np.array_equal(reward.cpu().numpy(), r["ratio"].astype(np.float32))
            assert np.array_equal(env.positions.cpu().numpy(), r["positions"])
            assert np.array_equal(env.stable.cpu().numpy(), r["stable"])
        assert (env.flags.cpu().numpy() == 0).all()
    assert torch.equal(dyn, torch.from_numpy(dynamic).cuda())


def test_lazy_container_list_as_model_py_builds_it():
    """`[tools.Container(...) for _ in range(B)]` (model.py:294) with tapenv.Container substituted: the B objects
    bind to ONE batch on the first add_new_block and behave like the reference objects."""
    torch = _torch()
    import tapenv
    static, dynamic = load_inputs("rand3d_n10.npz", 48)
    B, n, dim, size = 48, 10, 3, [5, 5, 50]
    r = oracle_rollout(static, dynamic, size, "C+P+S-lb-soft", "diff", "LB_GREEDY", seed=9)
    containers = [tapenv.Container(size, n, "C+P+S-lb-soft", "diff", packing_strategy="LB_GREEDY") for _ in range(B)]
    st = torch.from_numpy(static).cuda()
    for t in range(n):
        ptr = torch.from_numpy(r["ptr"][t]).cuda()
        blocks = torch.gather(st[:, 1:1 + dim], 2, ptr.view(-1, 1, 1).expand(-1, dim, 1)).squeeze(2).cpu().numpy()
        hms = [containers[b].add_new_block(blocks[b], False) for b in range(B)]
        assert hms[0].shape == (2, 5, 5)
        assert np.array_equal(np.stack(hms).reshape(B, -1), r["dec_dyn"][t])
    assert containers[0]._batch is containers[B - 1]._batch and containers[0]._batch.batch_size == B
    assert np.abs(np.array([c.calc_ratio() for c in containers]) - r["ratio"]).max() <= REWARD_TOL
    # a lone object is a batch of one
    one = tapenv.Container([5, 50], 3, "C+P+S-lb-soft", "diff")
    kat = np.load(golden_path("kat.npz"))
    for t, b in enumerate([[3, 3], [3, 3], [2, 3]]):
        assert np.array_equal(one.add_new_block(np.array(b, np.float32)), kat["G1_enc"][t])
    assert one.valid_size == 24 and one.current_blocks_num == 3 and one.heightmap.tolist() == [6, 6, 6, 3, 3]


def test_whole_episode_wrappers_known_answers():
    """tools.calc_positions_lb_greedy / pack.reward signatures: visual/draw_result.py:14-34 and SURVEY G1/G3."""
    torch = _torch()
    import tapenv
    pos, grid, stable, ratio, scores = tapenv.calc_positions_lb_greedy(np.array([[3, 2], [1, 1], [1, 2]]), [4, 6], "C+P+S-lb-hard")
    assert pos.tolist() == [[0, 0], [3, 0], [3, 1]] and stable == [True, True, True]
    assert ratio == 2.75 and scores == [9, 12, 0, 3, 3]
    # the reference's voxel `container` (tools.py:2168-2169): block ids k+1, 0 = empty
    assert grid.shape == (4, 6) and grid.tolist() == [[1, 1, 0, 0, 0, 0]] * 3 + [[2, 3, 3, 0, 0, 0]]
    kat = np.load(golden_path("kat.npz"))
    g1 = torch.tensor(kat["G1_blocks"], dtype=torch.float32).unsqueeze(0).repeat(3, 1, 1)
    pos, hm, stable, ratio, scores = tapenv.calc_positions_lb_greedy(g1, [5, 50], "C+P+S-lb-soft")
    assert np.array_equal(pos[2].cpu().numpy(), kat["G1_positions"]) and abs(float(ratio[1]) - 3 * 0.8775) < 1e-12
    g3 = torch.tensor(kat["G3_blocks"], dtype=torch.float32)
    pos, hm, stable, ratio, scores = tapenv.calc_positions_mcs(g3, [7, 100], "C+P+S-mcs-hard")
    assert np.array_equal(pos, kat["G3_positions"]) and abs(ratio - 3 * float(kat["G3_ratio"])) < 1e-12
    # pack.reward: tour indices into `static`
    static, dynamic = load_inputs("rand2d_n10.npz", 16)
    tour = np.stack([np.random.RandomState(i).permutation(10) + 10 * (i % 2) for i in range(16)])
    rw = tapenv.reward(torch.from_numpy(static), torch.from_numpy(tour), "C+P+S-lb-soft", "bot", True, 5, 50)
    from oracle import oracle
    for b in range(16):
        c = oracle.Container([5, 50], 10, "C+P+S-lb-soft", "full")
        for j in tour[b]:
            c.add_new_block(static[b, 1:3, j])
        assert abs(float(rw[b]) + 3 * c.calc_ratio()) <= 1e-6


def test_install_patches_reference_modules():
    import types
    import tapenv
    pack, tools = types.ModuleType("pack"), types.ModuleType("tools")
    pack.update_dynamic = pack.update_mask = pack.reward = tools.Container = "orig"
    tools.calc_positions_lb_greedy = tools.calc_positions_mcs = lambda blocks, size, rt: "orig-result"
    names = tapenv.install(pack, tools)
    assert set(names) == {"update_dynamic", "update_mask", "reward", "Container", "calc_positions_lb_greedy", "calc_positions_mcs"}
    # shapes beyond the compiled limits (49 cells > 32) go to the saved reference function (dataset generators, generate.py:908)
    assert tools.calc_positions_lb_greedy(np.ones((3, 3)), [7, 7, 50], "C+P+S-lb-hard") == "orig-result"
    assert pack.update_dynamic is tapenv.update_dynamic and tools.Container is tapenv.Container
    tapenv.uninstall()
    assert pack.update_mask == "orig" and tools.Container == "orig" and not hasattr(tools.calc_positions_mcs, "tapenv_original")


def test_rolling_style_container_outlives_the_window():
    """rolling.py:702-703 keeps ONE container (blocks_num = total 50, H = 250) while the network window stays at 10:
    5 windows of 10 RAND-3D blocks, masks re-initialised per window (rolling.py:325-335), state never cleared."""
    torch = _torch()
    import tapenv
    from oracle import oracle
    B, n, dim, size, total = 96, 10, 3, [5, 5, 250], 50
    static_all, dynamic_all = load_inputs("rand3d_n10.npz", 5 * B)
    env = tapenv.BatchedContainers(size, total, "C+P+S-lb-soft", "diff", batch_size=B, window=n)
    conts = [oracle.Container(size, total, "C+P+S-lb-soft", "diff") for _ in range(B)]
    rng = np.random.RandomState(4)
    for w in range(5):
        static, dynamic = static_all[w * B:(w + 1) * B], dynamic_all[w * B:(w + 1) * B]
        st, dyn = torch.from_numpy(static).cuda(), torch.from_numpy(dynamic).cuda()
        cur, mask = env.initial_mask(dyn)
        cur_o, mask_o, dyn_o = oracle.initial_mask(dynamic, n, 6), np.ones((B, 60), np.float32), dynamic
        assert np.array_equal(cur.cpu().numpy(), cur_o)
        for t in range(n):
            u = rng.random_sample((B, 60)) * (cur_o > 0)
            ptr = np.argmax(u, axis=1).astype(np.int64)
            dyn, cur, mask, dec_static, dec_dyn = env.step(torch.from_numpy(ptr).cuda(), st, dyn, mask)
            dyn_o = oracle.update_dynamic(dyn_o, static, ptr)
            cur_o, mask_o = oracle.update_mask(mask_o, dyn_o, static, ptr)
            blocks = static[np.arange(B)[:, None], 1 + np.arange(dim)[None, :], ptr[:, None]]
            enc = np.stack([conts[b].add_new_block(blocks[b]).reshape(-1) for b in range(B)])
            assert np.array_equal(dec_dyn.cpu().numpy().reshape(B, -1), enc.astype(np.float32)), (w, t)
            assert np.array_equal(cur.cpu().numpy(), cur_o) and np.array_equal(mask.cpu().numpy(), mask_o)
    assert int(env.current_blocks_num.min()) == total and (env.flags.cpu().numpy() == 0).all()
    assert np.array_equal(env.positions.cpu().numpy(), np.stack([c.positions for c in conts]))
    assert np.array_equal(env.heightmap.cpu().numpy().reshape(B, -1), np.stack([c.heightmap.reshape(-1) for c in conts]))
    want = np.array([c.calc_ratio() for c in conts])
    assert np.array_equal(env.calc_ratio().cpu().numpy(), want.astype(np.float32))
    # the stand-alone reference-style object with blocks_num = 50 (rolling.py:702)
    one = tapenv.Container(size, total, "C+P+S-lb-soft", "diff")
    ref1 = oracle.Container(size, total, "C+P+S-lb-soft", "diff")
    for i in range(12):
        blk = static_all[i, 1:4, i % 60]
        assert np.array_equal(one.add_new_block(blk), ref1.add_new_block(blk))


@pytest.mark.parametrize("fixture,size,num,strat,rt", [
    ("rand2d_n10.npz", [5, 50], 333, "LB_GREEDY", "C+P+S-lb-soft"),       # S = 20: 128-bit path
    ("rand3d_n10.npz", [5, 5, 50], 130, "LB_GREEDY", "C+P+S-lb-soft"),    # S = 60, 1800 bits per environment
    ("ppsg2d_n20.npz", [7, 50], 65, "MACS", "C+P+S-mcs-hard"),            # S = 40, 2400 bits
    ("rand2d_n10.npz", [5, 50], 40, "LB", "C+P+S-lb-soft"),               # LB keeps extra state: cleared by the plain reset
])
def test_packed_reset_expands_to_the_reference_tensors(fixture, size, num, strat, rt):
    """tapenv_reset_packed: u8 static + bit-row dynamic -> exactly the fp32 tensors PACKDataset holds, the initial masks
    of model.py:297-307, a cleared container; an episode started from it equals one started from the fp32 tensors."""
    torch = _torch()
    import tapenv
    static, dynamic = load_inputs(fixture, num)
    dim = len(size)
    n = static.shape[2] // (2 if dim == 2 else 6)
    su8, bits = tapenv.pack_inputs(static, dynamic)
    env = tapenv.BatchedContainers(size, n, rt, "diff", packing_strategy=strat, batch_size=num)
    ptr_seq = random_valid_ptrs(static, dynamic, size, seed=11)
    st, dy = torch.from_numpy(static).cuda(), torch.from_numpy(dynamic).cuda()

    def episode(start):
        s_, d_, cur, mask = start()
        cur0 = cur.clone()
        for t in range(n):
            d_, cur, mask, _, _ = env.step(torch.from_numpy(ptr_seq[t]).cuda(), s_, d_, mask)
        return cur0, env.heightmap.clone(), env.calc_ratio().clone(), env.positions.clone()

    def plain():
        cur, mask = env.reset(dy)
        return st, dy, cur, mask

    def packed():
        # dirty the state first: reset_packed must clear it
        env.step(torch.from_numpy(ptr_seq[0]).cuda(), st, dy, torch.ones(num, static.shape[2], device="cuda"))
        s_, d_, cur, mask = env.reset_packed(torch.from_numpy(su8).cuda(), torch.from_numpy(bits).cuda())
        assert torch.equal(s_, st) and torch.equal(d_, dy)
        assert bool((mask == 1).all()) and int(env.current_blocks_num.sum()) == 0 and int(env.heightmap.abs().sum()) == 0
        return s_, d_, cur, mask

    a, b = episode(plain), episode(packed)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    env.check_flags()


def test_packed_reset_scalar_path_and_errors():
    """S % 4 != 0 takes the element-wise expansion; wrong shapes are rejected before any launch."""
    torch = _torch()
    import tapenv
    rng = np.random.RandomState(0)
    B, n, R = 9, 7, 2
    S = n * R
    static = np.zeros((B, 3, S), np.float32)
    static[:, 0] = np.tile(np.arange(n), R)[None]
    static[:, 1:] = rng.randint(1, 5, size=(B, 2, S))
    dynamic = (rng.random_sample((B, 3 * n, S)) < 0.1).astype(np.float32)
    su8, bits = tapenv.pack_inputs(static, dynamic)
    env = tapenv.BatchedContainers([5, 50], n, "C+P+S-lb-soft", "diff", batch_size=B)
    s_, d_, cur, mask = env.reset_packed(torch.from_numpy(su8).cuda(), torch.from_numpy(bits).cuda())
    assert np.array_equal(s_.cpu().numpy(), static) and np.array_equal(d_.cpu().numpy(), dynamic)
    cur_ref, _ = env.reset(torch.from_numpy(dynamic).cuda())
    assert torch.equal(cur, cur_ref)
    with pytest.raises(ValueError):
        env.reset_packed(torch.from_numpy(su8).cuda(), torch.from_numpy(bits[:, :-1].copy()).cuda())
    with pytest.raises(ValueError):
        tapenv.pack_inputs(static, dynamic * 2)


@pytest.mark.parametrize("name", ["traj_2d_mulwith_lbg", "traj_2d_mul_macs", "traj_3d_mulwith_lbg", "traj_3d_mul_lbg_full"])
@pytest.mark.parametrize("fused", [True, False])
def test_two_container_inputs(name, fused):
    """input_type 'mul' / 'mul-with' through tapenv_step_mul (fused) and update_dynamic + update_mask + tapenv_add_blocks_mul
    (unfused) against the live-reference recording: both heightmaps after every step, cat(A,B) decoder input, decoder_static
    rows, masks, dynamic, positions; scores bit-equal (fp32 accumulation of model.py:503-507)."""
    torch = _torch()
    import tapenv
    t = load_traj(name)
    num = int(t["num"])
    it = str(t["input_type"])
    static = t["static"].astype(np.float32)
    _, dynamic = load_inputs(str(t["source"]), num)
    size = t["container_size"].tolist()
    dim = len(size)
    n = static.shape[2] // (2 if dim == 2 else 6)
    env = tapenv.BatchedContainerPairs(size, n, str(t["reward_type"]), str(t["heightmap_type"]),
                                       packing_strategy=str(t["packing_strategy"]), batch_size=num, input_type=it)
    st, dyn = torch.from_numpy(static).cuda(), torch.from_numpy(dynamic).cuda()
    cur, mask = env.reset(dyn)
    assert np.array_equal(cur.cpu().numpy(), t["cur_mask"][0])
    for k in range(n):
        ptr = torch.from_numpy(t["ptr"][k].astype(np.int64)).cuda()
        if fused:
            dyn, cur, mask, dec_static, dec_dyn = env.step(ptr, st, dyn, mask)
        else:
            dyn = tapenv.update_dynamic(dyn, st, ptr, it, True)
            cur, mask = tapenv.update_mask(mask, dyn, st, ptr, it, True)
            part = st[:, 1:-1] if it == "mul" else st[:, 1:]
            dec_static = torch.gather(part, 2, ptr.view(-1, 1, 1).expand(-1, part.shape[1], 1)).squeeze(2)
            tgt = torch.gather(st[:, -1], 1, ptr.view(-1, 1)).squeeze(1)
            dec_dyn = env.add_new_blocks(dec_static[:, :dim].contiguous(), tgt)
        assert np.array_equal(env.a.heightmap.cpu().numpy().reshape(num, -1), t["hm_a"][k]), k
        assert np.array_equal(env.b.heightmap.cpu().numpy().reshape(num, -1), t["hm_b"][k]), k
        assert np.array_equal(dec_dyn.cpu().numpy().reshape(num, -1), t["dec_dyn"][k]), k
        assert np.array_equal(dec_static.cpu().numpy(), t["dec_static"][k]), k
        assert np.array_equal(cur.cpu().numpy(), t["cur_mask"][k + 1]) and np.array_equal(mask.cpu().numpy(), t["mask"][k])
    assert np.array_equal(env.a.positions.cpu().numpy(), t["positions_a"])
    assert np.array_equal(env.b.positions.cpu().numpy(), t["positions_b"])
    assert np.array_equal(env.calc_ratio().cpu().numpy(), t["scores"])
    fin = np.unpackbits(t["dynamic_final"], axis=1)[:, :dynamic[0].size].reshape(dynamic.shape).astype(np.float32)
    assert np.array_equal(dyn.cpu().numpy(), fin)
    env.check_flags()
    if dim == 3:
        assert dec_dyn.shape == (num, 4 if str(t["heightmap_type"]) == "diff" else 2, size[0], size[1])


def test_two_container_bad_target_id_is_flagged():
    torch = _torch()
    import tapenv
    env = tapenv.BatchedContainerPairs([5, 50], 10, "C+P+S-lb-soft", "diff", batch_size=3, input_type="mul")
    out = env.add_new_blocks(torch.tensor([[2., 2.], [1., 3.], [2., 1.]]).cuda(), torch.tensor([0., 1., 2.]).cuda())
    assert out.shape == (3, 8)
    assert env.a.flags.cpu().tolist() == [0, 0, 4] and env.b.flags.cpu().tolist() == [0, 0, 4]
    assert env.a.heightmap.cpu().tolist()[0][:2] == [2, 2] and env.b.heightmap.cpu().tolist()[1][0] == 3
    with pytest.raises(ValueError):
        tapenv.BatchedContainerPairs([5, 50], 10, "C+P+S-lb-soft", "diff", batch_size=3, input_type="bot")


def _synthetic_inputs(rng, B, n, dim, max_edge, density):
    """Random well-formed TAP inputs of an arbitrary shape: static rows as PACKDataset lays them out (block id, edge
    lengths per rotation = permutations of the block's edges), dynamic = sparse 0/1 precedence bands."""
    import itertools
    R = 2 if dim == 2 else 6
    S = n * R
    edges = rng.randint(1, max_edge + 1, size=(B, n, dim))
    static = np.zeros((B, 1 + dim, S), np.float32)
    for r, p in enumerate(itertools.permutations(range(dim))):
        static[:, 0, r * n:(r + 1) * n] = np.arange(n)
        for d in range(dim):
            static[:, 1 + d, r * n:(r + 1) * n] = edges[:, :, p[d]]
    dynamic = (rng.random_sample((B, 3 * n, S)) < density).astype(np.float32)
    return static, dynamic


@pytest.mark.parametrize("seed", range(int(os.environ.get("TAPENV_FUZZ_SEEDS", "24"))))     # TAPENV_FUZZ_SEEDS=600 for a long run
def test_random_shapes_fuzz(seed):
    """Generic (not compile-time specialised) shapes: random n, container width/length, heightmap encoding, reward type and
    strategy; S % 4 != 0 takes the scalar tensor path.  Every step bit-exact against the oracle."""
    rng = np.random.RandomState(1000 + seed)
    dim = 2 if seed % 3 else 3
    if dim == 2:
        n = int(rng.randint(2, 25))
        size = [int(rng.randint(2, 17)), 400]
        strat = ["LB_GREEDY", "MACS", "LB_GREEDY", "LB"][seed % 4]
    else:
        n = int(rng.randint(2, 11))
        W = int(rng.randint(2, 7)); L = int(rng.randint(2, min(7, 32 // W + 1)))
        size = [W, L, 400]
        strat = ["LB_GREEDY", "LB", "MACS", "LB_GREEDY"][seed % 4]
    if strat == "MACS" and dim == 3:
        size[-1] = 200                                   # EMS coordinates are bytes in the MACS 3D kernel (height <= 255)
    if strat == "LB" and seed % 8 < 4:
        size[-1] = 130                                   # <= 256 levels: the warp form on level masks; 400: the one-thread walk
    if strat == "MACS":
        rt = ["C+P+S-mcs-soft", "C+P+S-mcs-hard", "C+P-mcs-soft", "mcs-hard"][int(rng.randint(4))]
    else:
        rt = ["C+P+S-lb-soft", "C+P+S-lb-hard", "C+P-lb-soft", "C+P-lb-hard"][int(rng.randint(4))]
    hm = ["full", "zero", "diff"][int(rng.randint(3))]
    B = int(rng.randint(1, 70))
    # MACS 3D walks the phantom (0,0,0) extents of unplaced blocks and raises IndexError when one is wider than the
    # container (tools.py:2915-2922): keep its blocks placeable
    edge_cap = min(size[:-1]) if (strat == "MACS" and dim == 3) else max(size[:-1])
    static, dynamic = _synthetic_inputs(rng, B, n, dim, max_edge=min(5, edge_cap), density=0.06)
    r = oracle_rollout(static, dynamic, size, rt, hm, strat, seed=seed)
    g = gpu_rollout(static, dynamic, r["ptr"], size, rt, hm, strat, fused=bool(seed % 2) or strat == "LB")
    if strat == "MACS" and dim == 3:
        assert (g["flags"] == 0).all()
    assert_same(g, r, r["ptr"], static, dim)


@pytest.mark.parametrize("fixture,size,strat,rt,B", [("rand2d_n10.npz", [5, 50], "LB_GREEDY", "C+P+S-lb-soft", 65536),
                                                     ("rand3d_n10.npz", [5, 5, 50], "LB_GREEDY", "C+P+S-lb-hard", 32768),
                                                     ("ppsg2d_n20.npz", [7, 100], "MACS", "C+P+S-mcs-hard", 8192)])
def test_full_size_step_properties(fixture, size, strat, rt, B):
    """Batches at and beyond BASELINE's sizes through size-independent properties of the transition: each block chosen
    once, masks only ever lose candidates, `dynamic` only ever loses entries, sum(heightmap) == valid + empty, valid ==
    volume of the placed blocks, and tiled copies of the fixture pool under identical pointers agree bit for bit (the
    environments are independent) -- plus the oracle on the pool itself."""
    torch = _torch()
    import tapenv
    static_p, dynamic_p = load_inputs(fixture)
    pool = min(static_p.shape[0], 512)
    static_p, dynamic_p = static_p[:pool], dynamic_p[:pool]
    dim = len(size)
    S = static_p.shape[2]
    n = S // (2 if dim == 2 else 6)
    reps = B // pool
    st = torch.from_numpy(static_p).cuda().repeat(reps, 1, 1)
    dyn = torch.from_numpy(dynamic_p).cuda().repeat(reps, 1, 1)
    env = tapenv.BatchedContainers(size, n, rt, "diff", packing_strategy=strat, batch_size=B)
    cur, mask = env.reset(dyn)
    weights = torch.arange(S, 0, -1, device="cuda", dtype=torch.float32)
    ptrs = []
    for t in range(n):
        ptr = torch.argmax(cur[:pool] * weights, dim=1).repeat(reps)                      # first accessible candidate
        prev_dyn, prev_mask = dyn, mask
        dyn, cur, mask, dec_static, dec_dyn = env.step(ptr, st, dyn, mask)
        assert bool((dyn <= prev_dyn).all()) and bool((mask <= prev_mask).all()) and bool((cur <= mask).all())
        ptrs.append(ptr)
    tour = torch.stack(ptrs, 1)
    assert bool((torch.sort(tour % n, dim=1).values == torch.arange(n, device="cuda")).all())
    assert float(mask.sum()) == 0.0 and float(dyn[:, :n].sum()) == 0.0
    hm = env.heightmap.reshape(B, -1).to(torch.int64)
    sc = env.scalars.to(torch.int64)
    assert bool((hm.sum(1) == sc[:, 0] + sc[:, 1]).all())
    placed_vol = (env.blocks.to(torch.int64).prod(2) * (env.stable.to(torch.int64) >= 0)).sum(1)
    if not rt.endswith("hard"):
        assert bool((sc[:, 0] == placed_vol).all())
    else:
        assert bool((sc[:, 0] <= placed_vol).all())
    r = env.calc_ratio()
    assert bool(torch.equal(hm.view(reps, pool, -1), hm[:pool].expand(reps, pool, -1)))
    assert bool(torch.equal(r.view(reps, pool), r[:pool].expand(reps, pool)))
    o = oracle_rollout(static_p, dynamic_p, size, rt, "diff", strat, ptr_seq=tour[:pool].T.cpu().numpy())
    assert np.array_equal(hm[:pool].cpu().numpy(), o["heightmap"][-1])
    assert np.array_equal(r[:pool].cpu().numpy(), o["ratio"].astype(np.float32), equal_nan=True)
    env.check_flags()


@pytest.mark.parametrize("name", ["bot2d", "mul2d", "mul3d"])
def test_pack_reward_known_answers(name):
    """tapenv.reward (pack.reward, pack.py:378-473 -- the deprecated re-packing reward) against values recorded from the live
    reference, including the two-container split of 'mul' / 'mul-with' (an empty part scores 0)."""
    torch = _torch()
    import tapenv
    z = np.load(golden_path("reward_kat.npz"))
    rt, it, W, H = [str(v) for v in z[name + "_args"]]
    static = torch.from_numpy(z[name + "_static"].astype(np.float32)).cuda()
    tour = torch.from_numpy(z[name + "_tour"]).cuda()
    r = tapenv.reward(static, tour, rt, it, True, int(W), int(H))
    assert r.dtype == torch.float32 and r.is_cuda
    assert np.abs(r.cpu().numpy().astype(np.float64) - z[name + "_reward"].astype(np.float64)).max() <= 1e-6


@pytest.mark.parametrize("packed", [False, True])
@pytest.mark.parametrize("use_graph", [False, True])
def test_host_pipeline_uploads(packed, use_graph):
    """tapenv.HostPipeline: host buffers in, host rewards out -- three separate pinned tensors or one contiguous HostBatch,
    fp32 or packed; more episodes than the pipeline is deep; every result equals the oracle episode for ITS inputs."""
    torch = _torch()
    import tapenv
    from oracle import oracle
    B, n, size = 300, 10, [5, 50]
    static, dynamic = load_inputs("rand2d_n10.npz", 4 * B)
    env = tapenv.BatchedContainers(size, n, "C+P+S-lb-soft", "diff", batch_size=B)
    pipe = tapenv.HostPipeline(env, n, depth=3, use_graph=use_graph, packed=packed)
    batches, want = [], []
    for i in range(4):
        st, dy = static[i * B:(i + 1) * B], dynamic[i * B:(i + 1) * B]
        ptrs = random_valid_ptrs(st, dy, size, seed=20 + i)
        want.append(oracle.episode_batch(st, dy, ptrs, size, "C+P+S-lb-soft", "diff", "LB_GREEDY", want=("reward",))["reward"])
        a, b = tapenv.pack_inputs(st, dy) if packed else (st, dy)
        if i % 2 == 0:                                     # one contiguous pinned batch
            hb = pipe.new_host_batch()
            hb.static.copy_(torch.from_numpy(a).view_as(hb.static)); hb.dynamic.copy_(torch.from_numpy(b).view_as(hb.dynamic))
            hb.ptr.copy_(torch.from_numpy(ptrs).view_as(hb.ptr))
            batches.append((hb,))
        else:                                              # three pinned tensors
            batches.append((torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory(), torch.from_numpy(ptrs).pin_memory()))
    got = []
    order = [0, 1, 2, 3, 1, 0]
    for i in order:
        if pipe.inflight == pipe.depth:
            got.append(pipe.result()[0].clone())
        pipe.submit(*batches[i])
    while pipe.inflight:
        r, sums = pipe.result()
        got.append(r.clone())
        assert float(sums[2]) == B
    assert len(got) == len(order)
    for i, r in zip(order, got):
        assert np.array_equal(r.numpy(), want[i]), i


@pytest.mark.parametrize("fixture,size,strat,rt", [("rand2d_n10.npz", [5, 50], "LB_GREEDY", "C+P+S-lb-soft"),
                                                   ("rand3d_n10.npz", [5, 5, 50], "LB_GREEDY", "C+P+S-lb-hard"),
                                                   ("ppsg2d_n20.npz", [7, 100], "MACS", "C+P+S-mcs-hard"),
                                                   ("rand2d_n10.npz", [5, 50], "LB", "C+P+S-lb-soft"),
                                                   ("rand3d_n10.npz", [5, 5, 50], "MACS", "C+P+S-mcs-soft")])
def test_reward_rides_on_the_last_step(fixture, size, strat, rt):
    """tapenv_step_reward: calc_ratio emitted by the decode step itself (any step, not only the last) equals the separate
    tapenv_reward launch bit for bit; tapenv_reward_sums equals the sums tapenv_reward produces."""
    torch = _torch()
    import tapenv
    B = 257
    static, dynamic = load_inputs(fixture, B)
    n = static.shape[2] // (2 if len(size) == 2 else 6)
    ptrs = random_valid_ptrs(static, dynamic, size, seed=4)
    env = tapenv.BatchedContainers(size, n, rt, "diff", packing_strategy=strat, batch_size=B)
    st, dyn = torch.from_numpy(static).cuda(), torch.from_numpy(dynamic).cuda()
    cur, mask = env.reset(dyn)
    rbuf = torch.full((B,), -7.0, device="cuda")
    for t in range(n):
        dyn, cur, mask, _, _ = env.step(torch.from_numpy(ptrs[t]).cuda(), st, dyn, mask, reward_out=rbuf if t in (2, n - 1) else None)
        if t in (2, n - 1):
            want, sums = env.calc_ratio(partial_sums=True)
            assert torch.equal(rbuf, want), t
            assert torch.equal(env.reward_sums(rbuf), sums)
    env.check_flags()


@pytest.mark.parametrize("dim,fixture,size", [(2, "rand2d_n10.npz", [5, 50]), (3, "rand3d_n10.npz", [5, 5, 50])])
def test_g6_pretrained_network_tours_on_gpu(dim, fixture, size):
    """G6: the tours the unmodified reference network (pretrained actor, greedy) chose, replayed through the fused step, the
    whole-episode kernel and the decode loop driven by a tour-replaying actor: rewards equal the ones DRL.forward returned."""
    torch = _torch()
    import tapenv
    z = np.load(golden_path("g6_tours.npz"))
    num = int(z["g6_%dd_num" % dim])
    static, dynamic = load_inputs(fixture, num)
    tour = z["g6_%dd_tour" % dim]
    want = -z["g6_%dd_reward" % dim].astype(np.float64)
    env = tapenv.BatchedContainers(size, 10, "C+P+S-lb-soft", "diff", batch_size=num)
    st, dyn = torch.from_numpy(static).cuda(), torch.from_numpy(dynamic).cuda()
    tq = torch.from_numpy(tour.T.copy()).cuda()
    run = tapenv.EpisodeRunner(env, st, dyn, tq, use_graph=False)
    assert np.abs(run.run().cpu().numpy().astype(np.float64) - want).max() <= 1e-6
    assert np.abs(env.episode(st, dyn, tq)[0].cpu().numpy().astype(np.float64) - want).max() <= 1e-6
    S = static.shape[2]

    def replay_actor(static_, dynamic_, dec_static, dec_dyn, state):          # logits that make the greedy loop follow the recorded tour
        t = 0 if state is None else state
        return torch.nn.functional.one_hot(tq[t], S).float() * 50.0, t + 1

    tour_idx, _, reward = tapenv.DecodeLoop(env, replay_actor, greedy=True).run(st, dyn)
    assert np.array_equal(tour_idx.cpu().numpy(), tour)
    assert np.abs(reward.cpu().numpy().astype(np.float64) - want).max() <= 1e-6


@pytest.mark.parametrize("W,n,rt", [(7, 48, "C+P+S-mcs-hard"), (5, 60, "C+P+S-mcs-soft"), (12, 40, "mcs-hard"), (3, 64, "C+P-mcs-hard")])
def test_macs_long_histories_two_slots(W, n, rt):
    """MACS 2D list B (tools.py:2531-2555) over more than 32 previous blocks (second history slot), many identical block tops
    (the duplicate test of :2538) and unplaced phantom entries: placement-only path against the oracle, every step."""
    torch = _torch()
    import tapenv
    from oracle import oracle
    B = 40
    rng = np.random.RandomState(W * 1000 + n)
    H = 4 * n + 8
    env = tapenv.BatchedContainers([W, H], n, rt, "diff", packing_strategy="MACS", batch_size=B)
    conts = [oracle.Container([W, H], n, rt, "diff", packing_strategy="MACS") for _ in range(B)]
    for t in range(n):
        # few distinct sizes -> many equal tops; with -hard, blocks that find no stable position stay unplaced (phantom entries)
        blocks = rng.randint(1, min(4, W + 1), size=(B, 2)).astype(np.float32)
        enc = env.add_new_blocks(torch.from_numpy(blocks).cuda()).cpu().numpy()
        want = np.stack([np.asarray(conts[b].add_new_block(blocks[b])).reshape(-1) for b in range(B)])
        assert np.array_equal(enc, want.astype(np.float32)), t
        assert np.array_equal(env.heightmap.cpu().numpy(), np.stack([c.heightmap for c in conts])), t
    assert np.array_equal(env.positions.cpu().numpy(), np.stack([c.positions for c in conts]))
    assert np.array_equal(env.stable.cpu().numpy(), np.stack([np.array(c.stable, dtype=np.uint8) for c in conts]))
    r = env.calc_ratio().cpu().numpy().astype(np.float64)
    want_r = np.array([c.calc_ratio() for c in conts])
    assert np.abs(r - want_r).max() <= REWARD_TOL


@pytest.mark.parametrize("fixture,size,rt,strat,it", [("rand2d_n10.npz", [5, 50], "C+P+S-lb-soft", "LB", "mul-with"),
                                                      ("rand3d_n10.npz", [5, 5, 50], "C+P+S-lb-hard", "LB", "mul"),
                                                      ("rand3d_n10.npz", [5, 5, 50], "C+P+S-mcs-soft", "MACS", "mul-with")])
def test_two_container_inputs_voxel_strategies(fixture, size, rt, strat, it):
    """tapenv_step_mul for the voxel-state strategies (LB, MACS 3D): r01 returned TAPENV_EUNSUPPORTED; against the oracle's
    two-container rollout (tests/rollout.py:oracle_rollout_mul), every step."""
    torch = _torch()
    import tapenv
    from tests.rollout import oracle_rollout_mul
    B = 96
    static, dynamic = load_inputs(fixture, B)
    dim = len(size)
    R = 2 if dim == 2 else 6
    n = static.shape[2] // R
    rng = np.random.RandomState(5)
    ids = np.tile(rng.randint(0, 2, size=(B, 1, n)).astype(np.float32), (1, 1, R))
    static = np.ascontiguousarray(np.concatenate([static, ids], 1))
    ptrs = random_valid_ptrs(static[:, :1 + dim], dynamic, size, seed=9)
    want = oracle_rollout_mul(static, dynamic, ptrs, size, rt, "diff", strat, it)
    env = tapenv.BatchedContainerPairs(size, n, rt, "diff", packing_strategy=strat, batch_size=B, input_type=it)
    st, dyn = torch.from_numpy(static).cuda(), torch.from_numpy(dynamic).cuda()
    cur, mask = env.reset(dyn)
    for k in range(n):
        dyn, cur, mask, dec_static, dec_dyn = env.step(torch.from_numpy(ptrs[k]).cuda(), st, dyn, mask)
        assert np.array_equal(env.a.heightmap.cpu().numpy().reshape(B, -1), want["hm_a"][k]), k
        assert np.array_equal(env.b.heightmap.cpu().numpy().reshape(B, -1), want["hm_b"][k]), k
        assert np.array_equal(dec_dyn.cpu().numpy().reshape(B, -1), want["dec_dyn"][k].astype(np.float32)), k
        assert np.array_equal(dec_static.cpu().numpy(), want["dec_static"][k]), k
        assert np.array_equal(cur.cpu().numpy(), want["cur_mask"][k + 1]) and np.array_equal(mask.cpu().numpy(), want["mask"][k])
    assert np.array_equal(env.a.positions.cpu().numpy(), want["positions_a"])
    assert np.array_equal(env.b.positions.cpu().numpy(), want["positions_b"])
    assert np.array_equal(env.calc_ratio().cpu().numpy(), want["scores"])
    env.check_flags()


@pytest.mark.parametrize("input_type,allow_rot", [("simple", False), ("simple", True), ("rot", True), ("rot-old", True), ("bot", True),
                                                  ("bot-rot", True), ("use-static", True), ("use-pnet", True), ("mul", True),
                                                  ("mul-with", True)])
def test_tensor_ops_every_input_type(input_type, allow_rot):
    """tapenv.update_dynamic / update_mask / the initial masks / the packed upload for EVERY input_type string of
    pack.py:285-309 against the oracle (itself checked against the live pack.* functions for the same layouts in
    tests/test_oracle_vs_reference.py) -- including the legacy 'rot-old' layout (n movement rows + one rotate-state row), whose
    FUSED step is refused like the reference's own decode loop (model.py:391-392 -> tools.py:2060 raises)."""
    torch = _torch()
    import tapenv
    from oracle import oracle
    dev = torch.device("cuda:0")
    rng = np.random.RandomState(21)
    for dim, n in ((2, 10), (3, 5), (2, 7)):                 # S % 4 != 0 at (2, 7): scalar tensor path for every layout
        R = (2 if dim == 2 else 6) if allow_rot else 1
        S, B = n * R, 37
        srows = 2 + dim if input_type in ("mul", "mul-with") else 1 + dim
        drows = n if input_type in ("simple", "rot") else (n + 1 if input_type == "rot-old" else 3 * n)
        static = np.zeros((B, srows, S), np.float32)
        static[:, 0] = np.tile(np.arange(n), R)
        static[:, 1:1 + dim] = rng.randint(1, 5, size=(B, dim, S))
        dynamic = (rng.random_sample((B, drows, S)) < 0.08).astype(np.float32)
        size = [5, 50] if dim == 2 else [5, 5, 50]
        if input_type in ("mul", "mul-with"):
            static[:, -1] = rng.randint(0, 2, size=(B, S))
            env = tapenv.BatchedContainerPairs(size, n, "C+P+S-lb-soft", "diff", batch_size=B, device=dev, input_type=input_type,
                                               allow_rot=allow_rot).a
        else:
            env = tapenv.BatchedContainers(size, n, "C+P+S-lb-soft", "diff", batch_size=B, device=dev, input_type=input_type,
                                           allow_rot=allow_rot)
        st, dyn = torch.from_numpy(static).to(dev), torch.from_numpy(dynamic).to(dev)
        cur0, mask0 = env.reset(dyn)
        assert np.array_equal(cur0.cpu().numpy(), oracle.initial_mask(dynamic, n, R)) and bool((mask0 == 1).all())
        # packed upload of the same layout
        su8, bits = tapenv.pack_inputs(static, dynamic)
        st2, dyn2, cur2, _ = env.reset_packed(torch.from_numpy(su8).to(dev), torch.from_numpy(bits).to(dev))
        assert torch.equal(st2, st) and torch.equal(dyn2, dyn) and torch.equal(cur2, cur0)
        mask = np.ones((B, S), np.float32)
        mask_t = torch.from_numpy(mask).to(dev)
        for t in range(n):
            ptr = rng.randint(0, S, size=B).astype(np.int64)
            dynamic = oracle.update_dynamic(dynamic, static, ptr, input_type, allow_rot)
            cur, mask = oracle.update_mask(mask, dynamic, static, ptr, input_type, allow_rot)
            ptr_t = torch.from_numpy(ptr).to(dev)
            dyn = tapenv.update_dynamic(dyn, st, ptr_t, input_type, allow_rot)
            cur_t, mask_t = tapenv.update_mask(mask_t, dyn, st, ptr_t, input_type, allow_rot)
            assert np.array_equal(dyn.cpu().numpy(), dynamic)
            assert np.array_equal(cur_t.cpu().numpy(), cur) and np.array_equal(mask_t.cpu().numpy(), mask)
        if input_type == "rot-old":
            with pytest.raises(tapenv.TapEnvError) as e:
                env.step(ptr_t, st, dyn, mask_t)
            assert e.value.code == tapenv._capi.EUNSUPPORTED


@pytest.mark.parametrize("size,strategy,rt", [([5, 50], "LB_GREEDY", "C+P+S-lb-hard"), ([5, 5, 50], "LB_GREEDY", "C+P+S-lb-soft"),
                                              ([7, 60], "MACS", "C+P+S-mcs-hard"), ([5, 5, 50], "MACS", "C+P+S-mcs-soft"),
                                              ([5, 50], "LB", "C+P+S-lb-soft"), ([5, 5, 50], "LB", "C+P+S-lb-hard")])
def test_container_proxy_attributes_match_reference_container(size, strategy, rt):
    """Every attribute of tools.Container a caller can read (tools.py:3628-3661; rolling.py:640-658, draw_container
    :3968-3996): heightmap, positions, stable, valid / empty size, current_blocks_num, blocks, container (the voxel grid --
    rebuilt from the placements where the kernels keep only the heightmap, read from the state for LB / MACS 3D),
    bounding_box, calc_CPS, get_heightmap -- per-object proxies (each its own batch of one) against the oracle."""
    import tapenv
    from oracle import oracle
    dim = len(size)
    rng = np.random.RandomState(17)
    n = 12
    for ep in range(4):
        hi = min(5, size[0]) + 1
        blocks = np.stack([rng.randint(1, hi, size=n) for _ in range(dim - 1)] + [rng.randint(1, 6, size=n)], 1).astype(np.float32)
        ours = tapenv.Container(size, n, rt, "diff", [7, 7, 50][:dim] if dim == 3 else [7, 50], packing_strategy=strategy)
        ref = oracle.Container(size, n, rt, "diff", packing_strategy=strategy)
        for t in range(n):
            enc = ours.add_new_block(blocks[t], False)
            want = ref.add_new_block(blocks[t], False)
            assert np.array_equal(np.asarray(enc), np.asarray(want))
            assert np.array_equal(ours.get_heightmap(), ref.get_heightmap())
        assert np.array_equal(ours.heightmap, ref.heightmap) and np.array_equal(ours.positions, ref.positions)
        assert ours.stable == ref.stable and ours.valid_size == ref.valid_size and ours.empty_size == ref.empty_size
        assert ours.current_blocks_num == n and np.array_equal(np.array(ours.blocks), blocks.astype(int))
        assert np.array_equal(ours.container, ref.container)
        assert np.array_equal(ours.bounding_box, np.zeros(dim)) and ours.rotate_state == [False] * n
        assert np.allclose(ours.calc_CPS(), ref.calc_CPS(), rtol=0, atol=1e-12)
        assert abs(ours.calc_ratio() - ref.calc_ratio()) <= REWARD_TOL


@pytest.mark.parametrize("size,hm,strategy,rt,with_id", [([5, 50], "diff", "LB_GREEDY", "C+P+S-lb-soft", True), ([5, 50], "full", "LB_GREEDY", "C+P+S-lb-hard", False),
                                                         ([5, 5, 50], "diff", "LB_GREEDY", "C+P+S-lb-soft", True), ([5, 5, 50], "zero", "LB_GREEDY", "C+P+S-lb-soft", False),
                                                         ([7, 60], "diff", "MACS", "C+P+S-mcs-soft", True), ([5, 5, 50], "diff", "MACS", "C+P+S-mcs-hard", False)])
def test_two_container_lists_as_unmodified_model_py_drives_them(size, hm, strategy, rt, with_id):
    """model.py:291-292 builds TWO lists of B containers for input_type 'mul' / 'mul-with'; per decode step and environment
    the block goes to one list's container while the other one is only asked for its heightmap (model.py:414-428, verbatim
    below).  The per-object proxies recognise the pattern (2B objects, rows of one [B,dim(+1)] array) and serve a step with ONE
    launch: returned heightmaps (after the loop, as model.py consumes them), final states and calc_ratio against two oracle
    containers per environment."""
    import tapenv
    from oracle import oracle
    from tapenv import containers as tc
    dim = len(size)
    B, n = 24, 8
    rng = np.random.RandomState(5 + dim)
    containers_a = [tapenv.Container(size, n, rt, hm, packing_strategy=strategy) for _ in range(B)]
    containers_b = [tapenv.Container(size, n, rt, hm, packing_strategy=strategy) for _ in range(B)]
    ref_a = [oracle.Container(size, n, rt, hm, packing_strategy=strategy) for _ in range(B)]
    ref_b = [oracle.Container(size, n, rt, hm, packing_strategy=strategy) for _ in range(B)]
    for t in range(n):
        hi = min(4, size[0]) + 1
        full = np.concatenate([rng.randint(1, hi, size=(B, dim)), rng.randint(0, 2, size=(B, 1))], 1).astype(np.float32)
        target_ids = full[:, -1].copy()
        blocks = full[:, :dim] if with_id else np.ascontiguousarray(full[:, :dim])      # 'mul-with' slices, 'mul' owns its array
        is_rotate = np.zeros(B, bool)
        heightmaps_a, heightmaps_b, want_a, want_b = [], [], [], []
        for batch_index in range(B):                                                   # model.py:418-428
            target_id = target_ids[batch_index]
            if target_id == 0:
                heightmaps_a.append(containers_a[batch_index].add_new_block(blocks[batch_index], is_rotate[batch_index]))
                heightmaps_b.append(containers_b[batch_index].get_heightmap())
                want_a.append(ref_a[batch_index].add_new_block(blocks[batch_index])); want_b.append(ref_b[batch_index].get_heightmap())
            elif target_id == 1:
                heightmaps_a.append(containers_a[batch_index].get_heightmap())
                heightmaps_b.append(containers_b[batch_index].add_new_block(blocks[batch_index], is_rotate[batch_index]))
                want_a.append(ref_a[batch_index].get_heightmap()); want_b.append(ref_b[batch_index].add_new_block(blocks[batch_index]))
        assert np.array_equal(np.array(heightmaps_a), np.array(want_a)) and np.array_equal(np.array(heightmaps_b), np.array(want_b))
    assert containers_a[0]._pair is not None and containers_a[0]._pair is containers_b[B - 1]._pair      # ONE batched pair, not 2B singles
    assert isinstance(containers_a[0]._group.batch, tc.BatchedContainerPairs)
    for b in range(B):                                                                 # model.py:503-507
        assert abs(containers_a[b].calc_ratio() - ref_a[b].calc_ratio()) <= REWARD_TOL
        assert abs(containers_b[b].calc_ratio() - ref_b[b].calc_ratio()) <= REWARD_TOL
    assert np.array_equal(containers_b[3].positions, ref_b[3].positions) and containers_a[5].stable == ref_a[5].stable
    # out-of-protocol use is refused loudly instead of returning heightmaps that would never be filled in
    containers_a[0].add_new_block(blocks[0])
    with pytest.raises(RuntimeError):
        containers_a[2].add_new_block(blocks[2])


@pytest.mark.parametrize("WL", [(8, 4), (4, 8), (16, 2), (2, 16), (32, 1), (1, 32)])
@pytest.mark.parametrize("strat,rt", [("LB_GREEDY", "C+P+S-lb-hard"), ("LB", "C+P+S-lb-soft"), ("MACS", "C+P+S-mcs-soft")])
def test_full_32_cell_3d_containers(WL, strat, rt):
    """The compiled limit of the 3D kernels: W*L = 32 heightmap cells = one lane / one mask bit per cell, bit 31 included
    (8x4, 4x8, 16x2, 2x16 and the degenerate 32x1 / 1x32), every step against the oracle -- fused and unfused."""
    W, L = WL
    rng = np.random.RandomState(W * 100 + L)
    size = [W, L, 120]
    n, B = 6, 12
    static, dynamic = _synthetic_inputs(rng, B, n, 3, max_edge=min(4, W, L), density=0.05)
    r = oracle_rollout(static, dynamic, size, rt, "diff", strat, seed=3)
    for fused in (True, False):
        g = gpu_rollout(static, dynamic, r["ptr"], size, rt, "diff", strat, fused=fused)
        assert (g["flags"] == 0).all()
        assert_same(g, r, r["ptr"], static, 3)


@pytest.mark.parametrize("strat,rt", [("LB_GREEDY", "C+P+S-lb-soft"), ("LB", "C+P+S-lb-hard"), ("MACS", "C+P+S-mcs-hard")])
def test_full_32_column_2d_containers(strat, rt):
    """The compiled limit in 2D: W = 32 columns = every lane / every bit of a level mask in use."""
    rng = np.random.RandomState(32)
    size = [32, 120]
    n, B = 12, 16
    static, dynamic = _synthetic_inputs(rng, B, n, 2, max_edge=6, density=0.05)
    r = oracle_rollout(static, dynamic, size, rt, "zero", strat, seed=4)
    for fused in (True, False):
        g = gpu_rollout(static, dynamic, r["ptr"], size, rt, "zero", strat, fused=fused)
        assert (g["flags"] == 0).all()
        assert_same(g, r, r["ptr"], static, 2)


@pytest.mark.parametrize("strat,rt", [("LB_GREEDY", "C+P+S-lb-hard"), ("LB", "C+P+S-lb-soft"), ("MACS", "C+P+S-mcs-soft")])
def test_full_64_candidate_masks(strat, rt):
    """The compiled limit of the tensor pass: S = n * R = 64 candidates (n = 32 blocks, 2 rotations) = every bit of the 64-bit
    accessibility words in use; every step (masks, dynamic, placements) against the oracle."""
    rng = np.random.RandomState(64)
    size = [9, 250]
    n, B = 32, 10
    static, dynamic = _synthetic_inputs(rng, B, n, 2, max_edge=4, density=0.02)
    r = oracle_rollout(static, dynamic, size, rt, "diff", strat, seed=6)
    for fused in (True, False):
        g = gpu_rollout(static, dynamic, r["ptr"], size, rt, "diff", strat, fused=fused)
        assert (g["flags"] == 0).all()
        assert_same(g, r, r["ptr"], static, 2)
