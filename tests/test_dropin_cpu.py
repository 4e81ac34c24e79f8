"""CPU: host-side pieces of the drop-in layer against the LIVE reference (needs the reference tree: /root/reference in the
build container, oracle/_ref on the GPU box) -- the voxel grid calc_positions_* hand to generate.calc_dependent, and the
network adapter tapenv.adapters.drl_actor_step against model.DRL.forward itself."""
import contextlib
import io

import numpy as np
import pytest

from oracle import refshim
from tests.golden_io import load_inputs

pytestmark = pytest.mark.skipif(not refshim.available(), reason="reference tree not present")


@pytest.mark.parametrize("dim,size,rt,fn", [(2, [7, 60], "C+P+S-lb-hard", "calc_positions_lb_greedy"),
                                            (2, [5, 60], "C+P+S-lb-soft", "calc_positions_lb_greedy"),
                                            (3, [5, 5, 60], "C+P+S-lb-hard", "calc_positions_lb_greedy"),
                                            (3, [7, 7, 60], "C+P+S-lb-soft", "calc_positions_lb_greedy"),
                                            (2, [7, 60], "C+P+S-mcs-hard", "calc_positions_mcs"),
                                            (2, [5, 60], "C+P+S-mcs-soft", "calc_positions_mcs")])
def test_voxel_container_rebuilt_from_placements(dim, size, rt, fn):
    """tools.calc_positions_*'s second return value (tools.py:2168-2169, :2661-2665) from (positions, blocks) alone."""
    from tapenv.episode import voxel_container
    tools = refshim.load(("tools",))["tools"]
    rng = np.random.RandomState(dim * 100 + len(rt))
    unplaced = 0
    for _ in range(40):
        blocks = rng.randint(1, 5, size=(10, dim))
        positions, container, stable, ratio, scores = getattr(tools, fn)(blocks.copy(), size, rt)
        unplaced += int(rt.endswith("hard")) * (10 - int(np.sum(stable)))
        got = voxel_container(positions, blocks, size)
        assert np.array_equal(got, container)
    if rt.endswith("hard") and "lb" in rt:
        assert unplaced > 0                      # the unplaced-block branch was exercised


@pytest.mark.parametrize("dim,ck", [(2, "2d-bot-C+P+S-lb-soft-width-5-note-sh-R-diff"),
                                    (3, "3d-bot-C+P+S-lb-soft-width-5-note-sh-R-diff")])
@pytest.mark.parametrize("incremental", [False, True])
def test_actor_step_adapter_reproduces_drl_forward(dim, ck, incremental):
    """The network half lifted out of model.DRL (tapenv.adapters.drl_actor_step) + the reference's own environment,
    driven like tapenv.DecodeLoop drives it == model.DRL.forward (model.py:254-515): same tours, same log-probabilities."""
    import torch
    from tapenv.adapters import drl_actor_step
    mods = refshim.load(("tools", "generate", "pack", "model"))
    pack, model, tools = mods["pack"], mods["model"], mods["tools"]
    B, n = 12, 10
    R = 2 if dim == 2 else 6
    static, dynamic = load_inputs("rand%dd_n10.npz" % dim, B)
    with contextlib.redirect_stdout(io.StringIO()):
        actor = model.DRL(dim, 30, 128, 256, False, "bot", True, 5, 50, dim, "C+P+S-lb-soft", "shape_heightmap", "diff",
                          "LB_GREEDY", pack.update_dynamic, pack.update_mask, 1, 0.1, 1.0)
    actor.load_state_dict(torch.load(refshim.checkpoint(ck), map_location="cpu"))
    actor.eval()
    st, dy = torch.from_numpy(static), torch.from_numpy(dynamic)
    dec_static = torch.zeros(B, dim, 1)
    dec_dyn = torch.zeros(B, 4, 1) if dim == 2 else torch.zeros(B, 2, 5, 5)
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        want_idx, want_logp, _, want_r = actor(st, dy, [dec_static, dec_dyn])

    step = drl_actor_step(actor, incremental=incremental)
    size = [5, 50] if dim == 2 else [5, 5, 50]
    conts = [tools.Container(size, n, "C+P+S-lb-soft", "diff", packing_strategy="LB_GREEDY") for _ in range(B)]
    mask = torch.ones(B, n * R)
    cur = mask.clone()
    blocked = dy[:, :n].sum(1) + dy[:, n:2 * n].sum(1) * dy[:, 2 * n:].sum(1)
    cur[blocked.ne(0)] = 0.
    d_s = torch.zeros(B, dim)
    d_d = torch.zeros(B, 4) if dim == 2 else torch.zeros(B, 2, 5, 5)
    state, dyn = None, dy
    idx, logps = [], []
    with torch.no_grad():
        for _ in range(n):
            logits, state = step(st, dyn, d_s, d_d, state)
            probs = torch.softmax(logits + cur.log(), dim=1)
            prob, ptr = torch.max(probs, 1)
            if incremental:                         # teacher-forced along the reference's tour: a 1e-6 change of the hidden
                ptr = want_idx[:, len(idx)]         # state may flip an argmax between two near-equal candidates
                prob = torch.gather(probs, 1, ptr.view(-1, 1)).squeeze(1)
            dyn = pack.update_dynamic(dyn, st, ptr, "bot", True)
            cur, mask = pack.update_mask(mask, dyn, st, ptr, "bot", True)
            state.last_ptr = ptr
            d_s = torch.gather(st[:, 1:], 2, ptr.view(-1, 1, 1).expand(-1, dim, 1)).squeeze(2)
            hms = [conts[b].add_new_block(d_s[b].numpy()) for b in range(B)]
            d_d = torch.FloatTensor(np.array(hms))
            idx.append(ptr); logps.append(prob.log())
    got_idx, got_logp = torch.stack(idx, 1), torch.stack(logps, 1)
    assert torch.equal(got_idx, want_idx)
    # full recomputation: the same torch ops in the same order -> bit-equal; incremental: fp32 rounding of a rank-3 update
    # accumulated over 10 steps (measured 1e-6 on the hidden state per step) -> 1e-4 on the log-probabilities
    tol = 0.0 if not incremental else 1e-4
    assert float((got_logp - want_logp).abs().max()) <= tol
    r = torch.tensor([c.calc_ratio() for c in conts], dtype=torch.float32)
    assert torch.equal(-r, want_r)


@pytest.mark.parametrize("dim,W", [(2, 7), (2, 0), (3, 5)])
def test_speculative_generate_blocks_host_logic(dim, W, monkeypatch):
    """tapenv.generators.generate_blocks (the rejection loop of generate.py:896-910 evaluated K draws at a time): the HOST
    logic -- speculation, rewinding np.random to exactly where the sequential loop stops, the rotation / dependency tail --
    with the batched GPU evaluation replaced by the reference's own calc_positions_lb_greedy, candidate by candidate.  Same
    outputs and same stream position as generate.generate_blocks under the same seed (the GPU evaluation itself is covered by
    tests/test_gpu_model_in_loop.py::test_create_dataset_identical_with_batched_generator)."""
    import torch
    from tapenv import generators
    mods = refshim.load(("tools", "generate"))
    tools, generate = mods["tools"], mods["generate"]
    calls = {"batches": 0, "cands": 0}

    def fake_calc(blocks, container_size, reward_type):
        blocks = np.asarray(blocks)
        if blocks.ndim == 2:
            return tools.calc_positions_lb_greedy(blocks.copy(), container_size, reward_type)
        calls["batches"] += 1
        calls["cands"] += len(blocks)
        res = [tools.calc_positions_lb_greedy(b.copy(), container_size, reward_type) for b in blocks]
        pos = torch.from_numpy(np.stack([r[0] for r in res]))
        stable = torch.tensor([[bool(v) for v in r[2]] for r in res])
        return pos, None, stable, None, None

    monkeypatch.setattr(generators, "calc_positions_lb_greedy", fake_calc)
    size = [W, 50] if dim == 2 else [W, W, 50]
    for seed in (3, 4, 5):
        np.random.seed(seed)
        want = [generate.generate_blocks(10, list(size), 1, [1, 5]) for _ in range(3)]
        want_next = np.random.random_sample(4)
        np.random.seed(seed)
        got = [generators.generate_blocks(10, list(size), 1, [1, 5], speculate=4) for _ in range(3)]
        got_next = np.random.random_sample(4)
        for a, b in zip(got, want):
            for x, y in zip(a, b):
                assert np.array_equal(np.asarray(x), np.asarray(y))
        assert np.array_equal(got_next, want_next)           # the global stream is where the sequential loop leaves it
    if W > 0:                                                # (with a random width the first, non-speculative draw usually fits)
        assert calls["batches"] > 0 and calls["cands"] >= calls["batches"]


@pytest.mark.parametrize("input_type", ["bot", "simple"])
def test_speculative_ppsg_generator_host_logic(input_type, monkeypatch):
    """tapenv.generators.generate_blocks_with_GT (the PPSG generator, generate.py:17-229: 20 unpacking orders of a perfect
    packing, each re-packed with calc_positions_lb_greedy, :112): the HOST logic -- the 20 tries' draws made up front, the
    interval-arithmetic calc_dependent on the perfect packing and on the accepted one, the unpacking check, np.random put back
    to where the sequential loop stops -- with the batched GPU packing replaced by the reference's own function, try by try.
    Same outputs and same stream position as generate.generate_blocks_with_GT under the same seed."""
    import torch
    from tapenv import generators
    mods = refshim.load(("tools", "generate"))
    tools, generate = mods["tools"], mods["generate"]
    calls = {"batches": 0}

    def fake_calc(blocks, container_size, reward_type):
        blocks = np.asarray(blocks)
        assert blocks.ndim == 3 and reward_type == "C+P+S-lb-hard"
        calls["batches"] += 1
        res = [tools.calc_positions_lb_greedy(b.copy(), container_size, reward_type) for b in blocks]
        return (torch.from_numpy(np.stack([r[0] for r in res])), None, torch.tensor([[bool(v) for v in r[2]] for r in res]), None, None)

    monkeypatch.setattr(generators, "calc_positions_lb_greedy", fake_calc)
    monkeypatch.setattr(generators, "_generate", generate)
    monkeypatch.setattr(generators, "_original_gt", generate.generate_blocks_with_GT)
    n = 10
    for seed in (1, 2, 3):
        for gt_h in (8, 12):
            np.random.seed(seed)
            want = [generate.generate_blocks_with_GT(n, [7, gt_h], [7, 100], 1, [1, 5], input_type, i) for i in range(2)]
            want_next = np.random.random_sample(4)
            np.random.seed(seed)
            got = [generators.generate_blocks_with_GT(n, [7, gt_h], [7, 100], 1, [1, 5], input_type, i) for i in range(2)]
            got_next = np.random.random_sample(4)
            for a, b in zip(got, want):
                assert len(a) == len(b) == 5
                for x, y in zip(a, b):
                    assert np.array_equal(np.asarray(x), np.asarray(y))
            assert np.array_equal(got_next, want_next)
    assert calls["batches"] > 0
