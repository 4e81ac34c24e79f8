"""GPU: tapenv.DecodeLoop -- model.DRL.forward's decode loop (model.py:342-515) without host round trips -- against a
step-by-step CPU replica of the same loop on the oracle environment, with an actor whose logits are exact in fp32."""
import numpy as np
import pytest
import torch

from oracle import oracle
from tests.golden_io import load_inputs

pytestmark = pytest.mark.gpu


def make_actor(S, scale=1.0):
    """Integer-valued (exact in fp32 on CPU and GPU) logits with a unique maximum per row: torch.max's tie-break differs
    between devices."""
    def actor(static, dynamic, dec_static, dec_dyn, state):
        t = 0 if state is None else state
        B = static.shape[0]
        col = torch.arange(S, device=static.device, dtype=torch.float32)[None, :]
        hm = dec_dyn.reshape(B, -1).abs().sum(1, keepdim=True)
        blocked = dynamic.sum(1)                                   # uses the CURRENT precedence tensor
        score = (static[:, 1:].sum(1) * 7 + col * 3 + hm * 5 + blocked * 2 + dec_static.sum(1, keepdim=True) + t) % 11
        return (score * 64 + col) * scale, t + 1
    return actor


def cpu_replica(static, dynamic, size, rt, hm_type, strat, actor, greedy=True):
    B, rows, S = static.shape
    dim = rows - 1
    R = 2 if dim == 2 else 6
    n = S // R
    conts = [oracle.Container(size, n, rt, hm_type, packing_strategy=strat) for _ in range(B)]
    mask = np.ones((B, S), np.float32)
    cur = oracle.initial_mask(dynamic, n, R)
    dyn = dynamic
    dec_static = np.zeros((B, dim), np.float32)
    enc_len = len(np.asarray(conts[0].get_heightmap()).reshape(-1))
    dec_dyn = np.zeros((B, enc_len), np.float32)
    state = None
    idx, logps = [], []
    st_t = torch.from_numpy(static)
    for _ in range(n):
        logits, state = actor(st_t, torch.from_numpy(dyn), torch.from_numpy(dec_static), torch.from_numpy(dec_dyn), state)
        probs = torch.softmax(logits + torch.from_numpy(cur).log(), dim=1)
        prob, ptr = torch.max(probs, 1)
        ptr_n = ptr.numpy().astype(np.int64)
        dyn = oracle.update_dynamic(dyn, static, ptr_n)
        cur, mask = oracle.update_mask(mask, dyn, static, ptr_n)
        dec_static = static[np.arange(B)[:, None], 1 + np.arange(dim)[None, :], ptr_n[:, None]]
        dec_dyn = np.stack([np.asarray(conts[b].add_new_block(dec_static[b])).reshape(-1) for b in range(B)]).astype(np.float32)
        idx.append(ptr_n); logps.append(prob.log().numpy())
    return np.stack(idx, 1), np.stack(logps, 1), np.array([c.calc_ratio() for c in conts])


@pytest.mark.parametrize("fixture,size,strat,rt", [("rand2d_n10.npz", [5, 50], "LB_GREEDY", "C+P+S-lb-soft"),
                                                   ("rand3d_n10.npz", [5, 5, 50], "LB_GREEDY", "C+P+S-lb-hard"),
                                                   ("ppsg2d_n20.npz", [7, 100], "MACS", "C+P+S-mcs-hard")])
@pytest.mark.parametrize("use_graph", [False, True])
def test_greedy_decode_loop_matches_cpu_replica(fixture, size, strat, rt, use_graph):
    import tapenv
    B = 96
    static, dynamic = load_inputs(fixture, B)
    S = static.shape[2]
    n = S // (2 if len(size) == 2 else 6)
    actor = make_actor(S)
    want_idx, want_logp, want_r = cpu_replica(static, dynamic, size, rt, "diff", strat, actor)
    env = tapenv.BatchedContainers(size, n, rt, "diff", packing_strategy=strat, batch_size=B)
    loop = tapenv.DecodeLoop(env, actor, greedy=True, use_graph=use_graph)
    st, dy = torch.from_numpy(static).cuda(), torch.from_numpy(dynamic).cuda()
    for rep in range(2):                                           # the second call replays the captured graph
        tour_idx, tour_logp, reward = loop.run(st, dy)
        assert np.array_equal(tour_idx.cpu().numpy(), want_idx)
        assert np.allclose(tour_logp.cpu().numpy(), want_logp, atol=1e-6)
        assert np.abs(reward.cpu().numpy().astype(np.float64) - want_r).max() <= 1e-6
    env.check_flags()


def test_sampling_decode_loop_is_valid_and_reproducible():
    import tapenv
    B = 256
    static, dynamic = load_inputs("rand2d_n10.npz", B)
    S = static.shape[2]
    env = tapenv.BatchedContainers([5, 50], 10, "C+P+S-lb-soft", "diff", batch_size=B)
    st, dy = torch.from_numpy(static).cuda(), torch.from_numpy(dynamic).cuda()
    outs = []
    for _ in range(2):
        g = torch.Generator(device="cuda").manual_seed(5)
        outs.append(tapenv.DecodeLoop(env, make_actor(S, scale=1.0 / 64), greedy=False, generator=g).run(st, dy))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][2], outs[1][2])
    tour = outs[0][0].cpu().numpy()
    # every block exactly once per environment (model.py:342-345), never an inaccessible candidate
    assert np.array_equal(np.sort(tour % 10, axis=1), np.tile(np.arange(10), (B, 1)))
    assert torch.isfinite(outs[0][1]).all() and (outs[0][1] <= 0).all()
    # replaying the sampled tour through the oracle reproduces the reward
    ref = oracle.episode_batch(static, dynamic, tour.T.copy(), [5, 50], "C+P+S-lb-soft", "diff", "LB_GREEDY", want=("reward",))
    assert np.array_equal(ref["reward"], outs[0][2].cpu().numpy())


def test_rolling_decode_loop_matches_cpu_replica():
    """tapenv.RollingDecodeLoop (rolling.validate's loop, greedy like actor.eval()) against a CPU replica on the oracle's
    InitialContainer + Container with the same integer-logit actor: identical pointers, windows and rewards."""
    import tapenv
    from tests.golden_io import load_rolling
    B, T, n, dim = 24, 50, 10, 3
    z = load_rolling("rolling3d_t50.npz", B)
    size, rt = [5, 5, 250], "C+P+S-lb-soft"
    S = n * 6
    actor = make_actor(S)
    # ---- CPU replica ----
    ics = [oracle.InitialContainer(z["adj"][b], z["blocks"][b], T, n, dim) for b in range(B)]
    conts = [oracle.Container(size, T, rt, "diff") for _ in range(B)]
    dec_static = np.zeros((B, dim), np.float32)
    dec_dyn = np.zeros((B, 2 * 25), np.float32)
    state, want_idx = None, []
    t = 0
    while t < T:
        outs = [ic.convert_to_input() for ic in ics]
        static = np.stack([o[0] for o in outs]); dyn = np.stack([o[1] for o in outs])
        last = ics[0].is_last_graph()
        mask = np.ones((B, S), np.float32)
        cur = oracle.initial_mask(dyn, n, 6)
        for _ in range(n if last else 1):
            logits, state = actor(torch.from_numpy(static), torch.from_numpy(dyn), torch.from_numpy(dec_static),
                                  torch.from_numpy(dec_dyn), state)
            probs = torch.softmax(logits + torch.from_numpy(cur).log(), dim=1)
            ptr = torch.max(probs, 1)[1].numpy().astype(np.int64)
            want_idx.append(ptr)
            dec_static = static[np.arange(B)[:, None], 1 + np.arange(dim)[None, :], ptr[:, None]]
            dec_dyn = np.stack([np.asarray(conts[b].add_new_block(dec_static[b])).reshape(-1) for b in range(B)]).astype(np.float32)
            if last:
                dyn = oracle.update_dynamic(dyn, static, ptr)
                cur, mask = oracle.update_mask(mask, dyn, static, ptr)
            t += 1
        for b, ic in enumerate(ics):
            ic.remove_block(ic.sub_graph_nodes[int(ptr[b]) % n])
    want_idx = np.stack(want_idx, 1)
    want_r = np.array([c.calc_ratio() for c in conts])
    # ---- device loop ----
    env = tapenv.BatchedContainers(size, T, rt, "diff", batch_size=B, window=n)
    win = tapenv.BatchedInitialContainers(z["adj"], z["blocks"], T, n, dim)
    loop = tapenv.RollingDecodeLoop(tapenv.RollingRunner(env, win), actor)
    tour_idx, tour_logp, reward = loop.run()
    assert np.array_equal(tour_idx.cpu().numpy(), want_idx)
    assert np.abs(reward.cpu().numpy().astype(np.float64) - want_r).max() <= 1e-6
    assert torch.isfinite(tour_logp).all()
    env.check_flags(); win.check_flags()
