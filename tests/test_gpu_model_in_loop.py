"""GPU: the UNMODIFIED reference network in the loop.

`model.DRL.forward` (model.py:254-515) -- the real class from the reference tree, the shipped pretrained actor, constructed
as trainer.py:461-480 constructs it -- runs on the B200 twice: with the reference's own environment (tools.Container on the
host, pack.update_* as torch ops) and after `tapenv.install(pack, tools)` (same model.py, same weights, environment = the
CUDA kernels behind the C ABI).  Tours, log-probabilities and rewards must be identical.  Then the same network modules
drive `tapenv.DecodeLoop` through `tapenv.adapters.drl_actor_step` (no host round trips, optionally one CUDA graph).
The reference modules come from oracle/_ref (staged by oracle/stage_ref.py; /root/reference does not exist on the GPU box)."""
import numpy as np
import pytest
import torch

from tests import ref_model
from tests.golden_io import golden_path, load_inputs

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_model.available(), reason="reference tree not staged (oracle/stage_ref.py)")]

SIZES = {2: [5, 50], 3: [5, 5, 50]}


def _inputs(dim, B):
    static, dynamic = load_inputs("rand%dd_n10.npz" % dim, B)
    return torch.from_numpy(static).cuda(), torch.from_numpy(dynamic).cuda()


@pytest.fixture
def installed():
    import tapenv
    mods = ref_model.reference_modules()
    names = tapenv.install(mods["pack"], mods["tools"])
    yield mods
    tapenv.uninstall()
    assert names


@pytest.mark.parametrize("dim", [2, 3])
def test_unmodified_drl_forward_greedy_install_equals_reference_env(dim):
    """eval() = greedy decode (model.py:370-372): reference env vs tapenv.install(), B=64, and the G6 recording."""
    import tapenv
    B = 64 if dim == 2 else 32
    st, dy = _inputs(dim, B)
    mods = ref_model.reference_modules()
    actor = ref_model.make_actor(dim, True).eval()
    assert actor.update_fn is mods["pack"].update_dynamic and actor.update_fn.__module__ == "pack"
    with torch.no_grad():
        tour_ref, logp_ref, _, r_ref = ref_model.forward(actor, st, dy)

    tapenv.install(mods["pack"], mods["tools"])
    try:
        actor2 = ref_model.make_actor(dim, True).eval()               # picks up pack.update_* = tapenv's, like trainer.py:461-480
        assert actor2.update_fn is tapenv.update_dynamic and mods["tools"].Container is tapenv.Container
        with torch.no_grad():
            tour, logp, _, r = ref_model.forward(actor2, st, dy)
    finally:
        tapenv.uninstall()
    assert mods["tools"].Container is not tapenv.Container
    assert torch.equal(tour, tour_ref)
    assert torch.equal(logp, logp_ref)
    assert float((r - r_ref).abs().max()) <= 1e-6 and torch.equal(r, r_ref)

    # G6: the same forward recorded from the fp32 CPU network in the build container (tests/golden/make_golden.py:make_g6).
    # A GPU BLAS / TF32 convolution may flip an argmax between near-equal candidates (r02a on a B200: 2D all but a few
    # tours agree, 3D -- 60 candidates, flatter distributions -- 59 %), so tours are compared per environment: wherever the
    # tour agrees the reward must agree (<= 1e-6), and a good share of the tours must agree.
    g6 = np.load(golden_path("g6_tours.npz"))
    gt, gr = g6["g6_%dd_tour" % dim], g6["g6_%dd_reward" % dim]
    num = min(B, gt.shape[0])
    same = (tour[:num].cpu().numpy() == gt[:num]).all(1)
    assert same.mean() >= (0.8 if dim == 2 else 0.3)
    assert np.abs(r[:num].cpu().numpy()[same] - gr[:num][same]).max() <= 1e-6


@pytest.mark.parametrize("dim", [2, 3])
def test_unmodified_drl_forward_training_mode_and_backward(dim, installed):
    """train() mode: Categorical sampling + pointer dropout under a fixed seed give the same tours with either environment
    (the env consumes no RNG); the REINFORCE loss of trainer.py:216-228 back-propagates through the tapenv tensors."""
    import tapenv
    B = 48
    st, dy = _inputs(dim, B)
    actor = ref_model.make_actor(dim, True).train()
    torch.manual_seed(7)
    tour, logp, _, r = ref_model.forward(actor, st, dy)
    loss = torch.mean(r.detach() * logp.sum(dim=1))                   # advantage * tour_logp (critic left out)
    loss.backward()
    g = actor.dynamic_encoder.conv.weight.grad
    assert g is not None and torch.isfinite(g).all() and float(g.abs().sum()) > 0
    tapenv.uninstall()
    mods = ref_model.reference_modules()
    actor_ref = ref_model.make_actor(dim, True).train()
    assert actor_ref.update_fn is mods["pack"].update_dynamic
    torch.manual_seed(7)
    tour_ref, logp_ref, _, r_ref = ref_model.forward(actor_ref, st, dy)
    assert torch.equal(tour, tour_ref) and torch.equal(logp, logp_ref) and torch.equal(r, r_ref)
    # every tour visits each block exactly once
    assert torch.equal(torch.sort(tour % 10, 1).values, torch.arange(10, device=tour.device).expand(B, -1))


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("use_graph", [False, True])
def test_decode_loop_with_reference_network_modules(dim, use_graph):
    """tapenv.adapters.drl_actor_step(actor) feeding DecodeLoop == model.DRL.forward (tours, log-probs, rewards)."""
    import tapenv
    B = 64
    st, dy = _inputs(dim, B)
    actor = ref_model.make_actor(dim, True).eval()
    with torch.no_grad():
        tour_ref, logp_ref, _, r_ref = ref_model.forward(actor, st, dy)
        env = tapenv.BatchedContainers(SIZES[dim], 10, "C+P+S-lb-soft", "diff", batch_size=B)
        loop = tapenv.DecodeLoop(env, tapenv.adapters.drl_actor_step(actor), greedy=True, use_graph=use_graph)
        for rep in range(2):
            tour, logp, reward = loop.run(st, dy)
            assert torch.equal(tour, tour_ref)
            assert float((logp - logp_ref).abs().max()) <= 1e-5
            assert torch.equal(-reward, r_ref)
    env.check_flags()


def test_incremental_dynamic_hidden_matches_encoder():
    """model.py:380 as a rank-3 correction (tapenv.adapters.incremental_dynamic_hidden) against the fp32 PyTorch reference
    (the encoder re-run on the updated tensor): |delta| <= 1e-5 absolute on activations of magnitude ~10."""
    import tapenv
    from tapenv.adapters import incremental_dynamic_hidden
    B, n = 512, 10
    st, dy = _inputs(2, B)
    actor = ref_model.make_actor(2, True).eval()
    g = torch.Generator(device="cuda").manual_seed(3)
    env = tapenv.BatchedContainers(SIZES[2], n, "C+P+S-lb-soft", "diff", batch_size=B)
    with torch.no_grad():
        cur, mask = env.reset(dy)
        hid = actor.dynamic_encoder(dy)
        dyn = dy
        for t in range(n):
            ptr = torch.multinomial(cur, 1, generator=g).squeeze(1)
            new, cur, mask, _, _ = env.step(ptr, st, dyn, mask)
            real = torch.gather(st[:, 0, :], 1, ptr.view(-1, 1)).long()
            rows = real + n * torch.arange(3, device="cuda").view(1, -1)
            hid = incremental_dynamic_hidden(actor.dynamic_encoder, hid, dyn, new, rows)
            want = actor.dynamic_encoder(new)
            assert float((hid - want).abs().max()) <= 1e-5 * (t + 1)
            dyn = new


def test_generate_blocks_runs_after_install(installed):
    """generate.generate_blocks calls tools.calc_positions_lb_greedy(...)[1] as a voxel grid
    (generate.py:908 -> calc_dependent); with tapenv installed the dataset generator must still run and produce the same
    sample as the reference under the same NumPy seed (2D through the kernels, 3D 7x7 falls back to the reference)."""
    import tapenv
    mods = installed
    gen = mods["generate"]
    assert mods["tools"].calc_positions_lb_greedy.tapenv_original is not None
    for size in ([7, 50], [7, 7, 50]):
        np.random.seed(99)
        got = gen.generate_blocks(10, size, 1, [1, 5])
        tapenv.uninstall()
        np.random.seed(99)
        want = gen.generate_blocks(10, size, 1, [1, 5])
        tapenv.install(mods["pack"], mods["tools"])
        for a, b in zip(got, want):
            assert np.array_equal(np.asarray(a), np.asarray(b))


@pytest.mark.parametrize("dim,W", [(2, 7), (2, 0), (3, 5), (3, 7)])
def test_create_dataset_identical_with_batched_generator(dim, W, tmp_path, monkeypatch):
    """pack.create_dataset (pack.py:580-667) with generate.generate_blocks replaced by the speculative GPU rejection loop
    (tapenv.generators): byte-identical dataset files under the same seed -- same block sets accepted, same NumPy stream
    position afterwards (the shuffle / random_integers calls between the samples see the same state).  W=7 in 3D (49 cells)
    and W=0 in 3D (random width up to 10) exceed the compiled limits and run the saved reference function."""
    import contextlib
    import io
    import os
    import tapenv
    mods = ref_model.reference_modules()
    pack, tools, generate = mods["pack"], mods["tools"], mods["generate"]

    def build(sub):
        d = tmp_path / sub
        d.mkdir()
        monkeypatch.chdir(d)
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            train_dir, valid_dir = pack.create_dataset(10, 10, 4, dim, W, 50, 1, [1, 5], seed=77)
        out = {}
        for sdir in (train_dir, valid_dir):
            for f in sorted(os.listdir(sdir)):
                out[os.path.join(sdir, f)] = open(os.path.join(sdir, f)).read()
        return out

    want = build("reference")
    tapenv.install(pack, tools, generate)
    try:
        assert generate.generate_blocks is tapenv.generators.generate_blocks
        got = build("tapenv")
    finally:
        tapenv.uninstall()
    assert generate.generate_blocks is not tapenv.generators.generate_blocks
    assert sorted(got) == sorted(want) and len(want) == 12
    for k in want:
        assert got[k] == want[k], k


@pytest.mark.parametrize("n,gt_size,ics", [(10, [7, 8], [7, 100]), (14, [7, 12], [7, 100]), (8, [5, 5, 4], [7, 7, 100])])
def test_ppsg_generator_identical_with_batched_tries(n, gt_size, ics):
    """generate.generate_blocks_with_GT (the PPSG generator behind pack.create_dataset_gt / BASELINE C4, generate.py:17-229)
    replaced by tapenv.generators.generate_blocks_with_GT: the 20 unpacking orders of every perfect packing are packed in ONE
    GPU batch instead of 20 host calls of calc_positions_lb_greedy (:112).  Same samples and the same NumPy stream position as
    the reference under the same seed; the 3D 7x7 initial container (49 cells) runs the saved reference function."""
    import time
    import tapenv
    mods = ref_model.reference_modules()
    pack, tools, generate = mods["pack"], mods["tools"], mods["generate"]
    np.random.seed(2024)
    t0 = time.perf_counter()
    want = [generate.generate_blocks_with_GT(n, list(gt_size), list(ics), 1, [1, 5], "bot", i) for i in range(2)]
    t_ref = time.perf_counter() - t0
    want_next = np.random.random_sample(3)
    tapenv.install(pack, tools, generate)
    try:
        assert generate.generate_blocks_with_GT is tapenv.generators.generate_blocks_with_GT
        np.random.seed(2024)
        t0 = time.perf_counter()
        got = [generate.generate_blocks_with_GT(n, list(gt_size), list(ics), 1, [1, 5], "bot", i) for i in range(2)]
        t_ours = time.perf_counter() - t0
        got_next = np.random.random_sample(3)
    finally:
        tapenv.uninstall()
    assert generate.generate_blocks_with_GT is not tapenv.generators.generate_blocks_with_GT
    for a, b in zip(got, want):
        for x, y in zip(a, b):
            assert np.array_equal(np.asarray(x), np.asarray(y))
    assert np.array_equal(got_next, want_next)
    print("PPSG n=%d: reference %.2f s, tapenv %.2f s" % (n, t_ref, t_ours))
