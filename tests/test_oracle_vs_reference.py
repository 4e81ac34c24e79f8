"""CPU, build container only (needs /root/reference): live differential checks of the oracle restatement and of
the host-side dataset loader against the UNMODIFIED Python reference.  Skipped where the reference is absent
(the GPU box); the committed fixtures under tests/golden/ carry the same pins there."""
import os
import shutil
import tempfile

import numpy as np
import pytest

from oracle import oracle, refshim

pytestmark = pytest.mark.skipif(not refshim.available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    return refshim.load(("tools", "generate", "pack"))


CASES = [
    (2, [5, 50], 10, "C+P+S-lb-soft", "LB_GREEDY", "diff", 120),
    (2, [5, 50], 10, "C+P+S-lb-hard", "LB_GREEDY", "zero", 120),
    (2, [4, 60], 12, "C+P-lb-hard", "LB_GREEDY", "full", 80),
    (2, [7, 100], 20, "C+P+S-mcs-hard", "MACS", "diff", 60),
    (2, [5, 50], 10, "C+P+S-mcs-soft", "MACS", "diff", 60),
    (2, [6, 60], 12, "C+P-mcs-soft", "MACS", "full", 40),
    (3, [5, 5, 50], 10, "C+P+S-lb-soft", "LB_GREEDY", "diff", 60),
    (3, [5, 5, 50], 10, "C+P+S-lb-hard", "LB_GREEDY", "zero", 60),
    (3, [4, 6, 80], 14, "C+P+S-lb-soft", "LB_GREEDY", "full", 30),
    (2, [5, 50], 10, "C+P+S-lb-soft", "LB", "diff", 100),
    (2, [6, 50], 12, "C+P+S-lb-hard", "LB", "zero", 100),
    (3, [5, 5, 50], 10, "C+P+S-lb-soft", "LB", "diff", 50),
    (3, [4, 6, 80], 12, "C+P+S-lb-hard", "LB", "full", 30),
    (3, [5, 5, 50], 10, "C+P+S-mcs-soft", "MACS", "diff", 40),      # calc_one_position_mcs_3d, tools.py:2751-3165
    (3, [5, 5, 50], 10, "C+P+S-mcs-hard", "MACS", "zero", 40),
    (3, [4, 6, 80], 14, "C+P-mcs-soft", "MACS", "full", 25),
    (3, [6, 4, 60], 12, "mcs-hard", "MACS", "diff", 25),
    (3, [5, 5, 50], 12, "C+P+S-mul-hard", "MUL", "diff", 25),
]


@pytest.mark.parametrize("dim,size,n,rt,strat,hm,episodes", CASES)
def test_container_differential(ref, dim, size, n, rt, strat, hm, episodes):
    tools = ref["tools"]
    rng = np.random.RandomState(hash((dim, n, rt)) % 2 ** 31)
    for ep in range(episodes):
        c = oracle.Container(size, n, rt, hm, packing_strategy=strat)
        r = tools.Container(size, n, rt, hm, packing_strategy=strat)
        for k in range(n):
            b = rng.randint(1, 5, size=dim).astype(np.float32)
            e = r.add_new_block(b.copy())
            a = c.add_new_block(b.copy())
            assert np.array_equal(np.asarray(a), np.asarray(e)), (ep, k)
            assert np.array_equal(c.heightmap, np.asarray(r.heightmap)) and np.array_equal(c.positions, np.asarray(r.positions))
            assert c.valid_size == r.valid_size and c.empty_size == r.empty_size
            assert list(c.stable) == [bool(x) for x in r.stable]
        assert c.calc_ratio() == r.calc_ratio()
        assert np.array_equal(c.container, np.asarray(r.container))
        if strat in ("MACS", "MUL") and dim == 2:
            assert c.level_free_space == [list(map(int, l)) for l in r.level_free_space]
        if strat in ("MACS", "MUL") and dim == 3:
            assert c.level_free_space == [[list(map(int, l)) for l in row] for row in r.level_free_space]


def test_mask_and_dynamic_differential(ref):
    import torch
    pack = ref["pack"]
    rng = np.random.RandomState(5)
    for dim, n in ((2, 10), (3, 10), (2, 20)):
        R = 2 if dim == 2 else 6
        S = n * R
        B = 32
        static = np.zeros((B, 1 + dim, S), np.float32)
        static[:, 0] = np.tile(np.arange(n), R)
        static[:, 1:] = rng.randint(1, 5, size=(B, dim, S))
        dynamic = (rng.random_sample((B, 3 * n, S)) < 0.1).astype(np.float32)
        mask = np.ones((B, S), np.float32)
        dyn_t, mask_t = torch.from_numpy(dynamic), torch.from_numpy(mask)
        for t in range(n):
            ptr = rng.randint(0, S, size=B).astype(np.int64)
            dynamic = oracle.update_dynamic(dynamic, static, ptr)
            cur, mask = oracle.update_mask(mask, dynamic, static, ptr)
            dyn_t = pack.update_dynamic(dyn_t, torch.from_numpy(static), torch.from_numpy(ptr), "bot", True)
            cur_t, mask_t = pack.update_mask(mask_t, dyn_t, torch.from_numpy(static), torch.from_numpy(ptr), "bot", True)
            assert np.array_equal(dynamic, dyn_t.numpy()) and np.array_equal(cur, cur_t.numpy()) and np.array_equal(mask, mask_t.numpy())


INPUT_LAYOUTS = [("simple", False), ("simple", True), ("rot", True), ("rot-old", True), ("bot", True), ("bot-rot", True),
                 ("use-static", True), ("mul", True), ("mul-with", True)]


def layout_rows(input_type, n, dim):
    """(static rows, dynamic rows) of every input_type (pack.py:186-223)."""
    srows = 2 + dim if input_type in ("mul", "mul-with") else 1 + dim
    drows = n if input_type in ("simple", "rot") else (n + 1 if input_type == "rot-old" else 3 * n)
    return srows, drows


@pytest.mark.parametrize("input_type,allow_rot", INPUT_LAYOUTS)
def test_tensor_ops_every_input_type(ref, input_type, allow_rot):
    """pack.update_dynamic / pack.update_mask for EVERY input_type string (pack.py:285-309, :338-365), including the
    legacy 'rot-old' layout (n movement rows + one rotate-state row, which here carries non-zero entries on purpose)."""
    import torch
    pack = ref["pack"]
    rng = np.random.RandomState(11)
    for dim, n in ((2, 10), (3, 5)):
        R = (2 if dim == 2 else 6) if allow_rot else 1
        S, B = n * R, 16
        srows, drows = layout_rows(input_type, n, dim)
        static = np.zeros((B, srows, S), np.float32)
        static[:, 0] = np.tile(np.arange(n), R)
        static[:, 1:1 + dim] = rng.randint(1, 5, size=(B, dim, S))
        dynamic = (rng.random_sample((B, drows, S)) < 0.08).astype(np.float32)
        mask = np.ones((B, S), np.float32)
        dyn_t, mask_t, st_t = torch.from_numpy(dynamic), torch.from_numpy(mask), torch.from_numpy(static)
        for t in range(n):
            ptr = rng.randint(0, S, size=B).astype(np.int64)
            dynamic = oracle.update_dynamic(dynamic, static, ptr, input_type, allow_rot)
            cur, mask = oracle.update_mask(mask, dynamic, static, ptr, input_type, allow_rot)
            dyn_t = pack.update_dynamic(dyn_t, st_t, torch.from_numpy(ptr), input_type, allow_rot)
            cur_t, mask_t = pack.update_mask(mask_t, dyn_t, st_t, torch.from_numpy(ptr), input_type, allow_rot)
            assert np.array_equal(dynamic, dyn_t.numpy()) and np.array_equal(cur, cur_t.numpy()) and np.array_equal(mask, mask_t.numpy())


@pytest.mark.parametrize("dim", [2, 3])
def test_packdataset_matches_reference_loader(ref, dim):
    """tapenv.PACKDataset builds the same four tensors as pack.PACKDataset from the same dataset directory."""
    from tapenv.dataset import PACKDataset
    pack = ref["pack"]
    cwd = os.getcwd()
    d = tempfile.mkdtemp(prefix="tapds_")
    os.chdir(d)
    try:
        train_dir, _ = pack.create_dataset(10, 24, 4, dim, 7, 50, 1, [1, 5], seed=99)
        for input_type, hm in (("bot", "diff"), ("bot", "full"), ("simple", "zero"), ("bot-rot", "diff"), ("rot-old", "diff"),
                               ("mul-with", "diff")):
            allow_rot = input_type != "simple"
            theirs = pack.PACKDataset(train_dir, 10, 24, 7, input_type, hm, True, 5)
            ours = PACKDataset(train_dir, 10, 24, 7, input_type, hm, True, 5)
            assert np.array_equal(ours.static.numpy(), theirs.static.numpy())
            assert np.array_equal(ours.dynamic.numpy(), theirs.dynamic.numpy())
            assert ours.decoder_static.shape == theirs.decoder_static.shape
            assert ours.decoder_dynamic.shape == theirs.decoder_dynamic.shape
            assert len(ours) == len(theirs) and all(np.array_equal(a.detach().numpy(), b.detach().numpy()) for a, b in zip(ours[3], theirs[3]))
        # mixed datasets (pack.py:67-97): first half from one directory, second half from another
        other_dir, _ = pack.create_dataset(10, 24, 6, dim, 7, 50, 1, [1, 5], seed=5)
        assert other_dir == train_dir                       # same directory name: generate the second set elsewhere
        os.makedirs("second", exist_ok=True)
        os.chdir("second")
        mix_dir, _ = pack.create_dataset(10, 24, 4, dim, 7, 50, 1, [1, 5], seed=123)
        mix_dir = os.path.join(os.getcwd(), mix_dir) + ("" if mix_dir.endswith("/") else "/")
        os.chdir(d)
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            theirs = pack.PACKDataset(train_dir, 10, 24, 7, "bot", "diff", True, 5, mix_data_file=mix_dir)
        ours = PACKDataset(train_dir, 10, 24, 7, "bot", "diff", True, 5, mix_data_file=mix_dir)
        assert np.array_equal(ours.static.numpy(), theirs.static.numpy())
        assert np.array_equal(ours.dynamic.numpy(), theirs.dynamic.numpy())
        assert not np.array_equal(ours.static.numpy()[:12], ours.static.numpy()[12:])
    finally:
        os.chdir(cwd)
        shutil.rmtree(d, ignore_errors=True)
