"""GPU (B200): the UNMODIFIED reference trainer -- trainer.train_pack (trainer.py:317-504: dataset creation, PACKDataset,
DRL + critic construction with pack.update_dynamic / pack.update_mask injected at :476-477) and trainer.train (:140-313:
REINFORCE with the critic baseline, Adam, gradient clipping, checkpoints, validation) -- run end to end twice on a tiny RAND
configuration under the same seeds: with the reference's own environment, and after tapenv.install(pack, tools, generate).
Nothing of trainer.py / model.py / pack.py is edited; only plotting is off (tests/ref_trainer.py).  The environment consumes
no random numbers and returns the same masks / heightmaps / rewards, so the two runs must log the same rewards and losses
and save the same weights."""
import numpy as np
import pytest

from tests import ref_model, ref_trainer

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def deterministic_torch():
    """Both arms of a comparison must compute the same network numbers: deterministic cuDNN / scatter implementations for the
    duration of a test (a harness setting -- nothing of the reference is touched)."""
    import torch
    prev = (torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark, torch.are_deterministic_algorithms_enabled(),
            torch.is_deterministic_algorithms_warn_only_enabled())
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    torch.use_deterministic_algorithms(True, warn_only=True)
    yield
    torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = prev[0], prev[1]
    torch.use_deterministic_algorithms(prev[2], warn_only=prev[3])


@pytest.mark.parametrize("dim,strategy,reward_type,input_type", [(2, "LB_GREEDY", "C+P+S-lb-soft", "bot"), (3, "LB_GREEDY", "C+P+S-lb-hard", "bot"),
                                                                 (2, "MACS", "C+P+S-mcs-soft", "bot"), (2, "LB_GREEDY", "C+P+S-lb-soft", "mul-with"), (3, "LB_GREEDY", "C+P+S-lb-soft", "mul"),
                                                                 (2, "LB_GREEDY", "C+P+S-lb-soft", "simple")])
def test_unmodified_train_pack_same_run_with_tapenv(dim, strategy, reward_type, input_type, tmp_path):
    import torch
    import tapenv
    if not ref_model.available():
        pytest.skip("reference tree not staged")
    mods = ref_trainer.modules()
    pack, tools, generate = mods["pack"], mods["tools"], mods["generate"]
    # ('mul' / 'mul-with': two container lists per batch, model.py:291-292 -- the per-object proxies serve each step with one launch;
    #  'simple': one precedence band, allow_rot off)
    kw = dict(obj_dim=dim, packing_strategy=strategy, reward_type=reward_type, input_type=input_type, allow_rot=input_type != "simple",
              train_size=64, valid_size=16, batch_size=32, epoch_num=2)
    want = ref_trainer.run_train_pack(str(tmp_path / "reference"), **kw)
    tapenv.install(pack, tools, generate)
    try:
        assert pack.update_dynamic is tapenv.update_dynamic and tools.Container is tapenv.Container
        got = ref_trainer.run_train_pack(str(tmp_path / "tapenv"), **kw)
    finally:
        tapenv.uninstall()
    assert got["files"] == want["files"] and "checkpoints/1/actor.pt" in got["files"]
    assert np.array_equal(got["rewards"], want["rewards"]), (got["rewards"], want["rewards"])
    np.testing.assert_allclose(got["losses"], want["losses"], rtol=1e-5, atol=1e-6)
    for k in want["actor"]:
        assert torch.allclose(got["actor"][k], want["actor"][k], rtol=1e-4, atol=1e-5), k


@pytest.mark.parametrize("strategy,reward_type,total", [("LB_GREEDY", "C+P+S-lb-soft", 20), ("LB_GREEDY", "C+P+S-lb-hard", 30),
                                                        ("MACS", "C+P+S-mcs-soft", 20)])
def test_unmodified_rolling_driver_same_statistics_with_tapenv(strategy, reward_type, total, tmp_path):
    """The UNMODIFIED rolling-inference driver (rolling.train_pack -> rolling.validate, rolling.py:575-760): a RollingDataset
    of generate.InitialContainer objects, rolling.DRL decoding one block per window refill, ONE tools.Container taking all
    `total` blocks -- run twice under the same seeds: reference environment vs tapenv.install(pack, tools, generate), which
    swaps in the GPU-backed InitialContainer, Container, update_dynamic / update_mask and the batched dataset generator.  The
    per-instance statistics files the driver writes (valid / box / empty size, stable count, packing height) must agree.
    (2D only: rolling.py's own DRL constructor raises for obj_dim = 3 -- HeightmapEncoder receives a tuple, rolling.py:254/:104.)"""
    import torch
    import tapenv
    from tests import ref_rolling
    if not ref_model.available():
        pytest.skip("reference tree not staged")
    mods = ref_rolling.modules()
    pack, tools, generate = mods["pack"], mods["tools"], mods["generate"]
    kw = dict(packing_strategy=strategy, reward_type=reward_type, total_blocks_num=total, valid_size=4)
    want = ref_rolling.run_rolling(str(tmp_path / "reference"), **kw)
    tapenv.install(pack, tools, generate)
    try:
        assert generate.InitialContainer is tapenv.rolling.InitialContainer
        got = ref_rolling.run_rolling(str(tmp_path / "tapenv"), **kw)
    finally:
        tapenv.uninstall()
    for name in ref_rolling.STATS:
        assert np.array_equal(got[name], want[name]), (name, got[name], want[name])
    assert (want["batch-valid_size.txt"] > 0).all()
