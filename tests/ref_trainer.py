"""Checker-side helper: run the UNMODIFIED reference trainer (trainer.train_pack, trainer.py:317-495 -> train :140-313) on a
tiny RAND configuration in a scratch directory and collect what it writes.  Plotting is the only thing switched off
(matplotlib is absent here: pyplot is a no-op stub, pack.render -- pure visualisation -- is replaced by a no-op for the call).
Used by tests/test_gpu_trainer.py; never by the product."""
import contextlib
import glob
import io
import os

import numpy as np

KWARGS = dict(task="train", note="t", use_cuda=True, cuda="0", cpu_threads=0, checkpoint=None, seed=12345,
              train_size=64, valid_size=16, epoch_num=2, batch_size=32, obj_dim=2, num_nodes=10, total_obj_num=10, dataset="RAND",
              unit=1.0, arm_size=1, min_size=1, max_size=5, container_width=5, container_length=5, container_height=50,
              initial_container_width=7, initial_container_length=7, initial_container_height=50,
              packing_strategy="LB_GREEDY", reward_type="C+P+S-lb-soft", input_type="bot", allow_rot=True,
              decoder_input_type="shape_heightmap", heightmap_type="diff", no_precedence=False, dropout=0.1, actor_lr=5e-4,
              critic_lr=5e-4, max_grad_norm=2., n_process_blocks=3, num_layers=1, encoder_hidden_size=128, decoder_hidden_size=256)


def modules():
    from oracle import refshim
    return refshim.load(("tools", "generate", "pack", "model", "trainer"))


def run_train_pack(workdir, **overrides):
    """trainer.train_pack(**kwargs) with cwd = workdir -> dict(rewards, losses, actor state_dict, files)."""
    import torch
    mods = modules()
    pack, trainer = mods["pack"], mods["trainer"]
    kw = dict(KWARGS)
    kw.update(overrides)
    os.makedirs(workdir, exist_ok=True)
    cwd = os.getcwd()
    render = pack.render
    pack.render = lambda *a, **k: None
    os.chdir(workdir)
    try:
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            trainer.train_pack(**kw)
        runs = glob.glob(os.path.join("pack", str(kw["num_nodes"]), "*"))
        assert len(runs) == 1, runs
        out = {"rewards": np.loadtxt(os.path.join(runs[0], "reawrds.txt")), "losses": np.loadtxt(os.path.join(runs[0], "losses.txt")),
               "actor": torch.load(os.path.join(runs[0], "actor.pt"), map_location="cpu"),
               "files": sorted(os.path.relpath(os.path.join(d, f), runs[0]) for d, _, fs in os.walk(runs[0]) for f in fs)}
    finally:
        os.chdir(cwd)
        pack.render = render
    return out
