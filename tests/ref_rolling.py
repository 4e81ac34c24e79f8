"""Checker-side helper: run the UNMODIFIED rolling-inference driver of the reference (rolling.train_pack -> rolling.validate,
rolling.py:575-760: RollingDataset of generate.InitialContainer objects, the one-step decode loop around rolling.DRL, one
tools.Container per run) on a tiny configuration in a scratch directory and collect the statistics files it writes.  Only
drawing is switched off (matplotlib is absent: tools.draw_container_2d / draw_container_voxel become no-ops for the call).
Used by tests/test_gpu_trainer.py; never by the product."""
import argparse
import contextlib
import glob
import io
import os

import numpy as np

ARGS = dict(task="pack", note="t", just_test=True, just_generate=False, use_cuda=True, cuda="0", cpu_threads=0, checkpoint=None,
            seed=12345, train_size=2, valid_size=4, epoch_num=1, batch_size=128, obj_dim=2, gt_data=False, mix_data=False,
            num_nodes=10, unit=1, arm_size=1, min_size=1, max_size=5, container_width=5, container_height=250,
            initial_container_width=7, initial_container_height=250, packing_strategy="LB_GREEDY", reward_type="C+P+S-lb-soft",
            input_type="bot", allow_rot=True, decoder_input_type="shape_heightmap", heightmap_type="diff", dropout=0.1,
            actor_lr=5e-4, critic_lr=5e-4, max_grad_norm=2., n_process_blocks=3, num_layers=1, encoder_hidden_size=128,
            decoder_hidden_size=256, total_blocks_num=20)
STATS = ("batch-valid_size.txt", "batch-box_size.txt", "batch-empty_size.txt", "batch-stable_num.txt", "batch-packing_height.txt")


def modules():
    from oracle import refshim
    return refshim.load(("tools", "generate", "pack", "model", "rolling"))


def run_rolling(workdir, **overrides):
    """rolling.train_pack(args) with cwd = workdir -> {file name: array} of the per-instance statistics."""
    mods = modules()
    tools, rolling = mods["tools"], mods["rolling"]
    kw = dict(ARGS)
    kw.update(overrides)
    for sub in ("rand_2d", "rand_3d"):               # the checkout ships ./data/<kind>/; rolling.get_dataset uses os.mkdir
        os.makedirs(os.path.join(workdir, "data", sub), exist_ok=True)
    cwd = os.getcwd()
    saved = (tools.draw_container_2d, tools.draw_container_voxel)
    tools.draw_container_2d = tools.draw_container_voxel = lambda *a, **k: None
    os.chdir(workdir)
    try:
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            rolling.train_pack(argparse.Namespace(**kw))
        runs = glob.glob(os.path.join(kw["task"], str(kw["num_nodes"]), "*"))
        assert len(runs) == 1, runs
        return {f: np.atleast_1d(np.loadtxt(os.path.join(runs[0], "render", "0", f))) for f in STATS}
    finally:
        os.chdir(cwd)
        tools.draw_container_2d, tools.draw_container_voxel = saved
