"""Checker-side helpers: build the UNMODIFIED reference network (model.DRL, constructed as trainer.py:461-480 does) with the
shipped pretrained actor, from oracle/refshim (the /root/reference tree in the build container, the staged oracle/_ref copy on
the GPU box).  Used by the model-in-the-loop tests and by bench.py's `model_in_loop` block; never by the product."""
import contextlib
import io

CHECKPOINTS = {2: "2d-bot-C+P+S-lb-soft-width-5-note-sh-R-diff", 3: "3d-bot-C+P+S-lb-soft-width-5-note-sh-R-diff"}


def available():
    from oracle import refshim
    import os
    return refshim.available() and os.path.isfile(refshim.checkpoint(CHECKPOINTS[2]))


def reference_modules():
    from oracle import refshim
    return refshim.load(("tools", "generate", "pack", "model"))


def make_actor(dim, use_cuda, update_fn=None, mask_fn=None, reward_type="C+P+S-lb-soft", packing_strategy="LB_GREEDY",
               container_width=5, container_height=50, dropout=0.1, pretrained=True):
    """model.DRL(STATIC_SIZE=dim, DYNAMIC_SIZE=3n, 128, 256, use_cuda, 'bot', True, W, H, dim, reward_type, 'shape_heightmap',
    'diff', packing_strategy, pack.update_dynamic, pack.update_mask, 1, dropout, 1.0)  -- scripts/train.sh defaults.
    update_fn / mask_fn default to WHATEVER pack.update_* currently are (i.e. tapenv's after tapenv.install())."""
    import torch
    from oracle import refshim
    mods = reference_modules()
    pack, model = mods["pack"], mods["model"]
    with contextlib.redirect_stdout(io.StringIO()):
        actor = model.DRL(dim, 30, 128, 256, use_cuda, "bot", True, container_width, container_height, dim, reward_type,
                          "shape_heightmap", "diff", packing_strategy,
                          update_fn if update_fn is not None else pack.update_dynamic,
                          mask_fn if mask_fn is not None else pack.update_mask, 1, dropout, 1.0)
    if pretrained:
        actor.load_state_dict(torch.load(refshim.checkpoint(CHECKPOINTS[dim]), map_location="cpu"))
    if use_cuda:
        actor = actor.cuda()
    return actor


def decoder_inputs(B, dim, device, W=5):
    """PACKDataset's zero decoder inputs (pack.py:228-266) for heightmap_type 'diff'."""
    import torch
    dec_static = torch.zeros(B, dim, 1, device=device)
    dec_dyn = torch.zeros(B, W - 1, 1, device=device) if dim == 2 else torch.zeros(B, 2, W, W, device=device)
    return [dec_static, dec_dyn]


def forward(actor, static, dynamic):
    """actor(static, dynamic, [decoder_static, decoder_dynamic]) as trainer.py:200 calls it (prints silenced)."""
    dim = static.shape[1] - 1
    with contextlib.redirect_stdout(io.StringIO()):
        return actor(static, dynamic, decoder_inputs(static.shape[0], dim, static.device))
