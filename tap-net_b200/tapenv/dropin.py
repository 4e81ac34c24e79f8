"""Point an UNMODIFIED TAP-Net checkout at the B200 environment.

    import pack, tools            # the reference's modules (trainer.py imports `pack` inside train_pack, model.py
    import tapenv                 # does `import tools` at the top)
    tapenv.install(pack, tools)   # before trainer.train_pack(**kwargs) runs
    ...
    tapenv.uninstall()

Replaces exactly the hot-path symbols: pack.update_dynamic, pack.update_mask, pack.reward,
tools.Container, tools.calc_positions_lb_greedy (and generate.InitialContainer when `generate` is passed).  Everything else of the reference keeps running as it is."""
import sys

from . import containers, episode, ops, rolling

_saved = []


def install(pack=None, tools=None, generate=None):
    """generate: the reference's `generate` module -- also replaces generate.InitialContainer (the rolling window,
    rolling.py:501) by the GPU-backed per-instance class."""
    pack = pack if pack is not None else sys.modules.get("pack")
    tools = tools if tools is not None else sys.modules.get("tools")
    if pack is None and tools is None:
        raise RuntimeError("tapenv.install: import the reference's `pack` / `tools` modules first (or pass them)")
    repl = []
    if pack is not None:
        repl += [(pack, "update_dynamic", ops.update_dynamic), (pack, "update_mask", ops.update_mask),
                 (pack, "reward", episode.reward)]
    if tools is not None:
        repl += [(tools, "Container", containers.Container),
                 (tools, "calc_positions_lb_greedy", episode.calc_positions_lb_greedy)]
    if generate is not None:
        repl += [(generate, "InitialContainer", rolling.InitialContainer)]
    for mod, name, new in repl:
        _saved.append((mod, name, getattr(mod, name, None)))
        setattr(mod, name, new)
    return [name for _, name, _ in repl]


def uninstall():
    while _saved:
        mod, name, old = _saved.pop()
        if old is None:
            delattr(mod, name)
        else:
            setattr(mod, name, old)
