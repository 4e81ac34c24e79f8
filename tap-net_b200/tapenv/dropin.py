"""Point an UNMODIFIED TAP-Net checkout at the B200 environment.

    import pack, tools            # the reference's modules (trainer.py imports `pack` inside train_pack, model.py
    import tapenv                 # does `import tools` at the top)
    tapenv.install(pack, tools)   # before trainer.train_pack(**kwargs) runs
    ...
    tapenv.uninstall()

Replaces exactly the hot-path symbols: pack.update_dynamic, pack.update_mask, pack.reward,
tools.Container, tools.calc_positions_lb_greedy, tools.calc_positions_mcs (and generate.InitialContainer +
generate.generate_blocks + generate.generate_blocks_with_GT when `generate` is passed).  Everything else of the reference keeps running as it is.  The two whole-episode functions fall back to the saved
reference function for shapes beyond the compiled limits (e.g. the 7x7 initial container of the 3D generators has 49 cells,
tapenv_limits.max_cells_3d is 32) -- the dataset generators call them with containers the network never sees."""
import functools
import sys

from . import _capi, containers, episode, generators, ops, rolling

_saved = []
_tools = None          # the reference's `tools` module while installed (Container.draw_container hands its drawing to it)


def installed_tools():
    return _tools


def _with_fallback(ours, original):
    """ours(...) unless the configuration is outside the compiled limits / unsupported; then the reference's own function."""
    if original is None:
        return ours

    @functools.wraps(ours)
    def call(blocks, container_size, reward_type):
        try:
            return ours(blocks, container_size, reward_type)
        except _capi.TapEnvError as e:
            if e.code in (_capi.ELIMIT, _capi.EUNSUPPORTED):
                return original(blocks, container_size, reward_type)
            raise
    call.tapenv_original = original
    return call


def install(pack=None, tools=None, generate=None):
    """generate: the reference's `generate` module -- also replaces generate.InitialContainer (the rolling window,
    rolling.py:501) by the GPU-backed per-instance class."""
    pack = pack if pack is not None else sys.modules.get("pack")
    tools = tools if tools is not None else sys.modules.get("tools")
    if pack is None and tools is None:
        raise RuntimeError("tapenv.install: import the reference's `pack` / `tools` modules first (or pass them)")
    global _tools
    _tools = tools
    repl = []
    if pack is not None:
        repl += [(pack, "update_dynamic", ops.update_dynamic), (pack, "update_mask", ops.update_mask),
                 (pack, "reward", episode.reward)]
    if tools is not None:
        repl += [(tools, "Container", containers.Container),
                 (tools, "calc_positions_lb_greedy",
                  _with_fallback(episode.calc_positions_lb_greedy, getattr(tools, "calc_positions_lb_greedy", None))),
                 (tools, "calc_positions_mcs",
                  _with_fallback(episode.calc_positions_mcs, getattr(tools, "calc_positions_mcs", None)))]
    if generate is not None:
        generators._original = getattr(generate, "generate_blocks", None)
        generators._original_gt = getattr(generate, "generate_blocks_with_GT", None)
        generators._generate = generate
        repl += [(generate, "InitialContainer", rolling.InitialContainer),
                 (generate, "generate_blocks", generators.generate_blocks)]
        if generators._original_gt is not None:
            repl.append((generate, "generate_blocks_with_GT", generators.generate_blocks_with_GT))
    for mod, name, new in repl:
        _saved.append((mod, name, getattr(mod, name, None)))
        setattr(mod, name, new)
    return [name for _, name, _ in repl]


def uninstall():
    global _tools
    _tools = None
    generators._original = generators._original_gt = generators._generate = None
    while _saved:
        mod, name, old = _saved.pop()
        if old is None:
            delattr(mod, name)
        else:
            setattr(mod, name, old)
