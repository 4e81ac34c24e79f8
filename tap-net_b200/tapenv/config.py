"""Host-side configuration of the packing environment: the reference's string-typed options
(tools.Container.__init__ tools.py:3611-3661, pack.update_* pack.py:285-309) -> struct tapenv_config."""
import ctypes as C
import math

from . import _capi

_cache = {}


def rotate_types(dim, allow_rot=True):
    """pack.py:306-309: dim! rotations when allow_rot else 1."""
    return math.factorial(dim) if allow_rot else 1


def make_config(batch, blocks_num, container_size, reward_type="C+P+S-lb-soft", heightmap_type="diff",
                packing_strategy="LB_GREEDY", input_type="bot", allow_rot=True, capacity=None):
    """Returns a validated _capi.Config.  Raises TapEnvError (a ValueError) on unknown enum strings or
    shapes beyond the compiled limits -- the reference would print '... OHHH' and die on a NameError."""
    size = tuple(int(v) for v in container_size)
    key = (int(batch), int(blocks_num), size, reward_type, heightmap_type, packing_strategy, input_type, bool(allow_rot),
           None if capacity is None else int(capacity))
    cfg = _cache.get(key)
    if cfg is None:
        cfg = _capi.Config()
        arr = (C.c_int32 * len(size))(*size)
        _capi.check(_capi.lib.tapenv_config_init(C.byref(cfg), key[0], key[1], len(size), 1 if allow_rot else 0, arr,
                                                 reward_type.encode(), packing_strategy.encode(),
                                                 heightmap_type.encode(), input_type.encode()), "config")
        if capacity is not None:                       # the container outlives the network window (rolling.py:702-703)
            cfg.capacity = int(capacity)
            _capi.check(_capi.lib.tapenv_config_check(C.byref(cfg)), "config")
        if len(_cache) > 256:
            _cache.clear()
        _cache[key] = cfg
    return cfg


def tensor_config(batch, static_rows, dyn_rows, S, input_type, allow_rot):
    """Config for the container-less tensor ops (update_dynamic / update_mask): the block dimension and
    blocks_num are derived from the tensor shapes exactly as pack.py:285-311 / :338-368 does."""
    if input_type in ("mul", "mul-with"):
        dim = static_rows - 2
    else:
        dim = static_rows - 1
    if dim not in (2, 3):
        raise _capi.TapEnvError(_capi.ESHAPE, "static has %d rows" % static_rows)
    R = rotate_types(dim, allow_rot)
    n = int(S / R)                                     # pack.py:311 int(dynamic.shape[-1] / rotate_types)
    if n * R != S:
        raise _capi.TapEnvError(_capi.ESHAPE, "S=%d is not blocks_num*rotate_types" % S)
    size = (1, 1) if dim == 2 else (1, 1, 1)
    cfg = make_config(batch, n, size, "C+P+S-lb-soft", "full", "LB_GREEDY", input_type, allow_rot)
    if cfg.dyn_rows != dyn_rows or cfg.static_rows != static_rows:
        raise _capi.TapEnvError(_capi.ESHAPE, "dynamic has %d rows, input_type %r needs %d" % (dyn_rows, input_type, cfg.dyn_rows))
    return cfg
