"""ctypes binding of include/tapenv.h -- the only way the Python side reaches the kernels.

There is deliberately NO fallback: if lib/libtapenv.so is missing or does not load,
importing this module raises, and so does every operator built on it.
"""
import ctypes as C
import os

from . import build as _build

c_void_p, c_int, c_int32, c_size_t, c_char_p = C.c_void_p, C.c_int, C.c_int32, C.c_size_t, C.c_char_p


class Config(C.Structure):  # struct tapenv_config
    _fields_ = [(n, c_int32) for n in (
        "batch", "blocks_num", "dim", "rotate_types", "width", "length", "height", "strategy",
        "heightmap_type", "reward_flags", "ratio_mode", "static_rows", "dyn_rows", "update_time", "capacity")]


class StateLayout(C.Structure):  # struct tapenv_state_layout
    _fields_ = [(n, c_size_t) for n in ("scalars", "heightmap", "positions", "blocks", "stable", "flags", "voxels", "lists", "pending", "total")]


class PeerComm(C.Structure):  # struct tapenv_peer_comm
    _fields_ = [("world", c_int32), ("rank", c_int32), ("peer", c_void_p * 8)]


class WindowConfig(C.Structure):  # struct tapenv_window_config
    _fields_ = [(n, c_int32) for n in ("batch", "total_blocks", "window", "dim", "rotate_types", "node_order", "blocks_are_rotations")]


class Limits(C.Structure):  # struct tapenv_limits
    _fields_ = [(n, c_int32) for n in ("max_width_2d", "max_cells_3d", "max_candidates", "max_blocks")]


OK, EINVAL, EENUM, ELIMIT, ESHAPE, ECUDA, EUNSUPPORTED = 0, -1, -2, -3, -4, -5, -6
LB_GREEDY, MACS, LB = 0, 1, 2
WINDOW_ORDER_REFERENCE, WINDOW_ORDER_SORTED = 0, 1

# every symbol include/tapenv.h declares: name -> (restype, argtypes)
P = c_void_p
CFG = C.POINTER(Config)
SYMBOLS = {
    "tapenv_version": (c_int, []),
    "tapenv_strerror": (c_char_p, [c_int]),
    "tapenv_get_limits": (None, [C.POINTER(Limits)]),
    "tapenv_config_init": (c_int, [CFG, c_int32, c_int32, c_int32, c_int32, C.POINTER(c_int32),
                                   c_char_p, c_char_p, c_char_p, c_char_p]),
    "tapenv_config_check": (c_int, [CFG]),
    "tapenv_state_bytes": (c_size_t, [CFG]),
    "tapenv_state_get_layout": (c_int, [CFG, C.POINTER(StateLayout)]),
    "tapenv_encoded_heightmap_len": (c_int32, [CFG]),
    "tapenv_reset": (c_int, [CFG, P, P, P, P, P]),
    "tapenv_initial_mask": (c_int, [CFG, P, P, P, P]),
    "tapenv_update_dynamic": (c_int, [CFG, P, P, P, P, P]),
    "tapenv_update_mask": (c_int, [CFG, P, P, P, P, P, P]),
    "tapenv_add_blocks": (c_int, [CFG, P, P, P, P]),
    "tapenv_step": (c_int, [CFG, P, P, P, P, P, P, P, P, P, P, P]),
    "tapenv_step_reward": (c_int, [CFG, P, P, P, P, P, P, P, P, P, P, P, P]),
    "tapenv_reward_sums": (c_int, [CFG, P, P, P, C.POINTER(PeerComm), P]),
    "tapenv_reward": (c_int, [CFG, P, P, P, P]),
    "tapenv_comm_bytes": (c_size_t, []),
    "tapenv_comm_status_offset": (c_size_t, []),
    "tapenv_reward_allreduce": (c_int, [CFG, P, P, P, P, C.POINTER(PeerComm), P]),
    "tapenv_episode": (c_int, [CFG, P, P, P, P, c_int32, P, P, P, P, P]),
    "tapenv_packed_words": (c_int32, [CFG]),
    "tapenv_reset_packed": (c_int, [CFG, P, P, P, P, P, P, P, P]),
    "tapenv_step_mul": (c_int, [CFG] + [P] * 10 + [c_int32, P, P]),
    "tapenv_add_blocks_mul": (c_int, [CFG] + [P] * 6),
    "tapenv_reward_mul": (c_int, [CFG, P, P, P, P]),
    "tapenv_window_state_bytes": (c_size_t, [C.POINTER(WindowConfig)]),
    "tapenv_window_reset": (c_int, [C.POINTER(WindowConfig), P, P]),
    "tapenv_window_next": (c_int, [C.POINTER(WindowConfig)] + [P] * 11),
    "tapenv_rolling_step": (c_int, [CFG, P, C.POINTER(WindowConfig)] + [P] * 13),
}


def _load():
    path = os.environ.get("TAPENV_LIB") or _build.LIB_PATH      # TAPENV_LIB: a tuning variant built by build.build(out=...)
    if not os.path.exists(path):
        raise ImportError(
            "tapenv: %s is missing -- build it with `python __graft_entry__.py build` "
            "(nvcc, sm_100a). There is no CPU fallback." % path)
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)       # AttributeError if the library does not export the symbol
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def limits():
    out = Limits()
    lib.tapenv_get_limits(C.byref(out))
    return out


class TapEnvError(ValueError):
    def __init__(self, code, where=""):
        self.code = code
        msg = lib.tapenv_strerror(code).decode()
        ValueError.__init__(self, "tapenv%s: %s (code %d)" % (" " + where if where else "", msg, code))


def check(code, where=""):
    if code != OK:
        raise TapEnvError(code, where)
