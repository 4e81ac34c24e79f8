"""TEST INFRASTRUCTURE ONLY -- ctypes face of oracle/tap_oracle.c.

The CPU oracle is the checker for the CUDA path.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs
may import this module; the product package never does.

Mirrors the reference's names so the parity tests read like reference usage:
``Container`` (tools.py:3603), ``update_dynamic`` / ``update_mask``
(pack.py:333 / :276), ``initial_mask`` (model.py:297-307) and a batched
``episode_batch`` driver used as the timed CPU baseline.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libtap_oracle.so")

LB_GREEDY, MACS = 0, 1
HM_TYPES = {"full": 0, "zero": 1, "diff": 2}
STRATEGIES = {"LB_GREEDY": 0, "MACS": 1, "MUL": 1, "LB": 2}


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("tap_oracle.c", "win_oracle.c", "Makefile")]
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        p = C.c_void_p
        L.tapo_env_new.restype = p
        L.tapo_env_new.argtypes = [C.c_int] * 5 + [C.c_char_p, C.c_int, C.c_int]
        L.tapo_env_free.argtypes = [p]
        L.tapo_env_clear.argtypes = [p]
        L.tapo_env_add_new_block.argtypes = [p, p, p]
        L.tapo_env_add_new_block.restype = C.c_int
        L.tapo_env_encode_heightmap.argtypes = [p, p]
        L.tapo_env_calc_ratio.argtypes = [p]
        L.tapo_env_calc_ratio.restype = C.c_double
        L.tapo_env_calc_cps.argtypes = [p, p, p, p]
        for name in ("k", "error", "strategy"):
            getattr(L, "tapo_env_" + name).argtypes = [p]
            getattr(L, "tapo_env_" + name).restype = C.c_int
        for name in ("valid", "empty"):
            getattr(L, "tapo_env_" + name).argtypes = [p]
            getattr(L, "tapo_env_" + name).restype = C.c_longlong
        for name in ("heightmap", "positions", "container", "stable"):
            getattr(L, "tapo_env_" + name).argtypes = [p]
            getattr(L, "tapo_env_" + name).restype = p
        L.tapo_env_lfs3.argtypes = [p, C.c_int, C.c_int, p]
        L.tapo_env_lfs3.restype = C.c_int
        L.tapo_env_lfs.argtypes = [p, C.c_int, p]
        L.tapo_env_lfs.restype = C.c_int
        L.tapo_update_dynamic.argtypes = [p, p, p] + [C.c_int] * 6 + [p]
        L.tapo_update_mask.argtypes = [p, p, p] + [C.c_int] * 5 + [p, p]
        L.tapo_is_stable_3d_mask.argtypes = [C.c_int, C.c_int, p]
        L.tapo_is_stable_3d_mask.restype = C.c_int
        L.tapo_is_stable_3d_masks.argtypes = [C.c_int, C.c_int, p, C.c_int, p]
        L.tapo_episode_batch.argtypes = ([C.c_int] * 6 + [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int]
                                         + [p] * 11 + [C.c_int, C.c_int, C.c_int])
        L.tapo_episode_batch.restype = C.c_int
        L.tapo_pyset_order.argtypes = [p, C.c_int, p]
        L.tapo_pyset_order.restype = C.c_int
        L.tapo_win_new.argtypes = [C.c_int, C.c_int, C.c_int, p, p, C.c_int]
        L.tapo_win_new.restype = p
        L.tapo_win_free.argtypes = [p]
        L.tapo_win_convert_to_input.argtypes = [p, p, p]
        L.tapo_win_convert_to_input.restype = C.c_int
        L.tapo_win_remove_block.argtypes = [p, C.c_int]
        L.tapo_win_is_last_graph.argtypes = [p]
        L.tapo_win_is_last_graph.restype = C.c_int
        L.tapo_win_nodes.argtypes = [p, p]
        L.tapo_win_nodes.restype = C.c_int
        L.tapo_win_error.argtypes = [p]
        L.tapo_win_error.restype = C.c_int
        L.tapo_rolling_batch.argtypes = ([C.c_int] * 6 + [C.c_char_p] + [C.c_int] * 4 + [p] * 5 + [C.c_int])
        L.tapo_rolling_batch.restype = C.c_int
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Container(object):
    """tools.Container (tools.py:3603-3966) backed by the C restatement."""

    def __init__(self, container_size, blocks_num, reward_type, heightmap_type="full",
                 initial_container_size=None, max_height=None, packing_strategy="LB_GREEDY"):
        self.container_size = list(container_size)
        self.block_dim = len(container_size)
        self.blocks_num = blocks_num
        self.reward_type = reward_type
        self.heightmap_type = heightmap_type
        W = int(container_size[0])
        Ln = int(container_size[1]) if self.block_dim == 3 else 1
        H = int(container_size[-1])
        self._W, self._L, self._H = W, Ln, H
        self._h = lib().tapo_env_new(self.block_dim, W, Ln, H, blocks_num, reward_type.encode(),
                                     HM_TYPES[heightmap_type], STRATEGIES[packing_strategy])
        if not self._h:
            raise ValueError("oracle limits exceeded")
        self.packing_strategy = "MACS" if lib().tapo_env_strategy(self._h) == MACS else packing_strategy

    def __del__(self):
        if getattr(self, "_h", None):
            lib().tapo_env_free(self._h)
            self._h = None

    def _enc_len(self):
        W, Ln = self._W, self._L
        if self.block_dim == 2:
            return W - 1 if self.heightmap_type == "diff" else W
        return 2 * W * Ln if self.heightmap_type == "diff" else W * Ln

    def _shape_enc(self, buf):
        W, Ln = self._W, self._L
        if self.block_dim == 2:
            return buf
        if self.heightmap_type == "diff":
            return buf.reshape(2, W, Ln)
        return buf.reshape(W, Ln)

    def add_new_block(self, block, is_rotate=False):
        blk = np.ascontiguousarray(np.asarray(block, dtype=np.float32))
        out = np.zeros(self._enc_len() + 4, dtype=np.int32)
        r = lib().tapo_env_add_new_block(self._h, _ptr(blk), _ptr(out))
        if r < 0 or lib().tapo_env_error(self._h):
            raise IndexError("oracle: container overflow / limit (error %d)" % lib().tapo_env_error(self._h))
        return self._shape_enc(out[: self._enc_len()].astype(np.int64))

    def get_heightmap(self, is_full=None):
        if is_full is not None:
            return self.heightmap
        out = np.zeros(self._enc_len() + 4, dtype=np.int32)
        lib().tapo_env_encode_heightmap(self._h, _ptr(out))
        return self._shape_enc(out[: self._enc_len()].astype(np.int64))

    def clear_container(self):
        lib().tapo_env_clear(self._h)

    def calc_ratio(self):
        return lib().tapo_env_calc_ratio(self._h)

    def calc_CPS(self):
        c, p, s = C.c_double(), C.c_double(), C.c_double()
        lib().tapo_env_calc_cps(self._h, C.byref(c), C.byref(p), C.byref(s))
        return c.value, p.value, s.value

    def _arr(self, name, count, ctype, dtype):
        addr = getattr(lib(), "tapo_env_" + name)(self._h)
        return np.ctypeslib.as_array(C.cast(addr, C.POINTER(ctype)), shape=(count,)).astype(dtype)

    @property
    def heightmap(self):
        a = self._arr("heightmap", self._W * self._L, C.c_int, np.int64)
        return a if self.block_dim == 2 else a.reshape(self._W, self._L)

    @property
    def positions(self):
        return self._arr("positions", self.blocks_num * self.block_dim, C.c_int, np.int64).reshape(
            self.blocks_num, self.block_dim)

    @property
    def container(self):
        a = self._arr("container", self._W * self._L * self._H, C.c_int, np.int64)
        return a.reshape(self._W, self._H) if self.block_dim == 2 else a.reshape(self._W, self._L, self._H)

    @property
    def stable(self):
        return [bool(v) for v in self._arr("stable", self.blocks_num, C.c_ubyte, np.uint8)]

    @property
    def valid_size(self):
        return int(lib().tapo_env_valid(self._h))

    @property
    def empty_size(self):
        return int(lib().tapo_env_empty(self._h))

    @property
    def current_blocks_num(self):
        return int(lib().tapo_env_k(self._h))

    @property
    def level_free_space(self):
        out = np.zeros(64, dtype=np.int32)
        res = []
        if self.block_dim == 3:                  # [z][y] -> x interval list (tools.py:3644-3648)
            for z in range(self._H):
                row = []
                for y in range(self._L):
                    n = lib().tapo_env_lfs3(self._h, z, y, _ptr(out))
                    row.append(out[:n].tolist())
                res.append(row)
            return res
        for z in range(self._H):
            n = lib().tapo_env_lfs(self._h, z, _ptr(out))
            res.append(out[:n].tolist())
        return res


def rotate_types(dim, allow_rot=True):
    import math
    return math.factorial(dim) if allow_rot else 1


def update_dynamic(dynamic, static, chosen_idx, input_type="bot", allow_rot=True):
    """pack.update_dynamic (pack.py:333-376) on numpy arrays."""
    dynamic = np.ascontiguousarray(dynamic, dtype=np.float32)
    static = np.ascontiguousarray(static, dtype=np.float32)
    ptr = np.ascontiguousarray(chosen_idx, dtype=np.int64)
    B, rows, S = dynamic.shape
    srows = static.shape[1]
    dim = srows - 2 if input_type in ("mul", "mul-with") else srows - 1
    n = S // rotate_types(dim, allow_rot)
    update_time = 1 if input_type in ("simple", "rot", "rot-old") else 3
    out = np.empty_like(dynamic)
    lib().tapo_update_dynamic(_ptr(dynamic), _ptr(static), _ptr(ptr), B, rows, S, srows, n, update_time, _ptr(out))
    return out


def update_mask(mask, dynamic, static, chosen_idx, input_type="bot", allow_rot=True):
    """pack.update_mask (pack.py:276-331) -> (new_mask, chosen_mask)."""
    mask = np.ascontiguousarray(mask, dtype=np.float32)
    dynamic = np.ascontiguousarray(dynamic, dtype=np.float32)
    ptr = np.ascontiguousarray(chosen_idx, dtype=np.int64)
    B, rows, S = dynamic.shape
    srows = static.shape[1]
    dim = srows - 2 if input_type in ("mul", "mul-with") else srows - 1
    R = rotate_types(dim, allow_rot)
    n = S // R
    new_mask = np.empty_like(mask)
    chosen = np.empty_like(mask)
    lib().tapo_update_mask(_ptr(mask), _ptr(dynamic), _ptr(ptr), B, rows, S, n, R, _ptr(new_mask), _ptr(chosen))
    return new_mask, chosen


def initial_mask(dynamic, n, R):
    """model.py:297-307: accessibility mask at t=0 (mask = ones)."""
    dynamic = np.ascontiguousarray(dynamic, dtype=np.float32)
    B, rows, S = dynamic.shape
    new_mask = np.empty((B, S), dtype=np.float32)
    chosen = np.empty((B, S), dtype=np.float32)
    lib().tapo_update_mask(None, _ptr(dynamic), None, B, rows, S, n, R, _ptr(new_mask), _ptr(chosen))
    return new_mask


def episode_batch(static, dynamic, ptr_seq, container_size, reward_type, heightmap_type="diff",
                  packing_strategy="LB_GREEDY", nthreads=1, want=("heightmap", "positions", "stable", "reward",
                                                                  "cur_mask", "mask", "dynamic", "dec_dyn"),
                  capacity=None):
    """Run whole episodes for a batch on the CPU oracle (the timed CPU baseline).

    static [B,1+dim,S] f32, dynamic [B,3n,S] f32, ptr_seq [steps,B] int64 -- or, rolling-style, the same with a
    leading window axis ([Wn,B,...], [Wn,B,...], [Wn,steps,B]) and `capacity` blocks per container."""
    static = np.ascontiguousarray(static, dtype=np.float32)
    dynamic = np.ascontiguousarray(dynamic, dtype=np.float32)
    ptr_seq = np.ascontiguousarray(ptr_seq, dtype=np.int64)
    if static.ndim == 3:
        static, dynamic, ptr_seq = static[None], dynamic[None], ptr_seq[None]
    nwin, B, srows, S = static.shape
    dim = srows - 1
    R = rotate_types(dim)
    n = S // R
    steps = ptr_seq.shape[1]
    cap = int(capacity) if capacity is not None else max(n, nwin * steps)
    W = int(container_size[0])
    Ln = int(container_size[1]) if dim == 3 else 1
    H = int(container_size[-1])
    cells = W * Ln
    hm_t = HM_TYPES[heightmap_type]
    enc = (W - 1 if hm_t == 2 else W) if dim == 2 else (2 * cells if hm_t == 2 else cells)
    o = {}
    if "heightmap" in want: o["heightmap"] = np.zeros((B, cells), np.int32)
    if "positions" in want: o["positions"] = np.zeros((B, cap, dim), np.int32)
    if "stable" in want: o["stable"] = np.zeros((B, cap), np.uint8)
    if "reward" in want: o["reward"] = np.zeros((B,), np.float32)
    if "cur_mask" in want: o["cur_mask"] = np.zeros((B, S), np.float32)
    if "mask" in want: o["mask"] = np.zeros((B, S), np.float32)
    if "dynamic" in want: o["dynamic"] = np.zeros((B, 3 * n, S), np.float32)
    if "dec_dyn" in want: o["dec_dyn"] = np.zeros((B, enc), np.int32)
    st = lib().tapo_episode_batch(dim, W, Ln, H, n, R, reward_type.encode(), hm_t, STRATEGIES[packing_strategy],
                                  B, steps, _ptr(static), _ptr(dynamic), _ptr(ptr_seq),
                                  _ptr(o.get("heightmap")), _ptr(o.get("positions")), _ptr(o.get("stable")),
                                  _ptr(o.get("reward")), _ptr(o.get("cur_mask")), _ptr(o.get("mask")),
                                  _ptr(o.get("dynamic")), _ptr(o.get("dec_dyn")), int(nthreads), int(nwin), int(cap))
    o["status"] = st
    return o


def is_stable_3d_mask(bx, by, occ):
    occ = np.ascontiguousarray(occ, dtype=np.uint8)
    return bool(lib().tapo_is_stable_3d_mask(bx, by, _ptr(occ)))


def is_stable_3d_masks(bx, by, masks):
    """Bulk tools.is_stable on support bitmasks (bit x*by+y) -> uint8 array."""
    masks = np.ascontiguousarray(masks, dtype=np.uint32)
    out = np.zeros(masks.shape[0], dtype=np.uint8)
    lib().tapo_is_stable_3d_masks(bx, by, _ptr(masks), masks.shape[0], _ptr(out))
    return out


# ----------------------------------------------------------------------------------------------
# rolling window (generate.InitialContainer, generate.py:1589-1825) -- win_oracle.c
# ----------------------------------------------------------------------------------------------
WINDOW_ORDER_REFERENCE, WINDOW_ORDER_SORTED = 0, 1


def pyset_order(keys):
    """Iteration order of set(keys) for small non-negative ints, restated from CPython's setobject.c."""
    keys = np.ascontiguousarray(keys, dtype=np.int32)
    out = np.zeros(max(len(keys), 1), np.int32)
    cnt = lib().tapo_pyset_order(_ptr(keys), len(keys), _ptr(out))
    assert cnt >= 0
    return out[:cnt].tolist()


class InitialContainer(object):
    """generate.InitialContainer restricted to what rolling.validate drives: convert_to_input / remove_block /
    is_last_graph / sub_graph_nodes.  Built from the five precedence graphs as [5,T,T] 0/1 matrices
    (adj[g,u,v] = edge u -> v: move, left, right, forward, backward) and the [R*T,dim] rotation-major blocks."""

    def __init__(self, adj, blocks, blocks_num, child_graph_size, block_dim, order=WINDOW_ORDER_REFERENCE):
        adj = np.ascontiguousarray(adj, dtype=np.uint8)
        blocks = np.ascontiguousarray(blocks, dtype=np.int32)
        self.T, self.n, self.dim = int(blocks_num), int(child_graph_size), int(block_dim)
        self.R = rotate_types(self.dim)
        assert adj.shape == (5, self.T, self.T) and blocks.shape == (self.R * self.T, self.dim)
        self._h = lib().tapo_win_new(self.T, self.n, self.dim, _ptr(adj), _ptr(blocks), int(order))
        if not self._h:
            raise ValueError("window oracle limits")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().tapo_win_free(self._h)
            self._h = None

    def convert_to_input(self):
        S = self.n * self.R
        static = np.zeros((1 + self.dim, S), np.float32)
        dynamic = np.zeros((3 * self.n, S), np.float32)
        rc = lib().tapo_win_convert_to_input(self._h, _ptr(static), _ptr(dynamic))
        if rc:
            raise ValueError("window cannot be filled (the reference raises in np.concatenate)")
        return static, dynamic

    def remove_block(self, block_id):
        lib().tapo_win_remove_block(self._h, int(block_id))

    def is_last_graph(self):
        return bool(lib().tapo_win_is_last_graph(self._h))

    @property
    def sub_graph_nodes(self):
        out = np.zeros(self.T, np.int32)
        k = lib().tapo_win_nodes(self._h, _ptr(out))
        return out[:k].tolist()

    @property
    def error(self):
        return lib().tapo_win_error(self._h)


def rolling_batch(adj, blocks, ptr_seq, container_size, window, reward_type, heightmap_type="diff",
                  packing_strategy="LB_GREEDY", order=WINDOW_ORDER_REFERENCE, nthreads=1):
    """rolling.validate's loop (rolling.py:575-640) for a batch, network replaced by the recorded ptr_seq [T,B]:
    adj [B,5,T,T] u8, blocks [B,R*T,dim] i32.  -> dict(reward f32 [B], heightmap i32 [B,cells], status)."""
    adj = np.ascontiguousarray(adj, dtype=np.uint8)
    blocks = np.ascontiguousarray(blocks, dtype=np.int32)
    ptr_seq = np.ascontiguousarray(ptr_seq, dtype=np.int64)
    B, _, T, _ = adj.shape
    dim = blocks.shape[2]
    W = int(container_size[0]); Ln = int(container_size[1]) if dim == 3 else 1; H = int(container_size[-1])
    reward = np.zeros(B, np.float32)
    hm = np.zeros((B, W * Ln), np.int32)
    st = lib().tapo_rolling_batch(dim, W, Ln, H, T, int(window), reward_type.encode(), HM_TYPES[heightmap_type],
                                  STRATEGIES[packing_strategy], int(order), B, _ptr(adj), _ptr(blocks), _ptr(ptr_seq),
                                  _ptr(reward), _ptr(hm), int(nthreads))
    return dict(reward=reward, heightmap=hm, status=st)
