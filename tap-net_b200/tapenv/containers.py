"""Batched packing containers: the B200 replacement of `[tools.Container(...) for _ in range(B)]`
(model.py:294) and of the per-environment Python loop at model.py:452-453 / :509-510.

`BatchedContainers` owns one opaque HBM state buffer for B environments and drives the kernels through
the C ABI (include/tapenv.h).  `Container` keeps the reference's per-environment class signature
(tools.py:3607-3966) as a thin view of one row, so an UNMODIFIED model.py can be pointed at it.
"""
import ctypes as C
import threading

import numpy as np
import torch

from . import _capi
from .config import make_config, rotate_types
from .ops import _dev, _p, _stream


class BatchedContainers(object):
    """B independent containers resident on one GPU.

    Constructor arguments follow tools.Container (tools.py:3611) + batch_size/device.
    `initial_container_size` / `max_height` are accepted and ignored exactly like the reference ignores
    them for the strategies built here (they only feed the two-container drawing code)."""

    def __init__(self, container_size, blocks_num, reward_type, heightmap_type="full", initial_container_size=None,
                 max_height=None, packing_strategy="LB_GREEDY", batch_size=1, device=None, input_type="bot",
                 allow_rot=True, window=None):
        if not torch.cuda.is_available():
            raise RuntimeError("tapenv: a CUDA device is required (no CPU fallback exists)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.container_size = [int(v) for v in container_size]
        self.block_dim = len(self.container_size)
        self.blocks_num = int(blocks_num)           # capacity of one container (tools.py:3611 blocks_num)
        self.batch_size = int(batch_size)
        self.reward_type = reward_type
        self.heightmap_type = heightmap_type
        # `window`: blocks the NETWORK sees per decode step (rows/columns of static/dynamic/mask).  It equals blocks_num
        # in training; rolling inference keeps one container of total_blocks_num behind a window of net_blocks_num
        # (rolling.py:702-703).  A container used without tensors (add_new_blocks only) may exceed the tensor limits.
        R = rotate_types(len(self.container_size), allow_rot)
        if window is None:
            window = self.blocks_num if self.blocks_num * R <= _capi.limits().max_candidates else 1
        self.window = int(window)
        self.cfg = make_config(batch_size, self.window, container_size, reward_type, heightmap_type, packing_strategy,
                               input_type, allow_rot, capacity=self.blocks_num)
        self.packing_strategy = "MACS" if self.cfg.strategy == _capi.MACS else packing_strategy
        self.S = self.cfg.blocks_num * self.cfg.rotate_types
        self.enc_len = int(_capi.lib.tapenv_encoded_heightmap_len(C.byref(self.cfg)))
        lay = _capi.StateLayout()
        _capi.check(_capi.lib.tapenv_state_get_layout(C.byref(self.cfg), C.byref(lay)), "layout")
        self._layout = lay
        self.state = torch.empty(max(int(lay.total), 1), dtype=torch.uint8, device=self.device)
        self._version = 0
        self.clear_container()

    # ---- state views (no copies) -------------------------------------------------------
    def _view(self, off, count, dtype, shape):
        item = torch.empty((), dtype=dtype).element_size()
        return self.state[off: off + count * item].view(dtype).view(shape)

    @property
    def _cells(self):
        return self.container_size[0] if self.block_dim == 2 else self.container_size[0] * self.container_size[1]

    @property
    def heightmap(self):
        """int32 [B,W] or [B,W,L] (tools.py:3630)."""
        B = self.batch_size
        shape = (B, self.container_size[0]) if self.block_dim == 2 else (B, self.container_size[0], self.container_size[1])
        return self._view(self._layout.heightmap, B * self._cells, torch.int32, shape)

    @property
    def scalars(self):
        """int32 [B,4]: valid_size, empty_size, number of stable blocks, current_blocks_num."""
        return self._view(self._layout.scalars, self.batch_size * 4, torch.int32, (self.batch_size, 4))

    @property
    def valid_size(self):
        return self.scalars[:, 0]

    @property
    def empty_size(self):
        return self.scalars[:, 1]

    @property
    def current_blocks_num(self):
        return self.scalars[:, 3]

    @property
    def positions(self):
        B, n, d = self.batch_size, self.blocks_num, self.block_dim
        return self._view(self._layout.positions, B * n * d, torch.int32, (B, n, d))

    @property
    def blocks(self):
        B, n, d = self.batch_size, self.blocks_num, self.block_dim
        return self._view(self._layout.blocks, B * n * d, torch.int32, (B, n, d))

    @property
    def stable(self):
        B, n = self.batch_size, self.blocks_num
        return self._view(self._layout.stable, B * n, torch.uint8, (B, n))

    @property
    def flags(self):
        """int32 [B] sticky anomaly bits (1: a stack passed container height -- NumPy would have raised)."""
        return self._view(self._layout.flags, self.batch_size, torch.int32, (self.batch_size,))

    def _shape_enc(self, t):
        if self.block_dim == 3:
            W, L = self.container_size[0], self.container_size[1]
            return t.view(self.batch_size, 2, W, L) if self.heightmap_type == "diff" else t.view(self.batch_size, W, L)
        return t

    # ---- operations --------------------------------------------------------------------
    def clear_container(self):
        """Container.clear_container for every environment (tools.py:3858-3885)."""
        self._version += 1
        with torch.cuda.device(self.device):
            _capi.check(_capi.lib.tapenv_reset(C.byref(self.cfg), _p(self.state), None, None, None, _stream()), "reset")

    def reset(self, dynamic):
        """clear + the initial accessibility masks of model.py:297-307 -> (current_mask, mask)."""
        dynamic = _dev(dynamic, "dynamic", torch.float32)
        B = self.batch_size
        if tuple(dynamic.shape) != (B, self.cfg.dyn_rows, self.S):
            raise _capi.TapEnvError(_capi.ESHAPE, "dynamic %s" % (tuple(dynamic.shape),))
        cur = torch.empty(B, self.S, dtype=torch.float32, device=self.device)
        mask = torch.empty_like(cur)
        self._version += 1
        with torch.cuda.device(self.device):
            _capi.check(_capi.lib.tapenv_reset(C.byref(self.cfg), _p(self.state), _p(dynamic), _p(cur), _p(mask),
                                               _stream()), "reset")
        return cur, mask

    def reset_packed(self, static_u8, dynamic_bits, out=None):
        """reset() for inputs uploaded in the packed host format (tapenv.pack_inputs: u8 static, bit-row dynamic):
        expands them to the fp32 tensors, clears the containers, emits the initial masks -- one launch.
        -> (static f32 [B,rows,S], dynamic f32 [B,3n,S], current_mask, mask)."""
        static_u8 = _dev(static_u8, "static_u8", torch.uint8)
        dynamic_bits = _dev(dynamic_bits, "dynamic_bits", torch.int32)
        B, S = self.batch_size, self.S
        words = int(_capi.lib.tapenv_packed_words(C.byref(self.cfg)))
        if tuple(static_u8.shape) != (B, self.cfg.static_rows, S) or tuple(dynamic_bits.shape) != (B, words):
            raise _capi.TapEnvError(_capi.ESHAPE, "packed inputs %s %s" % (tuple(static_u8.shape), tuple(dynamic_bits.shape)))
        f32 = dict(dtype=torch.float32, device=self.device)
        if out is None:
            static, dynamic = torch.empty(B, self.cfg.static_rows, S, **f32), torch.empty(B, self.cfg.dyn_rows, S, **f32)
            cur, mask = torch.empty(B, S, **f32), torch.empty(B, S, **f32)
        else:
            static, dynamic, cur, mask = out
        self._version += 1
        with torch.cuda.device(self.device):
            _capi.check(_capi.lib.tapenv_reset_packed(C.byref(self.cfg), _p(self.state), _p(static_u8), _p(dynamic_bits),
                                                      _p(static), _p(dynamic), _p(cur), _p(mask), _stream()), "reset_packed")
        return static, dynamic, cur, mask

    def initial_mask(self, dynamic):
        """The accessibility masks of a freshly (re)filled window, container untouched (model.py:297-307,
        rolling.py:325-335) -> (current_mask, mask)."""
        dynamic = _dev(dynamic, "dynamic", torch.float32)
        B = self.batch_size
        if tuple(dynamic.shape) != (B, self.cfg.dyn_rows, self.S):
            raise _capi.TapEnvError(_capi.ESHAPE, "dynamic %s" % (tuple(dynamic.shape),))
        cur = torch.empty(B, self.S, dtype=torch.float32, device=self.device)
        mask = torch.empty_like(cur)
        with torch.cuda.device(self.device):
            _capi.check(_capi.lib.tapenv_initial_mask(C.byref(self.cfg), _p(dynamic), _p(cur), _p(mask), _stream()), "initial_mask")
        return cur, mask

    def add_new_blocks(self, blocks):
        """Batched Container.add_new_block (tools.py:3663-3744): blocks f32 [B,dim] ->
        encoded heightmaps f32 [B,enc] (2D) / [B,2,W,L] or [B,W,L] (3D), what model.py:456-463 builds."""
        blocks = _dev(blocks, "blocks", torch.float32)
        if tuple(blocks.shape) != (self.batch_size, self.block_dim):
            raise _capi.TapEnvError(_capi.ESHAPE, "blocks %s" % (tuple(blocks.shape),))
        out = torch.empty(self.batch_size, self.enc_len, dtype=torch.float32, device=self.device)
        self._version += 1
        with torch.cuda.device(self.device):
            _capi.check(_capi.lib.tapenv_add_blocks(C.byref(self.cfg), _p(self.state), _p(blocks), _p(out), _stream()),
                        "add_blocks")
        return self._shape_enc(out)

    def step(self, ptr, static, dynamic, mask, out=None, reward_out=None):
        """The fused decode-step transition (model.py:376-458 env part) in ONE launch:
        returns (dynamic', current_mask, mask', decoder_static [B,dim], decoder_dynamic).
        reward_out (f32 [B], optional): also receives calc_ratio() of the state this step leaves behind -- pass it on the
        LAST decode step instead of calling calc_ratio() afterwards (model.py:499-515)."""
        ptr = _dev(ptr, "ptr", torch.int64)
        static = _dev(static, "static", torch.float32)
        dynamic = _dev(dynamic, "dynamic", torch.float32)
        mask = _dev(mask, "mask", torch.float32)
        B, S = self.batch_size, self.S
        if tuple(dynamic.shape) != (B, self.cfg.dyn_rows, S) or tuple(static.shape) != (B, self.cfg.static_rows, S) \
                or tuple(mask.shape) != (B, S) or tuple(ptr.shape) != (B,):
            raise _capi.TapEnvError(_capi.ESHAPE, "step tensors")
        if out is None:
            dyn_out = torch.empty_like(dynamic)
            cur = torch.empty_like(mask)
            mask_out = torch.empty_like(mask)
            dec_static = torch.empty(B, self.cfg.static_rows - 1, dtype=torch.float32, device=self.device)
            dec_dyn = torch.empty(B, self.enc_len, dtype=torch.float32, device=self.device)
        else:
            dyn_out, cur, mask_out, dec_static, dec_dyn = out
        self._version += 1
        with torch.cuda.device(self.device):
            if reward_out is None:
                _capi.check(_capi.lib.tapenv_step(C.byref(self.cfg), _p(self.state), _p(ptr), _p(static), _p(dynamic),
                                                  _p(mask), _p(dyn_out), _p(cur), _p(mask_out), _p(dec_static), _p(dec_dyn),
                                                  _stream()), "step")
            else:
                if reward_out.dtype != torch.float32 or tuple(reward_out.shape) != (B,) or not reward_out.is_cuda:
                    raise _capi.TapEnvError(_capi.ESHAPE, "reward_out")
                _capi.check(_capi.lib.tapenv_step_reward(C.byref(self.cfg), _p(self.state), _p(ptr), _p(static), _p(dynamic),
                                                         _p(mask), _p(dyn_out), _p(cur), _p(mask_out), _p(dec_static),
                                                         _p(dec_dyn), _p(reward_out), _stream()), "step_reward")
        return dyn_out, cur, mask_out, dec_static, self._shape_enc(dec_dyn)

    def reward_sums(self, reward, exchange=None, out=None):
        """(sum r, sum r^2, B) of a reward vector produced by step(..., reward_out=) -> f64 [3]; with exchange=PeerExchange
        -> (local sums, global sums), the cross-GPU reduction in the same launch (see calc_ratio).
        out=(sums, total) preallocated f64 [3] buffers (total ignored without an exchange)."""
        if out is not None:
            sums, total = out[0], (out[1] if exchange is not None else None)
        else:
            sums = torch.empty(3, dtype=torch.float64, device=self.device)
            total = torch.empty(3, dtype=torch.float64, device=self.device) if exchange is not None else None
        with torch.cuda.device(self.device):
            _capi.check(_capi.lib.tapenv_reward_sums(C.byref(self.cfg), _p(reward), _p(sums), _p(total),
                                                     C.byref(exchange.comm) if exchange is not None else None, _stream()), "reward_sums")
        return sums if exchange is None else (sums, total)

    def calc_ratio(self, partial_sums=False, exchange=None):
        """Container.calc_ratio for every environment -> f32 [B] (tools.py:3908-3966, model.py:509-510).
        partial_sums=True also returns the f64 [3] (sum r, sum r^2, B) operand of the reward all-reduce.
        exchange=<tapenv.dist.PeerExchange>: the cross-GPU reduction is fused behind the reward kernel over NVLink
        peer memory; returns (reward, local sums, global sums) -- the global sums are identical on every rank."""
        r = torch.empty(self.batch_size, dtype=torch.float32, device=self.device)
        if exchange is not None:
            sums = torch.empty(3, dtype=torch.float64, device=self.device)
            total = torch.empty(3, dtype=torch.float64, device=self.device)
            with torch.cuda.device(self.device):
                _capi.check(_capi.lib.tapenv_reward_allreduce(C.byref(self.cfg), _p(self.state), _p(r), _p(sums), _p(total),
                                                              C.byref(exchange.comm), _stream()), "reward_allreduce")
            return r, sums, total
        sums = torch.empty(3, dtype=torch.float64, device=self.device) if partial_sums else None
        with torch.cuda.device(self.device):
            _capi.check(_capi.lib.tapenv_reward(C.byref(self.cfg), _p(self.state), _p(r), _p(sums), _stream()), "reward")
        return (r, sums) if partial_sums else r

    def episode(self, static, dynamic, ptr_seq):
        """Whole-episode entry: reset + len(ptr_seq) steps + calc_ratio in one launch.
        ptr_seq int64 [steps,B].  Returns (reward f32 [B], current_mask, mask, decoder_dynamic)."""
        static = _dev(static, "static", torch.float32)
        dynamic = _dev(dynamic, "dynamic", torch.float32)
        ptr_seq = _dev(ptr_seq, "ptr_seq", torch.int64)
        B, S = self.batch_size, self.S
        r = torch.empty(B, dtype=torch.float32, device=self.device)
        cur = torch.empty(B, S, dtype=torch.float32, device=self.device)
        mask = torch.empty_like(cur)
        dec_dyn = torch.empty(B, self.enc_len, dtype=torch.float32, device=self.device)
        self._version += 1
        with torch.cuda.device(self.device):
            _capi.check(_capi.lib.tapenv_episode(C.byref(self.cfg), _p(self.state), _p(static), _p(dynamic), _p(ptr_seq),
                                                 int(ptr_seq.shape[0]), _p(r), _p(cur), _p(mask), _p(dec_dyn), _stream()),
                        "episode")
        return r, cur, mask, self._shape_enc(dec_dyn)

    def check_flags(self):
        """Raise IndexError if any environment overflowed its container (the reference's NumPy would)."""
        bad = int((self.flags != 0).sum().item())
        if bad:
            raise IndexError("tapenv: %d environment(s) stacked beyond container height / block capacity" % bad)

    # ---- per-environment views -----------------------------------------------------------
    def __len__(self):
        return self.batch_size

    def __getitem__(self, b):
        if not -self.batch_size <= b < self.batch_size:
            raise IndexError(b)
        return Container._view_of(self, b % self.batch_size)

    def __iter__(self):
        return (Container._view_of(self, b) for b in range(self.batch_size))


class BatchedContainerPairs(object):
    """The two container lists model.py builds for input_type 'mul' / 'mul-with' (model.py:286-292:
    containers_a / containers_b), as two BatchedContainers of one configuration.  Every candidate column of `static`
    names its target container in the last row (pack.py:212-216); `step` routes the chosen block accordingly and
    returns both heightmaps, cat(A, B) along dim 1 (model.py:421-447)."""

    def __init__(self, container_size, blocks_num, reward_type, heightmap_type="full", packing_strategy="LB_GREEDY",
                 batch_size=1, device=None, input_type="mul-with", allow_rot=True):
        if input_type not in ("mul", "mul-with"):
            raise _capi.TapEnvError(_capi.EENUM, "BatchedContainerPairs is for input_type 'mul' / 'mul-with'")
        kw = dict(packing_strategy=packing_strategy, batch_size=batch_size, device=device, input_type=input_type, allow_rot=allow_rot)
        self.a = BatchedContainers(container_size, blocks_num, reward_type, heightmap_type, **kw)
        self.b = BatchedContainers(container_size, blocks_num, reward_type, heightmap_type, **kw)
        self.input_type = input_type
        self.cfg, self.device, self.batch_size, self.S = self.a.cfg, self.a.device, self.a.batch_size, self.a.S
        self.block_dim, self.enc_len = self.a.block_dim, self.a.enc_len
        self.dec_static_rows = self.block_dim + (1 if input_type == "mul-with" else 0)     # model.py:388-394

    def clear_container(self):
        self.a.clear_container(); self.b.clear_container()

    def reset(self, dynamic):
        cur, mask = self.a.reset(dynamic)
        self.b.clear_container()
        return cur, mask

    def _shape(self, t):
        B = self.batch_size
        if self.block_dim == 2:
            return t.view(B, 2 * self.enc_len)                       # [B, 2*enc] (+ unsqueeze(2) at the caller, model.py:424-430)
        W, L = self.a.container_size[0], self.a.container_size[1]
        planes = 2 if self.a.heightmap_type == "diff" else 1
        return t.view(B, 2 * planes, W, L)                           # model.py:431-441

    def step(self, ptr, static, dynamic, mask, out=None):
        """-> (dynamic', current_mask, mask', decoder_static [B,dec_static_rows], decoder_dynamic cat(A,B))."""
        ptr = _dev(ptr, "ptr", torch.int64)
        static = _dev(static, "static", torch.float32)
        dynamic = _dev(dynamic, "dynamic", torch.float32)
        mask = _dev(mask, "mask", torch.float32)
        B, S = self.batch_size, self.S
        if tuple(dynamic.shape) != (B, self.cfg.dyn_rows, S) or tuple(static.shape) != (B, self.cfg.static_rows, S) \
                or tuple(mask.shape) != (B, S) or tuple(ptr.shape) != (B,):
            raise _capi.TapEnvError(_capi.ESHAPE, "step tensors")
        if out is None:
            f32 = dict(dtype=torch.float32, device=self.device)
            dyn_out, cur, mask_out = torch.empty_like(dynamic), torch.empty_like(mask), torch.empty_like(mask)
            dec_static = torch.empty(B, self.dec_static_rows, **f32)
            dec_dyn = torch.empty(B, 2, self.enc_len, **f32)
        else:
            dyn_out, cur, mask_out, dec_static, dec_dyn = out
        self.a._version += 1; self.b._version += 1
        with torch.cuda.device(self.device):
            _capi.check(_capi.lib.tapenv_step_mul(C.byref(self.cfg), _p(self.a.state), _p(self.b.state), _p(ptr), _p(static),
                                                  _p(dynamic), _p(mask), _p(dyn_out), _p(cur), _p(mask_out), _p(dec_static),
                                                  self.dec_static_rows, _p(dec_dyn), _stream()), "step_mul")
        return dyn_out, cur, mask_out, dec_static, self._shape(dec_dyn)

    def add_new_blocks(self, blocks, target_ids, raw=False):
        """model.py:421-428 for the batch: blocks f32 [B,dim], target_ids [B] (0 -> A, 1 -> B) -> cat(A, B) heightmaps
        (raw=True: f32 [B, 2, enc_len], A's and B's encoding side by side)."""
        blocks = _dev(blocks, "blocks", torch.float32)
        target_ids = _dev(target_ids, "target_ids", torch.float32).reshape(-1)
        if tuple(blocks.shape) != (self.batch_size, self.block_dim) or target_ids.numel() != self.batch_size:
            raise _capi.TapEnvError(_capi.ESHAPE, "blocks / target_ids")
        out = torch.empty(self.batch_size, 2, self.enc_len, dtype=torch.float32, device=self.device)
        self.a._version += 1; self.b._version += 1
        with torch.cuda.device(self.device):
            _capi.check(_capi.lib.tapenv_add_blocks_mul(C.byref(self.cfg), _p(self.a.state), _p(self.b.state), _p(blocks),
                                                        _p(target_ids), _p(out), _stream()), "add_blocks_mul")
        return out if raw else self._shape(out)

    def calc_ratio(self):
        """(calc_ratio(A) + calc_ratio(B)) / 2 in fp32 (model.py:503-507) -> f32 [B]."""
        r = torch.empty(self.batch_size, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _capi.check(_capi.lib.tapenv_reward_mul(C.byref(self.cfg), _p(self.a.state), _p(self.b.state), _p(r), _stream()), "reward_mul")
        return r

    def check_flags(self):
        self.a.check_flags(); self.b.check_flags()


class _Group(object):
    """Containers constructed back to back with identical arguments (model.py:294's list comprehension)."""

    def __init__(self, args):
        self.args = args
        self.members = []
        self.batch = None        # None (undecided) | "single" | BatchedContainers
        self.closed = False


class _PairProxy(object):
    """Two container lists driven by an UNMODIFIED model.py for input_type 'mul' / 'mul-with' (model.py:291-292, :414-428):
    per decode step and environment b EITHER `containers_a[b].add_new_block(blocks[b]); containers_b[b].get_heightmap()` OR the
    other way round -- which list receives the block is only revealed call by call.  The return values are merely collected
    in lists and converted after the loop (`torch.FloatTensor(heightmaps_a)`, :430-447), so add_new_block hands out row VIEWS
    of a host buffer that is filled when the step's last row has arrived: then ONE launch (tapenv_add_blocks_mul with the
    targets inferred from which list was called) and ONE D2H copy serve the whole batch.  get_heightmap needs no launch: a
    container that is only looked at keeps the encoding the previous step returned for it."""

    def __init__(self, pairs):
        self.pairs = pairs
        self.B = pairs.batch_size
        self.step = None          # the decode step being collected: src array, blocks, targets, next row, the two host buffers
        self.last = None          # ((enc_a, enc_b) of the last finished step, (version_a, version_b) they describe)
        a = pairs.a
        if a.block_dim == 3:
            W, L = a.container_size[0], a.container_size[1]
            self.shape = (self.B, 2, W, L) if a.heightmap_type == "diff" else (self.B, W, L)
        else:
            self.shape = (self.B, a.enc_len)
        # freshly cleared containers: every encoding of an empty heightmap is zeros
        self.last = ((np.zeros(self.shape, np.int64), np.zeros(self.shape, np.int64)), (pairs.a._version, pairs.b._version))

    def pending(self):
        return self.step is not None

    def add(self, c, block):
        st = self.step
        base = getattr(block, "base", None)
        dim = self.pairs.block_dim
        if st is None:
            if c._row != 0 or base is None or base.ndim != 2 or base.shape[0] != self.B or base.shape[1] < dim:
                raise RuntimeError("tapenv: with two container lists, add_new_block must be called for rows 0..B-1 in order with "
                                   "rows of one [B,dim] array (model.py:412-428); use BatchedContainerPairs.add_new_blocks")
            st = self.step = {"src": base, "blocks": np.ascontiguousarray(base[:, :dim], dtype=np.float32),
                              "targets": np.zeros(self.B, np.float32), "next": 0,
                              # (until the step is complete the rows hold a glaring sentinel, not plausible zeros)
                              "enc": (np.full(self.shape, -2 ** 62, np.int64), np.full(self.shape, -2 ** 62, np.int64))}
        b = c._row
        if base is not st["src"] or b != st["next"]:
            self.step = None
            raise RuntimeError("tapenv: add_new_block rows must come, in order, from the batch handed to row 0")
        st["targets"][b] = c._side
        st["next"] = b + 1
        out = st["enc"][c._side][b]
        if b + 1 == self.B:
            self._finish()
        return out

    def _finish(self):
        st, pr = self.step, self.pairs
        self.step = None
        dev = pr.device
        raw = pr.add_new_blocks(torch.from_numpy(st["blocks"]).to(dev), torch.from_numpy(st["targets"]).to(dev), raw=True)
        host = raw.cpu().numpy().astype(np.int64)                        # ONE launch + ONE D2H per step
        st["enc"][0][...] = host[:, 0].reshape(self.shape)               # fills the views handed out during the step
        st["enc"][1][...] = host[:, 1].reshape(self.shape)
        self.last = (st["enc"], (pr.a._version, pr.b._version))

    def get(self, c):
        pr = self.pairs
        last = self.last
        if last is not None and last[1] == (pr.a._version, pr.b._version):
            return last[0][c._side][c._row]
        return None


_tls = threading.local()
_ctor_cache = {}


class Container(object):
    """tools.Container (tools.py:3607) signature, usable exactly like the reference class.

    Drop-in behaviour for an UNMODIFIED model.py: `[tools.Container(...) for _ in range(B)]` (model.py:294)
    creates B of these back to back; they only remember their construction order.  The first
    `add_new_block` (model.py:452-453 calls it once per environment with `blocks[b]`, a row VIEW of one [B,dim]
    ndarray) reveals the batch: row 0's call binds all B objects to ONE BatchedContainers, launches the step
    for the whole batch from `block.base`, copies the encoded heightmaps to the host once, and every later
    row call just returns its row.  Used any other way (a single object, rows out of order, blocks that are
    not views of one array) each object becomes its own batch of one -- correct, just not batched."""

    def __init__(self, container_size, blocks_num, reward_type, heightmap_type="full", initial_container_size=None,
                 max_height=None, packing_strategy="LB_GREEDY", _batch=None, _row=0):
        self._batch = _batch
        self._row = _row
        self._group = None
        self._pair = None          # set when this object is one of the 2B containers of a 'mul' / 'mul-with' model (_PairProxy)
        self._side = 0
        self.initial_container_size = initial_container_size
        if _batch is not None:
            self._adopt(_batch)
            return
        # model.py:294 constructs B of these back to back with identical arguments: validate the strings once per distinct
        # argument tuple (4 096 constructions per DRL.forward at the C2 batch -- r02: 5.5 -> 1.5 us each)
        try:
            key = (tuple(container_size), blocks_num, reward_type, heightmap_type, packing_strategy)
            proto = _ctor_cache.get(key)
        except TypeError:                                # unhashable arguments (an ndarray size ...): the slow path below
            key, proto = None, None
        if proto is None:
            R = rotate_types(len(container_size), True)
            n_eff = int(blocks_num) if int(blocks_num) * R <= _capi.limits().max_candidates else 1
            cfg = make_config(1, n_eff, container_size, reward_type, heightmap_type, packing_strategy,
                              capacity=int(blocks_num))                                             # validates the strings
            size = [int(v) for v in container_size]
            proto = (size, len(size), int(blocks_num), "MACS" if cfg.strategy == _capi.MACS else packing_strategy,
                     (tuple(size), int(blocks_num), reward_type, heightmap_type, packing_strategy))
            if key is not None:
                if len(_ctor_cache) > 256:
                    _ctor_cache.clear()
                _ctor_cache[key] = proto
        self.container_size = list(proto[0])
        self.block_dim = proto[1]
        self.blocks_num = proto[2]
        self.reward_type = reward_type
        self.heightmap_type = heightmap_type
        self.packing_strategy = proto[3]
        args = proto[4]
        g = getattr(_tls, "open_group", None)
        if g is None or g.closed or g.args != args:
            g = _Group(args)
            _tls.open_group = g
        self._group = g
        self._row = len(g.members)
        g.members.append(self)

    def _adopt(self, batch):
        self.container_size = batch.container_size
        self.block_dim = batch.block_dim
        self.blocks_num = batch.blocks_num
        self.reward_type = batch.reward_type
        self.heightmap_type = batch.heightmap_type
        self.packing_strategy = batch.packing_strategy

    def _bind(self, batch_hint=None):
        if self._batch is not None:
            if self._pair is not None and self._pair.pending():
                raise RuntimeError("tapenv: a decode step of the two container lists is still being collected "
                                   "(add_new_block has not been called for every row yet)")
            return
        g = self._group
        g.closed = True
        size, n, rt, hm, strat = g.args
        if g.batch is None:
            if batch_hint is not None and batch_hint > 1 and batch_hint == len(g.members) and self._row == 0:
                g.batch = BatchedContainers(size, n, rt, hm, packing_strategy=strat, batch_size=batch_hint)
                for m in g.members:                      # rows 1..B-1 of this very step already take the O(1) path
                    m._batch = g.batch
                return
            if batch_hint is not None and 2 * batch_hint == len(g.members) and self._row in (0, batch_hint):
                # two lists of B containers built back to back (model.py:291-292) and row 0's block: 'mul' / 'mul-with'
                pairs = BatchedContainerPairs(size, n, rt, hm, packing_strategy=strat, batch_size=batch_hint, input_type="mul-with")
                g.batch = pairs
                proxy = _PairProxy(pairs)
                for i, m in enumerate(g.members):
                    m._side, m._row = divmod(i, batch_hint)
                    m._batch = pairs.b if m._side else pairs.a
                    m._pair = proxy
                return
            else:
                g.batch = "single"
        if isinstance(g.batch, str):                     # "single"
            self._batch = BatchedContainers(size, n, rt, hm, packing_strategy=strat, batch_size=1)
            self._row = 0
        else:
            self._batch = g.batch

    @classmethod
    def _view_of(cls, batch, row):
        views = batch.__dict__.setdefault("_row_views", {})
        v = views.get(row)
        if v is None:
            v = cls(None, None, None, _batch=batch, _row=row)
            views[row] = v
        return v

    def add_new_block(self, block, is_rotate=False):
        if self._pair is not None:
            return self._pair.add(self, block)
        bt = self._batch
        if bt is not None:
            # rows 1..B-1 of a decode step (model.py:452-453): the batch was launched by row 0; O(1) per call -- the row is
            # recognised by being a view of the very [B,dim] array row 0 handed in (model.py:412), no data is touched
            pending = bt.__dict__.get("_pending")
            if pending is not None and pending["next"] == self._row and getattr(block, "base", None) is pending["src"]:
                b = self._row
                pending["next"] = b + 1
                if b + 1 == bt.batch_size:
                    bt.__dict__["_pending"] = None
                return pending["enc"][b]
        block = np.asarray(block, dtype=np.float32)
        base = block.base
        if self._batch is None:
            # ('mul-with' slices the block out of a [B, dim+1] array, model.py:409-410: the base keeps the id column)
            hint = base.shape[0] if (base is not None and base.ndim == 2 and base.shape[1] in (self.block_dim, self.block_dim + 1)) else None
            if hint is not None and base.shape[1] != self.block_dim and 2 * hint != len(self._group.members):
                hint = None
            self._bind(hint)
            if self._pair is not None:
                return self._pair.add(self, block)
        bt, b = self._batch, self._row
        pending = bt.__dict__.get("_pending")
        if pending is not None and pending["next"] == b and b > 0:   # same protocol, rows that are copies: compare the values
            pending["next"] = b + 1
            if not np.array_equal(pending["blocks"][b], block):
                raise RuntimeError("tapenv: add_new_block rows must come from the batch handed to row 0")
            if pending["next"] == bt.batch_size:
                bt.__dict__["_pending"] = None
            return pending["enc"][b]
        if bt.batch_size > 1:
            if b != 0 or base is None or base.shape != (bt.batch_size, bt.block_dim):
                raise RuntimeError("tapenv: in a batch, add_new_block must be called for rows 0..B-1 in order with "
                                   "rows of one [B,dim] array (model.py:412,452-453); use BatchedContainers.add_new_blocks")
            blocks = np.ascontiguousarray(base, dtype=np.float32)
        else:
            blocks = block.reshape(1, -1)
        enc = bt.add_new_blocks(torch.from_numpy(blocks).to(bt.device)).cpu().numpy().astype(np.int64)   # ONE launch + ONE D2H per step
        if bt.batch_size > 1:
            bt.__dict__["_pending"] = {"next": 1, "blocks": blocks, "src": base, "enc": enc}
        return enc[b]

    def get_heightmap(self, is_full=None):
        if is_full is None:
            if self._batch is None:
                # a freshly constructed container that nobody has touched: every encoding of an empty heightmap is zeros.  Not
                # binding here matters: with two container lists the first call of a decode step may be containers_a[0]'s
                # get_heightmap() (model.py:425-426), and the batch is only revealed by the add_new_block that follows.
                W = [int(v) for v in self.container_size[:-1]]
                if self.heightmap_type == "diff":
                    return np.zeros(W[0] - 1 if self.block_dim == 2 else [2] + W, dtype=np.int64)
                return np.zeros(W, dtype=np.int64)
            if self._pair is not None:
                enc = self._pair.get(self)
                if enc is not None:
                    return enc
        h = self.heightmap
        if is_full is not None or self.heightmap_type == "full":
            return h
        if self.heightmap_type == "zero":
            return h - h.min()
        if self.block_dim == 2:
            return h[1:] - h[:-1]
        out = np.zeros((2,) + h.shape, dtype=h.dtype)
        out[0, 1:, :] = h[1:, :] - h[:-1, :]
        out[1, :, 1:] = h[:, 1:] - h[:, :-1]
        return out

    def clear_container(self):
        self._bind()
        if self._batch.batch_size != 1:
            raise RuntimeError("tapenv: clear the whole batch with BatchedContainers.clear_container()")
        self._batch.clear_container()

    def calc_ratio(self):
        """One launch + one D2H copy per batch state (cached until the batch changes), then row reads:
        model.py:509-510 calls this once per environment."""
        self._bind()
        bt = self._batch
        cache = bt.__dict__.get("_ratio_cache")
        if cache is None or cache[0] != bt._version:
            cache = (bt._version, bt.calc_ratio().cpu().numpy())
            bt.__dict__["_ratio_cache"] = cache
            # once per batch state: conditions under which the reference's NumPy raises IndexError (a stack above the
            # container height, more blocks than blocks_num) must not pass silently through an unmodified model.py
            bt.check_flags()
        return float(cache[1][self._row])

    def calc_CPS(self):
        self._bind()
        v, e, s, k = [int(x) for x in self._batch.scalars[self._row].tolist()]
        if k == 0:
            return 0, 0, 0
        h = int(self.heightmap.max())
        cells = self._batch._cells
        return v / (cells * h), v / (e + v), s / k

    # attributes the reference's callers read (rolling.py:640-658, model.py:1175)
    @property
    def container(self):
        """The reference's voxel grid (tools.py:3629: 0 empty / -1 empty under a block / k+1 block id), int [W(,L),H]: read
        from the state for the strategies that keep it (LB, MACS 3D), rebuilt from the recorded placements otherwise."""
        from . import episode
        self._bind()
        bt = self._batch
        if bt.cfg.strategy == _capi.LB or (bt.cfg.strategy == _capi.MACS and self.block_dim == 3):
            cells, H = bt._cells, int(self.container_size[-1])
            v = bt._view(bt._layout.voxels, bt.batch_size * cells * H, torch.int16, (bt.batch_size, cells, H))[self._row]
            return v.cpu().numpy().astype(np.int64).reshape([int(x) for x in self.container_size])
        k = self.current_blocks_num
        return episode.voxel_container(self.positions[:k], bt.blocks[self._row].cpu().numpy()[:k], self.container_size)

    @property
    def blocks(self):
        """The blocks added so far (tools.py:3631: a list, one entry per add_new_block call)."""
        self._bind()
        k = self.current_blocks_num
        return [b for b in self._batch.blocks[self._row].cpu().numpy()[:k].astype(np.int64)]

    @property
    def rotate_state(self):
        """tools.py:3634 -- the flags model.py passes to add_new_block are only drawn, never used by the packing: not kept."""
        return [False] * self.blocks_num

    @property
    def max_height(self):
        return 2 * int(self.container_size[0])           # tools.py:3624-3627 (never read by the packing functions)

    @property
    def bounding_box(self):
        """tools.py:3633: zeros -- add_new_block never stores the bounding box the placement functions return (:3708)."""
        return np.zeros(self.block_dim)

    def draw_container(self, save_name, order=None):
        """Container.draw_container (tools.py:3968-3996).  Visualisation stays the reference's: this container's blocks and
        positions (2D) or its rebuilt voxel grid (3D) are handed to tools.draw_container_2d / tools.draw_container_voxel of
        the module tapenv.install() patched.  (The rotate flags model.py passes to add_new_block are not kept: zeros.)"""
        from . import dropin, episode
        tools = dropin.installed_tools()
        if tools is None:
            raise RuntimeError("tapenv: draw_container needs the reference's drawing functions (tapenv.install(pack, tools))")
        if self.initial_container_size is None:
            print('Do not know the initial_conainer_size')
            return
        self._bind()
        k = self.current_blocks_num
        blocks = self._batch.blocks[self._row].cpu().numpy().astype(int)[:k]
        positions = self.positions
        if order is None:
            order = [i for i in range(k)]
        C, P, S = self.calc_CPS()
        title = "Compactness: %.3f\nPyramidality: %.3f\nStability: %.3f\n" % (C, P, S)
        rotate_state = np.zeros(self.blocks_num, dtype=bool)
        if self.block_dim == 2:
            tools.draw_container_2d(blocks, positions.astype('int'), self.container_size, self.reward_type, order, self.stable, None,
                                    rotate_state, save_title=title, save_name=save_name)
        else:
            grid = episode.voxel_container(positions[:k], blocks, self.container_size)
            tools.draw_container_voxel(grid, k, reward_type=self.reward_type, order=order, rotate_state=rotate_state,
                                       feasibility=None, save_title=title, save_name=save_name)

    @property
    def heightmap(self):
        self._bind()
        return self._batch.heightmap[self._row].cpu().numpy().astype(np.int64)

    @property
    def positions(self):
        self._bind()
        return self._batch.positions[self._row].cpu().numpy().astype(np.int64)

    @property
    def stable(self):
        self._bind()
        return [bool(v) for v in self._batch.stable[self._row].tolist()]

    @property
    def valid_size(self):
        self._bind()
        return int(self._batch.scalars[self._row, 0].item())

    @property
    def empty_size(self):
        self._bind()
        return int(self._batch.scalars[self._row, 1].item())

    @property
    def current_blocks_num(self):
        self._bind()
        return int(self._batch.scalars[self._row, 3].item())
