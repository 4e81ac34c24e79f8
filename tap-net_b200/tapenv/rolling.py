"""Rolling-window packing (rolling.py:575-640) for a batch of instances on one GPU.

`BatchedInitialContainers` is generate.InitialContainer (generate.py:1589-1825) for B instances at once: the five
precedence graphs live in HBM as predecessor bit masks, the mutable part (gm's node set, after_nodes_list,
sub_graph_nodes) is 64 bytes per instance, and `convert_to_input` is one kernel launch for the whole batch instead of
five networkx sub-graph copies per instance and step.  `RollingRunner` is rolling.validate's per-instance loop with
the batch axis restored: ONE launch per decode step (tapenv_rolling_step: place the chosen block, drop it from the
window, refill, emit the next window's static / dynamic / masks).
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _capi
from .containers import BatchedContainers
from .ops import _dev, _p, _stream


# node_order=REFERENCE reproduces an accident of the reference's dependencies, not a property of the algorithm: networkx's
# subgraph view iterates a Python set when 2*window < total (coreviews.FilterAtlas.__iter__), i.e. CPython's open-addressing
# layout (Objects/setobject.c: 8 -> 32 -> 128 slot tables, LINEAR_PROBES 9, PERTURB_SHIFT 5).  Derived from and pinned
# against these versions; a different interpreter / networkx may enumerate differently (then use WINDOW_ORDER_SORTED, or
# re-derive window.cuh:pyset_order_warp).
REFERENCE_ORDER_DERIVED_FROM = {"cpython": (3, 12), "networkx": "3.6"}


def reference_order_matches_this_interpreter():
    """Does THIS interpreter's set() iterate the way the kernel's restatement assumes?  Two closed-form regimes are probed:
    6 keys without a collision modulo 32 sit in slot key & 31 of the 32-slot table; 20 keys have grown the table to 128 slots
    and iterate in ascending order."""
    keys = [2, 5, 7, 11, 20, 44]
    big = list(range(0, 60, 3))
    return list(set(keys)) == sorted(keys, key=lambda k: k & 31) and list(set(big)) == big


def _warn_if_order_unpinned(node_order):
    import sys
    import warnings
    if int(node_order) != _capi.WINDOW_ORDER_REFERENCE:
        return
    if sys.version_info[:2] != REFERENCE_ORDER_DERIVED_FROM["cpython"] or not reference_order_matches_this_interpreter():
        warnings.warn("tapenv: node_order=REFERENCE reproduces CPython %d.%d / networkx %s set iteration order; this interpreter "
                      "differs, so a live reference may permute `dynamic` differently (use WINDOW_ORDER_SORTED)"
                      % (REFERENCE_ORDER_DERIVED_FROM["cpython"] + (REFERENCE_ORDER_DERIVED_FROM["networkx"],)))


def rotation_structured(blocks, total_blocks, dim):
    """True iff blocks[..., r*T+i, :] == blocks[..., i, perm_r] for every rotation r of itertools.permutations(range(dim))
    -- the layout generate.generate_blocks writes (numpy array [B,R*T,dim] or [R*T,dim])."""
    import itertools
    blocks = np.asarray(blocks)
    T = int(total_blocks)
    base = blocks[..., :T, :]
    return all(np.array_equal(blocks[..., r * T:(r + 1) * T, :], base[..., list(p)])
               for r, p in enumerate(itertools.permutations(range(dim))))


def pack_graphs(adj):
    """adj [B,5,T,T] 0/1 with adj[b,g,u,v] = edge u -> v (deps_g[u,v] == True, generate.py:1636-1664)
    -> predecessor masks int64 [B,5,T]: bit u of pred[b,g,v]."""
    adj = np.asarray(adj)
    B, G, T, _ = adj.shape
    assert G == 5 and T <= 64
    w = (np.uint64(1) << np.arange(T, dtype=np.uint64))[None, None, :, None]
    pred = (adj.astype(np.uint64) * w).sum(axis=2, dtype=np.uint64)
    return pred.view(np.int64)


class BatchedInitialContainers(object):
    """B generate.InitialContainer objects.  graphs: int64 [B,5,T] predecessor masks (pack_graphs) or the [B,5,T,T]
    adjacency; blocks: [B,R*T,dim] rotation-major block sizes (rolling.py:483-485)."""

    def __init__(self, graphs, blocks, blocks_num, child_graph_size, block_dim, device=None,
                 node_order=_capi.WINDOW_ORDER_REFERENCE, input_type="bot", blocks_are_rotations=None):
        if not torch.cuda.is_available():
            raise RuntimeError("tapenv: a CUDA device is required (no CPU fallback exists)")
        if input_type != "bot":
            raise _capi.TapEnvError(_capi.EUNSUPPORTED, "InitialContainer.convert_to_input only works for 'bot' inputs (generate.py:1790-1806)")
        _warn_if_order_unpinned(node_order)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        graphs = np.asarray(graphs) if not isinstance(graphs, torch.Tensor) else graphs
        if not isinstance(graphs, torch.Tensor):
            if graphs.ndim == 4:
                graphs = pack_graphs(graphs)
            graphs = torch.from_numpy(np.ascontiguousarray(graphs, dtype=np.int64))
        if not isinstance(blocks, torch.Tensor):
            blocks = torch.from_numpy(np.ascontiguousarray(blocks, dtype=np.int32))
        self.blocks_num, self.child_graph_size, self.block_dim = int(blocks_num), int(child_graph_size), int(block_dim)
        self.rotate_types = math.factorial(self.block_dim)
        B = int(graphs.shape[0])
        T, n, R, dim = self.blocks_num, self.child_graph_size, self.rotate_types, self.block_dim
        if tuple(graphs.shape) != (B, 5, T) or tuple(blocks.shape) != (B, R * T, dim):
            raise _capi.TapEnvError(_capi.ESHAPE, "graphs %s / blocks %s" % (tuple(graphs.shape), tuple(blocks.shape)))
        self.batch_size = B
        self.S = n * R
        self.graphs = graphs.to(self.device, torch.int64).contiguous()
        self.blocks = blocks.to(self.device, torch.int32).contiguous()
        # blocks_are_rotations: None = check `blocks` now; pass False/True explicitly when the buffer will be overwritten
        # later (host pipelines) -- False is always correct, True promises the rotation layout of generate_blocks
        rot = self._rotation_structured() if blocks_are_rotations is None else bool(blocks_are_rotations)
        self.wcfg = _capi.WindowConfig(B, T, n, dim, R, int(node_order), int(rot))
        nbytes = int(_capi.lib.tapenv_window_state_bytes(C.byref(self.wcfg)))
        if B > 0 and nbytes == 0:
            raise _capi.TapEnvError(_capi.ELIMIT, "window config (total_blocks <= 64, window <= 32, window*R <= 64)")
        self.state = torch.empty(max(nbytes, 4), dtype=torch.uint8, device=self.device)
        self.sub_graph_nodes = torch.full((B, n), -1, dtype=torch.int32, device=self.device)   # sorted, as generate.py:1766 leaves it
        self.remaining = torch.full((B,), T, dtype=torch.int32, device=self.device)             # len(after_nodes_list)
        self._pending = None
        self.reset()

    def _rotation_structured(self):
        """True iff blocks[:, r*T+i, :] == blocks[:, i, perm_r] for every rotation (how generate.generate_blocks lays the
        rotations out); checked on the device once, lets the kernel read only the un-rotated rows."""
        import itertools
        T, dim = self.blocks_num, self.block_dim
        if self.blocks.numel() == 0:
            return True
        base = self.blocks[:, :T, :]
        ok = True
        for r, p in enumerate(itertools.permutations(range(dim))):
            ok = ok and bool(torch.equal(self.blocks[:, r * T:(r + 1) * T, :], base[:, :, list(p)]))
        return ok

    def refresh_blocks(self):
        """Call after overwriting `self.blocks` in place (e.g. a host pipeline re-using the buffers)."""
        self.wcfg.blocks_are_rotations = int(self._rotation_structured())

    def reset(self):
        """Back to the freshly constructed InitialContainer (generate.py:1666-1673)."""
        with torch.cuda.device(self.device):
            _capi.check(_capi.lib.tapenv_window_reset(C.byref(self.wcfg), _p(self.state), _stream()), "window_reset")
        self._pending = None

    # ---- the reference's three calls ------------------------------------------------------------------
    def remove_block(self, ptr):
        """rolling.py:636-640 for the whole batch: drop sub_graph_nodes[ptr mod n] of every instance.  Like the
        reference ("the graph will update when you call convert_to_input", generate.py:1811) the removal is applied by
        the next convert_to_input launch."""
        self._pending = _dev(ptr, "ptr", torch.int64)

    def convert_to_input(self, out=None, masks=None):
        """-> (static f32 [B,1+dim,S], dynamic f32 [B,3n,S]) (generate.py:1770-1808).  masks=(cur, mask) optionally
        receives the initial masks rolling.DRL.forward derives from `dynamic` (rolling.py:325-335)."""
        B, n, S, dim = self.batch_size, self.child_graph_size, self.S, self.block_dim
        if out is None:
            static = torch.empty(B, 1 + dim, S, dtype=torch.float32, device=self.device)
            dynamic = torch.empty(B, 3 * n, S, dtype=torch.float32, device=self.device)
        else:
            static, dynamic = out
        cur, mask = masks if masks is not None else (None, None)
        with torch.cuda.device(self.device):
            _capi.check(_capi.lib.tapenv_window_next(C.byref(self.wcfg), _p(self.state), _p(self.graphs), _p(self.blocks),
                                                     _p(self._pending), _p(static), _p(dynamic), _p(cur), _p(mask),
                                                     _p(self.sub_graph_nodes), _p(self.remaining), _stream()), "window_next")
        self._pending = None
        return static, dynamic

    def is_last_graph(self):
        """generate.py:1824-1825 per instance -> bool [B] (device).  Every instance admits exactly one node per call
        after the first, so the flags agree across a batch of equally sized instances."""
        return self.remaining == 0

    # ---- state views ------------------------------------------------------------------------------------
    @property
    def flags(self):
        """int32 [B] sticky: 1 no in-degree-0 node (the reference would spin), 2 window not full at convert (the reference
        raises), 4 pointer outside the window."""
        return self.state.view(torch.int32).view(self.batch_size, 16)[:, 13]

    def check_flags(self):
        bad = int((self.flags != 0).sum().item())
        if bad:
            raise IndexError("tapenv: %d rolling window(s) in an invalid state (flags %s)" % (bad, sorted(set(self.flags.tolist()) - {0})))


class RollingRunner(object):
    """rolling.validate's loop (rolling.py:589-640) + the env section of rolling.DRL.forward for B instances, the
    network replaced by a pointer source.  Per instance: T - n + 1 windows; the first T - n are decoded for ONE step
    (one_step=True), the last completely.  `step(ptr)` is one launch while windows remain, the fused decode step
    (tapenv_step) inside the last window."""

    def __init__(self, env, windows, ptr_seq=None, use_graph=False, partial_sums=True, exchange=None, overlap_exchange=None):
        from .runner import RewardTail
        assert isinstance(env, BatchedContainers) and isinstance(windows, BatchedInitialContainers)
        assert env.batch_size == windows.batch_size and env.window == windows.child_graph_size
        assert env.blocks_num >= windows.blocks_num, "the container must hold total_blocks_num blocks (rolling.py:702-703)"
        self.env, self.win = env, windows
        B, n, S, dim = env.batch_size, windows.child_graph_size, windows.S, windows.block_dim
        dev = env.device
        f32 = dict(dtype=torch.float32, device=dev)
        self.static = [torch.empty(B, 1 + dim, S, **f32) for _ in range(2)]
        self.dynamic = [torch.empty(B, 3 * n, S, **f32) for _ in range(2)]
        self.cur = [torch.empty(B, S, **f32) for _ in range(2)]
        self.mask = [torch.empty(B, S, **f32) for _ in range(2)]
        self.dec_static = torch.empty(B, dim, **f32)
        self.dec_dyn = torch.empty(B, env.enc_len, **f32)
        self.total = windows.blocks_num
        self.t = 0
        self.slot = 0                                 # ping-pong index of dynamic / masks
        self.sslot = 0                                # ... of static (unchanged inside the last window)
        # whole-episode replay of a recorded pointer sequence int64 [T,B] (EpisodeRunner's interface)
        self.ptr_seq = ptr_seq
        self.partial_sums = partial_sums
        self.exchange = exchange
        self.tail = RewardTail(env, partial_sums, exchange, overlap_exchange)
        self.reward = None
        self.reward_buf = None
        # clear + window reset + first window, T steps (the last also emits the rewards), the sums (+ exchange) launch
        self.launches_per_episode = 3 + self.total + self.tail.launches
        self.graph = None
        if use_graph:
            assert ptr_seq is not None
            self._capture()

    @property
    def steps_total(self):
        return self.total

    def begin(self):
        """clear_container (rolling.py:659) + a fresh InitialContainer + the first window -> (static, dynamic, cur_mask)."""
        self.env.clear_container()
        self.win.reset()
        self.t, self.slot, self.sslot = 0, 0, 0
        self.win.convert_to_input(out=(self.static[0], self.dynamic[0]), masks=(self.cur[0], self.mask[0]))
        return self.static[0], self.dynamic[0], self.cur[0]

    def step(self, ptr, reward_out=None):
        """One decode step for every instance; returns (static, dynamic, cur_mask, decoder_static, decoder_dynamic)
        the network sees next.  reward_out (f32 [B]): on the LAST step, also receive calc_ratio()."""
        env, win = self.env, self.win
        ptr = _dev(ptr, "ptr", torch.int64)
        n, T = win.child_graph_size, self.total
        s = self.slot
        if self.t < T - n:                             # one_step window: place + advance the window, ONE launch (every strategy)
            o = s ^ 1
            with torch.cuda.device(env.device):
                _capi.check(_capi.lib.tapenv_rolling_step(
                    C.byref(env.cfg), _p(env.state), C.byref(win.wcfg), _p(win.state), _p(win.graphs), _p(win.blocks),
                    _p(ptr), _p(self.dec_static), _p(self.dec_dyn), _p(self.static[self.sslot ^ 1]), _p(self.dynamic[o]),
                    _p(self.cur[o]), _p(self.mask[o]), _p(win.sub_graph_nodes), _p(win.remaining), _stream()), "rolling_step")
            env._version += 1
            self.slot = o
            self.sslot ^= 1
        else:                                          # last window: the ordinary fused decode step
            o = s ^ 1
            env.step(ptr, self.static[self.sslot], self.dynamic[s], self.mask[s],
                     out=(self.dynamic[o], self.cur[o], self.mask[o], self.dec_static, self.dec_dyn), reward_out=reward_out)
            self.slot = o
        self.t += 1
        o = self.slot
        return self.static[self.sslot], self.dynamic[o], self.cur[o], self.dec_static, env._shape_enc(self.dec_dyn)

    def _episode(self):
        self.begin()
        if self.reward_buf is None:
            self.reward_buf = torch.empty(self.env.batch_size, dtype=torch.float32, device=self.env.device)
        for t in range(self.total):
            self.step(self.ptr_seq[t], reward_out=self.reward_buf if t == self.total - 1 else None)   # calc_ratio rides on the last step
        self.reward = self.reward_buf
        self.tail.inline(self.reward)

    @property
    def sums(self):
        return self.tail.sums

    @property
    def total_sums(self):
        return self.tail.total

    def _capture(self):
        dev = self.env.device
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            self._episode()                           # warm-up outside capture
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._episode()
        self.graph = g

    def run(self, ptr_seq=None):
        """One rolling episode for every instance with the recorded pointers int64 [T,B] (given here, or the
        `ptr_seq` buffer handed to the constructor -- required for CUDA-graph replay); returns calc_ratio f32 [B]."""
        if ptr_seq is not None:
            assert self.graph is None, "a captured runner replays its own ptr_seq buffer"
            self.ptr_seq = ptr_seq
        self.tail.before_episode()
        if self.graph is not None:
            self.graph.replay()
        else:
            self._episode()
        self.tail.after_episode(self.reward)
        return self.reward


class RollingHostPipeline(object):
    """Rolling episodes whose inputs live in HOST memory (what rolling.RollingDataset holds per instance: the
    precedence graphs and the block list): double-buffered H2D upload of (graphs, blocks, ptr_seq) on a copy stream,
    episode replay on the compute stream, D2H of rewards + sums.  Same protocol as runner.HostPipeline."""

    def __init__(self, env, total_blocks, window, depth=2, use_graph=True, exchange=None,
                 node_order=_capi.WINDOW_ORDER_REFERENCE, blocks_are_rotations=False):
        """blocks_are_rotations=True promises that every submitted `blocks` array has generate_blocks' rotation layout
        (rotation_structured(blocks) -- check once per dataset); the default reads all R*T rows."""
        self.env = env
        dev = env.device
        B, dim = env.batch_size, env.block_dim
        R = math.factorial(dim)
        self.depth = depth
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.slots = []
        for _ in range(depth):
            graphs = torch.zeros(B, 5, total_blocks, dtype=torch.int64, device=dev)
            blocks = torch.ones(B, R * total_blocks, dim, dtype=torch.int32, device=dev)
            pq = torch.zeros(total_blocks, B, dtype=torch.int64, device=dev)
            win = BatchedInitialContainers(graphs, blocks, total_blocks, window, dim, device=dev, node_order=node_order,
                                           blocks_are_rotations=blocks_are_rotations)
            runner = RollingRunner(env, win, ptr_seq=pq, use_graph=use_graph, partial_sums=True, exchange=exchange)
            self.slots.append(dict(win=win, ptr=pq, runner=runner, uploaded=torch.cuda.Event(), consumed=torch.cuda.Event(),
                                   done=torch.cuda.Event(), reward=torch.empty(B, dtype=torch.float32).pin_memory(),
                                   sums=torch.empty(3, dtype=torch.float64).pin_memory(), busy=False))
        self.head = self.tail = self.inflight = 0
        self.h2d_bytes = B * 5 * total_blocks * 8 + B * R * total_blocks * dim * 4 + total_blocks * B * 8
        self.d2h_bytes = B * 4 + 24

    def submit(self, graphs_h, blocks_h, ptr_h, after_episode=None):
        if self.inflight == self.depth:
            raise RuntimeError("pipeline full: call result() first")
        s = self.slots[self.head]
        compute = torch.cuda.current_stream(self.env.device)
        with torch.cuda.stream(self.copy_stream):
            if s["busy"]:
                self.copy_stream.wait_event(s["consumed"])
            s["win"].graphs.copy_(graphs_h, non_blocking=True)
            s["win"].blocks.copy_(blocks_h, non_blocking=True)
            s["ptr"].copy_(ptr_h, non_blocking=True)
            s["uploaded"].record(self.copy_stream)
        compute.wait_event(s["uploaded"])
        r = s["runner"].run()
        s["consumed"].record(compute)
        if after_episode is not None:
            after_episode(s["runner"])
        s["reward"].copy_(r, non_blocking=True)
        tail = s["runner"].tail
        if tail.overlap:                                         # the statistics arrive on the side stream
            with torch.cuda.stream(tail.stream):
                s["sums"].copy_(tail.total if tail.total is not None else tail.sums, non_blocking=True)
                tail.reduced.record(tail.stream)
        else:
            s["sums"].copy_(tail.total if tail.total is not None else tail.sums, non_blocking=True)
        s["done"].record(compute)
        s["busy"] = True
        self.head = (self.head + 1) % self.depth
        self.inflight += 1

    def result(self):
        if self.inflight == 0:
            raise RuntimeError("nothing in flight")
        s = self.slots[self.tail]
        s["done"].synchronize()
        tail = s["runner"].tail
        if tail.overlap:
            tail.reduced.synchronize()
        if tail.exchange is not None and s["sums"][0] != s["sums"][0]:      # NaN totals: the exchange timed out on a peer
            tail.exchange.check()
        self.tail = (self.tail + 1) % self.depth
        self.inflight -= 1
        return s["reward"], s["sums"]


# ----------------------------------------------------------------------------------------------------------------
# Dataset side (host, one-off): the precedence graphs of an initial packing and rolling.RollingDataset
# ----------------------------------------------------------------------------------------------------------------
def calc_dependent(blocks, positions, container_size, arm_size=1):
    """The five precedence relations generate.calc_dependent derives from the voxel grid of an initial packing
    (generate.py:575-647 for 2D, :649-752 for 3D), computed from the block intervals instead of by scanning voxels.
    blocks / positions: int [T,dim] (un-rotated sizes, corner positions).  Returns bool [5,T,T] in InitialContainer's
    graph order (move, left, right, forward, backward) with out[g,u,v] == True <=> deps_g[u,v] (edge u -> v).

      move      3D: u rests directly above v in some footprint column -- consecutive blocks of that column's stack,
                whatever the gap (the voxel scan stops at the first block it meets, :676-701).
                2D: u is anywhere above v over a shared column (np.unique over the whole strip, :603-613).
      left/right (and forward/backward in 3D): v needs that side free to be rotated; u -> v if u occupies the
                neighbouring strip (2D: `arm_size` columns; 3D: the one-cell plane, restricted to the block's middle
                cells along the other axis) at or above v's mid height; v -> v when v touches the wall."""
    blocks = np.asarray(blocks, dtype=np.int64)
    pos = np.asarray(positions, dtype=np.int64)
    T, dim = blocks.shape
    lo, hi = pos, pos + blocks                                   # half-open extents per axis
    out = np.zeros((5, T, T), dtype=bool)
    idx = np.arange(T)
    zmid = lo[:, -1] + (blocks[:, -1] - 1) // 2                  # z + int((bz-1)/2)

    def overlap(a_lo, a_hi, b_lo, b_hi):                         # [T,1] x [1,T] half-open interval intersection
        return (a_lo[:, None] < b_hi[None, :]) & (b_lo[None, :] < a_hi[:, None])

    if dim == 2:
        W = int(container_size[0])
        xov = overlap(lo[:, 0], hi[:, 0], lo[:, 0], hi[:, 0])
        below = xov & (lo[None, :, 1] < lo[:, None, 1])          # [u,v]: v has cells under u's bottom edge
        above = xov & (hi[:, None, 1] > hi[None, :, 1])          # [u,v]: u has cells over v's top edge
        out[0] = (below | above) & (idx[:, None] != idx[None, :])
        tall = hi[:, None, 1] > zmid[None, :]                    # [u,v]: u reaches v's mid height
        wall_l = lo[:, 0] < arm_size
        left = overlap(lo[:, 0], hi[:, 0], lo[:, 0] - arm_size, lo[:, 0]) & tall
        out[1] = np.where(wall_l[None, :], np.eye(T, dtype=bool), left)
        wall_r = hi[:, 0] > W - arm_size
        right = overlap(lo[:, 0], hi[:, 0], hi[:, 0], hi[:, 0] + arm_size) & tall
        out[2] = np.where(wall_r[None, :], np.eye(T, dtype=bool), right)
        return out

    W, L = int(container_size[0]), int(container_size[1])
    # move: per footprint column, consecutive blocks of the stack
    order = np.argsort(lo[:, 2], kind="stable")
    top_block = -np.ones((W, L), dtype=np.int64)                 # highest block seen so far per column
    for u in order:
        x0, x1, y0, y1 = lo[u, 0], hi[u, 0], lo[u, 1], hi[u, 1]
        under = np.unique(top_block[x0:x1, y0:y1])
        out[0, u, under[under >= 0]] = True
        top_block[x0:x1, y0:y1] = u
    # rotation: middle cells along the other horizontal axis (y_mid_1:y_mid_2, :705-713), from mid height upwards
    def mid_range(a_lo, size):
        m1 = a_lo + (size - 1) // 2
        m2 = a_lo + size // 2
        return m1, np.where(m1 == m2, m2 + 1, m2)
    ym1, ym2 = mid_range(lo[:, 1], blocks[:, 1])
    xm1, xm2 = mid_range(lo[:, 0], blocks[:, 0])
    tall = hi[:, None, 2] > zmid[None, :]
    eye = np.eye(T, dtype=bool)
    y_mid = overlap(lo[:, 1], hi[:, 1], ym1, ym2)                # [u,v]: u covers one of v's middle y cells
    x_mid = overlap(lo[:, 0], hi[:, 0], xm1, xm2)

    def plane(u_lo, u_hi, cell):                                  # [u,v]: u occupies coordinate cell[v] along that axis
        return (u_lo[:, None] <= cell[None, :]) & (cell[None, :] < u_hi[:, None])
    out[1] = np.where((lo[:, 0] == 0)[None, :], eye, plane(lo[:, 0], hi[:, 0], lo[:, 0] - 1) & y_mid & tall)
    out[2] = np.where((hi[:, 0] == W)[None, :], eye, plane(lo[:, 0], hi[:, 0], hi[:, 0]) & y_mid & tall)
    out[3] = np.where((lo[:, 1] == 0)[None, :], eye, plane(lo[:, 1], hi[:, 1], lo[:, 1] - 1) & x_mid & tall)
    out[4] = np.where((hi[:, 1] == L)[None, :], eye, plane(lo[:, 1], hi[:, 1], hi[:, 1]) & x_mid & tall)
    return out


class RollingDataset(object):
    """rolling.RollingDataset (rolling.py:462-534): reads blocks.txt / pos.txt of a rolling dataset directory and
    builds the initial containers -- here ONE BatchedInitialContainers for all `num_samples` instances (attribute
    `initial_containers`), plus the reference's zero decoder inputs.  dep_*.txt / container.txt are not needed: like
    the reference, the graphs are recomputed from the geometry (generate.py:1619)."""

    def __init__(self, data_file, total_blocks_num, net_blocks_num, num_samples, block_dim, seed, input_type,
                 heightmap_type, allow_rot, container_width, initial_container_width, initial_container_height,
                 mix_data_file=None, unit=1, device=None, node_order=_capi.WINDOW_ORDER_REFERENCE):
        if seed is None:
            seed = np.random.randint(123456)
        np.random.seed(seed)
        torch.manual_seed(seed)
        T, dim = int(total_blocks_num), int(block_dim)
        R = math.factorial(dim)
        blocks = np.atleast_2d(np.loadtxt(data_file + "blocks.txt")).astype("int")
        positions = np.atleast_2d(np.loadtxt(data_file + "pos.txt")).astype("int")
        data_size = len(blocks) // R
        blocks = blocks.reshape(data_size, R, dim, T).transpose(0, 1, 3, 2).reshape(data_size, R * T, dim)   # :483-485
        positions = positions.reshape(len(positions), dim, T).transpose(0, 2, 1)                              # :490-491
        ics = ([initial_container_width, initial_container_height] if dim == 2
               else [initial_container_width, initial_container_width, initial_container_height])
        N = int(num_samples)
        adj = np.stack([calc_dependent(blocks[b, :T], positions[b], ics) for b in range(N)])
        self.blocks, self.positions, self.graphs = blocks[:N], positions[:N], pack_graphs(adj)
        self.initial_containers = BatchedInitialContainers(self.graphs, blocks[:N], T, net_blocks_num, dim, device=device,
                                                           node_order=node_order, input_type=input_type)
        static_dim, hm_num = dim, 1                                                                           # :504-533
        if heightmap_type == "diff":
            hm_w = container_width * unit - 1 if dim == 2 else container_width * unit
            if dim == 3:
                hm_num = 2
        else:
            hm_w = container_width * unit
        self.decoder_static = torch.zeros(1, static_dim, 1, requires_grad=True)
        if dim == 2:
            self.decoder_dynamic = torch.zeros(1, int(hm_w), 1, requires_grad=True)
        else:
            self.decoder_dynamic = torch.zeros(1, hm_num, int(hm_w), int(hm_w), requires_grad=True)
        self.num_samples = N


class InitialContainer(object):
    """generate.InitialContainer with the reference's constructor and call protocol (generate.py:1589-1825), for an
    UNMODIFIED rolling.py (`generate.InitialContainer = tapenv.rolling.InitialContainer`, or tapenv.install(generate=...)):
    one instance per object, NumPy in / NumPy out, the window logic on the GPU (batch of one).  rolling.validate walks
    the instances one by one, so this keeps its structure; the batched classes above are the fast path."""

    def __init__(self, blocks, positions, blocks_num, initial_container_size, allow_bot, child_graph_size,
                 input_type="bot", device=None, node_order=_capi.WINDOW_ORDER_REFERENCE):
        blocks = np.asarray(blocks).astype(np.int64)
        positions = np.asarray(positions).astype(np.int64)
        T, dim = int(blocks_num), len(initial_container_size)
        self.input_type, self.block_dim, self.blocks_num = input_type, dim, T
        self.rotate_types = math.factorial(dim)
        self.blocks, self.positions = blocks, positions
        self.all_bot, self.child_graph_size = allow_bot, int(child_graph_size)
        adj = calc_dependent(blocks[:T], positions, initial_container_size)          # generate.py:1619
        if not allow_bot:
            adj[1:] = False                                                          # rotation graphs stay empty (:1641)
        self.deps = adj
        self._batch = BatchedInitialContainers(adj[None], blocks[None].astype(np.int32), T, child_graph_size, dim,
                                               device=device, node_order=node_order, input_type=input_type)
        self.sub_graph_nodes = []
        self._remaining = T

    def convert_to_input(self):
        static, dynamic = self._batch.convert_to_input()
        nodes = self._batch.sub_graph_nodes[0].cpu().numpy()
        self.sub_graph_nodes = [int(v) for v in nodes if v >= 0]
        self._remaining = int(self._batch.remaining[0].item())
        if len(self.sub_graph_nodes) != self.child_graph_size:
            raise ValueError("all the input array dimensions ... must match exactly (window of %d nodes, %d expected)"
                             % (len(self.sub_graph_nodes), self.child_graph_size))   # np.concatenate, generate.py:1788
        return static[0].cpu().numpy().astype(np.int64), dynamic[0].cpu().numpy().astype(np.float64)

    def remove_block(self, block_id):
        try:
            idx = self.sub_graph_nodes.index(int(block_id))
        except ValueError:
            return                                                                   # generate.py:1819-1821: silently ignored
        self._batch.remove_block(torch.tensor([idx], dtype=torch.int64, device=self._batch.device))
        self.sub_graph_nodes.remove(int(block_id))

    def is_last_graph(self):
        return self._remaining == 0
