"""Replay of whole packing episodes through the per-step C-ABI calls, with preallocated ping-pong
buffers and optional CUDA-graph capture of the n-step launch sequence (reset + n fused steps + reward).

This is the host-side loop model.py:294-515 runs around the environment, minus the network: the
pointer for every step comes from a `ptr_seq` tensor (a recorded policy) instead of the actor.
"""
import torch

from .containers import BatchedContainers


_side_streams = {}


def _side_stream(device):
    """One high-priority side stream per device for end-of-episode statistics without an exchange."""
    key = torch.device(device).index
    st = _side_streams.get(key)
    if st is None:
        lo, hi = torch.cuda.Stream.priority_range()
        st = _side_streams[key] = torch.cuda.Stream(device=device, priority=hi)
    return st


class RewardTail(object):
    """The end of an episode: the deterministic (sum r, sum r^2, B) of the reward vector and, on several GPUs, the exchange
    of those triples (tapenv_reward_sums / PeerExchange) -- the operand of the critic-baseline statistics
    (trainer.py:216-225).

    overlap=False: one launch at the tail of the episode (captured in its CUDA graph).
    overlap=True (default): the launch goes to a high-priority side stream (the exchange's, if there is one) right behind
    the last decode step; the next episode's reset does NOT wait for the reduction or for the poll on the slowest rank
    (r01: the in-stream poll cost ~4 us per 62 us episode at 8 GPUs; r02k at 2 GPUs: 60.2 -> 55.0 us per episode).
    `sums` / `total` are then valid after `reduced` (wait_total())."""

    def __init__(self, env, partial_sums=True, exchange=None, overlap=None):
        self.env, self.exchange = env, exchange
        self.enabled = bool(partial_sums) or exchange is not None
        self.overlap = self.enabled and (True if overlap is None else bool(overlap))
        dev = env.device
        self.stream = exchange.stream if exchange is not None else (_side_stream(dev) if self.overlap else None)
        self.sums = torch.zeros(3, dtype=torch.float64, device=dev) if self.enabled else None
        self.total = torch.zeros(3, dtype=torch.float64, device=dev) if exchange is not None else None
        self.posted = torch.cuda.Event()
        self.reduced = torch.cuda.Event()
        self.pending = False

    @property
    def launches(self):
        return 1 if self.enabled else 0

    def inline(self, reward):
        """Inside the episode's launch sequence (graph-capturable)."""
        if self.enabled and not self.overlap:
            self.env.reward_sums(reward, exchange=self.exchange, out=(self.sums, self.total))

    def before_episode(self):
        """The previous use of this runner's reward / sums buffers must have been consumed by its exchange."""
        if self.overlap and self.pending:
            torch.cuda.current_stream(self.env.device).wait_event(self.reduced)
            self.pending = False

    def after_episode(self, reward):
        if not (self.enabled and self.overlap):
            return
        cur = torch.cuda.current_stream(self.env.device)
        xs = self.stream
        self.posted.record(cur)
        xs.wait_event(self.posted)
        with torch.cuda.stream(xs):
            self.env.reward_sums(reward, exchange=self.exchange, out=(self.sums, self.total))
            self.reduced.record(xs)
        self.pending = True

    def wait_total(self):
        """Make the current stream wait for the (overlapped) exchange of the last episode."""
        if self.overlap and self.pending:
            torch.cuda.current_stream(self.env.device).wait_event(self.reduced)


class EpisodeRunner(object):
    """static [B,rows,S] / dynamic [B,3n,S] / ptr_seq [steps,B]  -- one window per episode (training), or the same
    tensors with a leading window axis ([Wn,B,...], [Wn,B,...], [Wn,steps,B]) for rolling-style episodes in which ONE
    container keeps filling while the network window is refilled Wn times (rolling.py:575-658): the container
    is cleared once, every later window only recomputes the masks (BatchedContainers.initial_mask)."""

    def __init__(self, env, static, dynamic, ptr_seq, use_graph=False, partial_sums=True, exchange=None, packed=None,
                 overlap_exchange=None):
        assert isinstance(env, BatchedContainers)
        self.env = env
        self.tail = RewardTail(env, partial_sums, exchange, overlap_exchange)
        # packed = (static_u8 [B,rows,S] uint8, dynamic_bits [B,words] int32) device buffers in the compact upload format
        # (tapenv.pack_inputs): the episode then starts with reset_packed, which fills `static` / `dynamic` from them
        self.packed = packed
        self.exchange = exchange                      # tapenv.dist.PeerExchange: fuse the cross-GPU reward reduction
        dev = env.device
        B, S = env.batch_size, env.S
        if static.dim() == 3:
            static, dynamic, ptr_seq = static.unsqueeze(0), dynamic.unsqueeze(0), ptr_seq.unsqueeze(0)
        self.static = static
        self.dynamic = dynamic
        self.ptr_seq = ptr_seq                        # int64 [Wn, steps, B] on device
        self.windows = int(ptr_seq.shape[0])
        self.steps = int(ptr_seq.shape[1])
        f32 = dict(dtype=torch.float32, device=dev)
        self.dyn_buf = [torch.empty_like(dynamic[0]), torch.empty_like(dynamic[0])]
        self.cur_buf = [torch.empty(B, S, **f32), torch.empty(B, S, **f32)]
        self.mask_buf = [torch.empty(B, S, **f32), torch.empty(B, S, **f32)]
        self.dec_static = torch.empty(B, env.cfg.static_rows - 1, **f32)
        self.dec_dyn = torch.empty(B, env.enc_len, **f32)
        self.reward_buf = torch.empty(B, **f32)
        self.reward = None
        self.partial_sums = partial_sums
        # reset / initial mask + fused steps per window (the last one also emits the rewards) + the sums (+ exchange) launch
        self.launches_per_episode = self.windows * (1 + self.steps) + self.tail.launches
        self.graph = None
        if use_graph:
            self._capture()

    def _episode(self):
        env = self.env
        for w in range(self.windows):
            if self.packed is not None:
                assert self.windows == 1
                _, _, cur, mask = env.reset_packed(self.packed[0], self.packed[1],
                                                   out=(self.static[0], self.dynamic[0], self.cur_buf[1], self.mask_buf[1]))
            else:
                cur, mask = env.reset(self.dynamic[w]) if w == 0 else env.initial_mask(self.dynamic[w])
            dyn = self.dynamic[w]
            for t in range(self.steps):
                out = (self.dyn_buf[t & 1], self.cur_buf[t & 1], self.mask_buf[t & 1], self.dec_static, self.dec_dyn)
                last = w == self.windows - 1 and t == self.steps - 1          # calc_ratio rides on the last decode step
                dyn, cur, mask, _, _ = env.step(self.ptr_seq[w, t], self.static[w], dyn, mask, out=out,
                                                reward_out=self.reward_buf if last else None)
        self.reward = self.reward_buf
        if self.steps == 0:
            self.reward = env.calc_ratio()
        self.tail.inline(self.reward)
        self.final = (dyn, cur, mask)

    @property
    def sums(self):
        return self.tail.sums

    @property
    def total(self):
        return self.tail.total

    def _capture(self):
        s = torch.cuda.Stream(device=self.env.device)
        s.wait_stream(torch.cuda.current_stream(self.env.device))
        with torch.cuda.stream(s):
            self._episode()                           # warm-up outside capture
        torch.cuda.current_stream(self.env.device).wait_stream(s)
        torch.cuda.synchronize(self.env.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._episode()
        self.graph = g

    def run(self):
        """One episode for the whole batch; returns the f32 [B] reward tensor (calc_ratio, not negated).
        With an overlapped exchange `sums` / `total` belong to the exchange's side stream: tail.wait_total() first."""
        self.tail.before_episode()
        if self.graph is not None:
            self.graph.replay()
        else:
            self._episode()
        self.tail.after_episode(self.reward)
        return self.reward


def _carve(buf, specs):
    """Views (dtype, shape) laid out back to back in the uint8 buffer `buf`, each 256-byte aligned."""
    out, off = [], 0
    for dtype, shape in specs:
        n = 1
        for v in shape:
            n *= int(v)
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        out.append(buf[off: off + nbytes].view(dtype).view(*shape))
        off = (off + nbytes + 255) // 256 * 256
    return out


def _carve_bytes(specs):
    off = 0
    for dtype, shape in specs:
        n = 1
        for v in shape:
            n *= int(v)
        off = (off + n * torch.empty((), dtype=dtype).element_size() + 255) // 256 * 256
    return max(off, 256)


class HostBatch(object):
    """One episode's inputs in ONE contiguous pinned host buffer (HostPipeline.new_host_batch): `static`, `dynamic`,
    `ptr` are views a loader fills in place; submit() then uploads the batch with a single H2D copy."""

    def __init__(self, specs):
        self.buffer = torch.empty(_carve_bytes(specs), dtype=torch.uint8).pin_memory()
        self.static, self.dynamic, self.ptr = _carve(self.buffer, specs)


class HostPipeline(object):
    """Episodes whose inputs live in HOST memory: pipelined H2D upload of (static, dynamic, ptr_seq) on a copy stream,
    episode replay on the compute stream, D2H of the rewards -- the upload of the next episodes overlaps the kernels of
    the current one.  This is the end-to-end path a trainer's DataLoader drives (trainer.py:189-192: one .cuda() per batch).

        pipe = HostPipeline(env, steps)
        pipe.submit(static_pinned, dynamic_pinned, ptr_seq_pinned)     # returns immediately; three H2D copies
        hb = pipe.new_host_batch(); hb.static[...] = ...; pipe.submit(hb)   # or: one contiguous pinned batch, ONE copy
        rewards, sums = pipe.result()                                 # pinned f32 [B] / f64 [3] of the OLDEST submitted episode
    """

    def __init__(self, env, steps, depth=3, use_graph=True, windows=1, exchange=None, packed=False):
        """packed=True: submit() takes (static_u8, dynamic_bits, ptr_seq) in the compact format of tapenv.pack_inputs /
        PACKDataset.packed() -- 20x fewer PCIe bytes; the fp32 tensors are produced on the device (reset_packed)."""
        self.env = env
        self.packed = packed
        dev = env.device
        B, S = env.batch_size, env.S
        cfg = env.cfg
        self.depth = depth
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.slots = []
        if packed:
            import ctypes as C
            from . import _capi
            words = int(_capi.lib.tapenv_packed_words(C.byref(cfg)))
            self.specs = [(torch.uint8, (B, cfg.static_rows, S)), (torch.int32, (B, words)), (torch.int64, (windows, steps, B))]
        else:
            self.specs = [(torch.float32, (windows, B, cfg.static_rows, S)), (torch.float32, (windows, B, cfg.dyn_rows, S)),
                          (torch.int64, (windows, steps, B))]
        for _ in range(depth):
            staging = torch.zeros(_carve_bytes(self.specs), dtype=torch.uint8, device=dev)      # same layout as a HostBatch
            a, b_, pq = _carve(staging, self.specs)
            if packed:
                st = torch.empty(windows, B, cfg.static_rows, S, dtype=torch.float32, device=dev)
                dy = torch.empty(windows, B, cfg.dyn_rows, S, dtype=torch.float32, device=dev)
                pk = (a, b_)
            else:
                st, dy, pk = a, b_, None
            runner = EpisodeRunner(env, st, dy, pq, use_graph=use_graph, partial_sums=True, exchange=exchange, packed=pk)
            self.slots.append(dict(staging=staging, static=a, dynamic=b_, ptr=pq, runner=runner,
                                   uploaded=torch.cuda.Event(), consumed=torch.cuda.Event(), done=torch.cuda.Event(),
                                   reward=torch.empty(B, dtype=torch.float32).pin_memory(),
                                   sums=torch.empty(3, dtype=torch.float64).pin_memory(), busy=False))
        self.head = 0      # next slot to submit into
        self.tail = 0      # oldest slot in flight
        self.inflight = 0
        self.h2d_bytes = windows * ((B * cfg.static_rows * S + B * cfg.dyn_rows * S) * 4 + steps * B * 8)
        if packed:
            self.h2d_bytes = B * cfg.static_rows * S + B * words * 4 + steps * B * 8
        self.d2h_bytes = B * 4 + 24

    def new_host_batch(self):
        """A pinned host batch with the staging layout of this pipeline (fill .static / .dynamic / .ptr in place)."""
        return HostBatch(self.specs)

    def submit(self, static_h, dynamic_h=None, ptr_h=None, after_episode=None):
        if self.inflight == self.depth:
            raise RuntimeError("pipeline full: call result() first")
        s = self.slots[self.head]
        compute = torch.cuda.current_stream(self.env.device)
        with torch.cuda.stream(self.copy_stream):
            if s["busy"]:
                self.copy_stream.wait_event(s["consumed"])       # the slot's previous episode has read its inputs
            if isinstance(static_h, HostBatch):                  # one contiguous pinned batch: ONE H2D copy
                s["staging"].copy_(static_h.buffer, non_blocking=True)
            else:
                s["static"].copy_(static_h.view_as(s["static"]), non_blocking=True)
                s["dynamic"].copy_(dynamic_h.view_as(s["dynamic"]), non_blocking=True)
                s["ptr"].copy_(ptr_h.view_as(s["ptr"]), non_blocking=True)
            s["uploaded"].record(self.copy_stream)
        compute.wait_event(s["uploaded"])
        r = s["runner"].run()
        s["consumed"].record(compute)
        if after_episode is not None:
            after_episode(s["runner"])                           # e.g. the cross-rank reduction of runner.sums
        s["reward"].copy_(r, non_blocking=True)
        tail = s["runner"].tail
        if tail.overlap:                                         # the statistics arrive on the side stream
            with torch.cuda.stream(tail.stream):
                s["sums"].copy_(tail.total if tail.total is not None else tail.sums, non_blocking=True)
                tail.reduced.record(tail.stream)                 # "consumed" now includes the D2H read of the totals
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
