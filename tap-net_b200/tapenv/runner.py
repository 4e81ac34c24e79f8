"""Replay of whole packing episodes through the per-step C-ABI calls, with preallocated ping-pong
buffers and optional CUDA-graph capture of the n-step launch sequence (reset + n fused steps + reward).

This is the host-side loop model.py:294-515 runs around the environment, minus the network: the
pointer for every step comes from a `ptr_seq` tensor (a recorded policy) instead of the actor.
"""
import torch

from .containers import BatchedContainers


class EpisodeRunner(object):
    def __init__(self, env, static, dynamic, ptr_seq, use_graph=False, partial_sums=True):
        assert isinstance(env, BatchedContainers)
        self.env = env
        dev = env.device
        B, S = env.batch_size, env.S
        self.static = static
        self.dynamic = dynamic
        self.ptr_seq = ptr_seq                        # int64 [steps, B] on device
        self.steps = int(ptr_seq.shape[0])
        f32 = dict(dtype=torch.float32, device=dev)
        self.dyn_buf = [torch.empty_like(dynamic), torch.empty_like(dynamic)]
        self.cur_buf = [torch.empty(B, S, **f32), torch.empty(B, S, **f32)]
        self.mask_buf = [torch.empty(B, S, **f32), torch.empty(B, S, **f32)]
        self.dec_static = torch.empty(B, env.cfg.static_rows - 1, **f32)
        self.dec_dyn = torch.empty(B, env.enc_len, **f32)
        self.reward = None
        self.sums = None
        self.partial_sums = partial_sums
        self.launches_per_episode = 1 + self.steps + 1 + (1 if partial_sums else 0)
        self.graph = None
        if use_graph:
            self._capture()

    def _episode(self):
        env = self.env
        cur, mask = env.reset(self.dynamic)
        dyn = self.dynamic
        for t in range(self.steps):
            out = (self.dyn_buf[t & 1], self.cur_buf[t & 1], self.mask_buf[t & 1], self.dec_static, self.dec_dyn)
            dyn, cur, mask, _, _ = env.step(self.ptr_seq[t], self.static, dyn, mask, out=out)
        res = env.calc_ratio(partial_sums=self.partial_sums)
        self.reward, self.sums = res if self.partial_sums else (res, None)
        self.final = (dyn, cur, mask)

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
        """One episode for the whole batch; returns the f32 [B] reward tensor (calc_ratio, not negated)."""
        if self.graph is not None:
            self.graph.replay()
        else:
            self._episode()
        return self.reward
