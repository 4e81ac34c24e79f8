"""The pointer-network decode loop of model.DRL.forward (model.py:342-515) with the environment on the device and no
host round trips.

The reference's loop synchronises with the host several times per step -- `mask.byte().any()` (model.py:344), the
re-sampling `while` (:367), `.cpu().numpy()` of the chosen blocks (:407-412), B Python `add_new_block` calls and a
`torch.FloatTensor(heightmaps).cuda()` upload (:452-458).  None of them is needed: every step removes exactly one block
from EVERY environment, so an episode is exactly `blocks_num` steps; a candidate whose mask entry is 0 has probability 0
after `softmax(logits + log(mask))`, so the re-sampling loop never fires; and the fused step returns decoder_static /
decoder_dynamic on the device.  `DecodeLoop` is that loop; the network stays the caller's:

    def actor_step(static, dynamic, decoder_static, decoder_dynamic, state):     # the reference's encoders + pointer
        ...
        return logits, state                                                     # logits f32 [B,S] BEFORE masking

    loop = tapenv.DecodeLoop(env, actor_step, greedy=False)
    tour_idx, tour_logp, reward = loop.run(static, dynamic)                      # == DRL.forward's (tour_idx, tour_logp, _, -R)

With `use_graph=True` the whole episode (actor included) is captured in one CUDA graph on first use -- the actor must
then be capturable (static shapes, no host syncs) and `run` must be fed tensors of the same shapes (they are copied
into the captured input buffers).
"""
import torch

from .containers import BatchedContainers


class DecodeLoop(object):
    def __init__(self, env, actor_step, greedy=False, use_graph=False, generator=None):
        assert isinstance(env, BatchedContainers)
        self.env, self.actor_step, self.greedy = env, actor_step, greedy
        self.use_graph, self.generator = use_graph, generator
        self.steps = env.window                       # blocks the network sees per episode (model.py:342: sequence_size / rotations)
        self._graph = None

    def _episode(self, static, dynamic):
        env = self.env
        B = env.batch_size
        current_mask, mask = env.reset(dynamic)                                  # model.py:294-307
        dec_static = torch.zeros(B, env.cfg.static_rows - 1, device=env.device)  # PACKDataset's zero decoder inputs (pack.py:228-266)
        dec_dyn = env._shape_enc(torch.zeros(B, env.enc_len, device=env.device))
        state = None
        idx, logps = [], []
        for _ in range(self.steps):
            logits, state = self.actor_step(static, dynamic, dec_static, dec_dyn, state)
            probs = torch.softmax(logits + current_mask.log(), dim=1)            # model.py:356
            if self.greedy:
                prob, ptr = torch.max(probs, 1)                                  # model.py:370-371
                logp = prob.log()
            else:
                ptr = torch.multinomial(probs, 1, generator=self.generator).squeeze(1)     # Categorical(probs).sample(), model.py:362-364
                logp = torch.log(torch.gather(probs, 1, ptr.unsqueeze(1)).squeeze(1))      # m.log_prob(ptr), model.py:368
            dynamic, current_mask, mask, dec_static, dec_dyn = env.step(ptr, static, dynamic, mask)   # model.py:376-463
            if hasattr(state, "last_ptr"):
                state.last_ptr = ptr                                             # tapenv.adapters: rows zeroed by this step
            idx.append(ptr)
            logps.append(logp)
        reward = env.calc_ratio()                                                # model.py:499-515
        return torch.stack(idx, 1), torch.stack(logps, 1), reward

    def run(self, static, dynamic):
        """-> (tour_idx int64 [B,n], tour_logp f32 [B,n], reward f32 [B] = calc_ratio, NOT negated)."""
        if not self.use_graph:
            return self._episode(static, dynamic)
        if self._graph is None:
            self._static_in, self._dynamic_in = static.clone(), dynamic.clone()
            s = torch.cuda.Stream(device=self.env.device)
            s.wait_stream(torch.cuda.current_stream(self.env.device))
            with torch.cuda.stream(s):
                self._episode(self._static_in, self._dynamic_in)                 # warm-up outside capture
            torch.cuda.current_stream(self.env.device).wait_stream(s)
            torch.cuda.synchronize(self.env.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._out = self._episode(self._static_in, self._dynamic_in)
            self._graph = g
        self._static_in.copy_(static)
        self._dynamic_in.copy_(dynamic)
        self._graph.replay()
        return self._out


class RollingDecodeLoop(object):
    """rolling.validate's loop (rolling.py:589-640) for a batch with the network as a callback: every step the actor sees
    the CURRENT window (static, dynamic) and the decoder inputs of the previous placement, picks a candidate (greedy like
    actor.eval(), or sampled), and ONE launch places the block and rebuilds the window (RollingRunner.step).

        loop = tapenv.RollingDecodeLoop(tapenv.RollingRunner(env, windows), actor_step)
        tour_idx, tour_logp, reward = loop.run()          # tour_idx[:, t] = pointer into the window of step t
    """

    def __init__(self, runner, actor_step, greedy=True, generator=None):
        self.runner, self.actor_step, self.greedy, self.generator = runner, actor_step, greedy, generator

    def run(self):
        run = self.runner
        env = run.env
        B = env.batch_size
        static, dynamic, current_mask = run.begin()
        dec_static = torch.zeros(B, env.block_dim, device=env.device)            # RollingDataset's zero decoder inputs (rolling.py:521-533)
        dec_dyn = env._shape_enc(torch.zeros(B, env.enc_len, device=env.device))
        state = None
        idx, logps = [], []
        for _ in range(run.total):
            logits, state = self.actor_step(static, dynamic, dec_static, dec_dyn, state)
            probs = torch.softmax(logits + current_mask.log(), dim=1)            # rolling.py:389
            if self.greedy:
                prob, ptr = torch.max(probs, 1)                                  # rolling.py:399-400
                logp = prob.log()
            else:
                ptr = torch.multinomial(probs, 1, generator=self.generator).squeeze(1)
                logp = torch.log(torch.gather(probs, 1, ptr.unsqueeze(1)).squeeze(1))
            static, dynamic, current_mask, dec_static, dec_dyn = run.step(ptr)
            idx.append(ptr)
            logps.append(logp)
        return torch.stack(idx, 1), torch.stack(logps, 1), env.calc_ratio()
