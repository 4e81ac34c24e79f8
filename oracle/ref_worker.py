"""TEST INFRASTRUCTURE ONLY -- the UNMODIFIED Python reference's environment path, timed on host cores.

This is the reference arm of bench.py (`--impl reference`, `cpu_baseline.kind == "reference"`): worker processes import the
reference's own `pack` / `tools` modules (oracle/refshim.py: /root/reference in the build container, the staged oracle/_ref
copy on the GPU box) and run, per environment shard, exactly what model.DRL.forward runs around the network
(model.py:294-307, :376-384, :404-412, :452-453, :509-510):

    containers = [tools.Container(...) for _ in range(B)]          # model.py:294
    initial mask                                                   # model.py:297-307
    per decode step: pack.update_dynamic, pack.update_mask (torch CPU ops), the gather of the chosen blocks,
                     for b in range(B): containers[b].add_new_block(blocks[b], is_rotate[b])     # the per-env Python loop
    scores[b] = containers[b].calc_ratio()                         # model.py:509-510

The pointer of every step is the recorded policy of bench.py (`policy_pick`: the floor(u*count)-th accessible candidate of
the CURRENT mask, u from a fixed RandomState) -- computed from this process's own masks, so the GPU arm and this arm follow
identical trajectories iff their masks agree.  The reference is single-threaded (trainer.py:155, no multiprocessing): one
worker = the faithful 1-core figure; `nproc` workers over disjoint shards = all host cores.
Nothing here is imported by tap-net_b200/.
"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def policy_pick(cur_mask, u):
    """cur_mask [B,S] (0/1), u [B] uniform in [0,1) -> int64 [B]: the floor(u*count)-th accessible candidate."""
    m = np.asarray(cur_mask) > 0
    cnt = m.sum(1)
    k = np.minimum((np.asarray(u) * cnt).astype(np.int64), np.maximum(cnt - 1, 0))
    return (np.cumsum(m, 1) > k[:, None]).argmax(1).astype(np.int64)


def reference_episode(mods, static, dynamic, u, size, reward_type, heightmap_type, strategy, trace=False):
    """One episode of the reference's env path for a shard.  static [B,1+dim,S], dynamic [B,3n,S] float32 numpy,
    u [n,B].  Returns dict(reward f32 [B], seconds, and with trace=True the per-step heightmaps / masks / pointers)."""
    import torch
    pack, tools = mods["pack"], mods["tools"]
    st, dyn = torch.from_numpy(static), torch.from_numpy(dynamic)
    B, rows, S = static.shape
    dim = rows - 1
    R = 2 if dim == 2 else 6
    n = S // R
    t0 = time.perf_counter()
    containers = [tools.Container(size, n, reward_type, heightmap_type, packing_strategy=strategy) for _ in range(B)]
    mask = torch.ones(B, S)
    current_mask = mask.clone()                                       # model.py:301-307
    move_mask = dyn[:, :n, :].sum(1)
    rotate_mask = dyn[:, n:2 * n, :].sum(1) * dyn[:, 2 * n:3 * n, :].sum(1)
    current_mask[(rotate_mask + move_mask).ne(0)] = 0.
    tr = dict(ptr=[], heightmap=[], cur_mask=[current_mask.numpy().copy()], dec_dyn=[]) if trace else None
    for t in range(n):
        ptr = torch.from_numpy(policy_pick(current_mask.numpy(), u[t]))
        dyn = pack.update_dynamic(dyn, st, ptr, "bot", True)          # model.py:376
        current_mask, mask = pack.update_mask(mask, dyn, st, ptr, "bot", True)   # model.py:384
        decoder_static = torch.gather(st[:, 1:, :], 2, ptr.view(-1, 1, 1).expand(-1, dim, 1))    # model.py:404-406
        is_rotate = (ptr < n).numpy().astype("bool")
        blocks = decoder_static.transpose(2, 1).squeeze(1).numpy()    # model.py:412
        heightmaps = []
        for batch_index in range(B):                                  # model.py:452-453
            heightmaps.append(containers[batch_index].add_new_block(blocks[batch_index], is_rotate[batch_index]))
        if trace:
            tr["ptr"].append(ptr.numpy().copy())
            tr["heightmap"].append(np.stack([np.asarray(c.heightmap).reshape(-1) for c in containers]))
            tr["cur_mask"].append(current_mask.numpy().copy())
            tr["dec_dyn"].append(np.stack([np.asarray(h, dtype=np.float32).reshape(-1) for h in heightmaps]))
    scores = torch.zeros(B)
    for batch_index in range(B):                                      # model.py:509-510
        scores[batch_index] = containers[batch_index].calc_ratio()
    out = {"reward": scores.numpy().copy(), "seconds": time.perf_counter() - t0}
    if trace:
        out.update({k: np.stack(v) for k, v in tr.items()})
        out["positions"] = np.stack([np.asarray(c.positions) for c in containers])
    return out


def _worker_main(conn, ref_dir):
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["MKL_NUM_THREADS"] = "1"
    if ref_dir:
        os.environ["TAPNET_REFERENCE"] = ref_dir
    for p in (ROOT,):
        if p not in sys.path:
            sys.path.insert(0, p)
    import io
    import contextlib
    import torch
    torch.set_num_threads(1)
    from oracle import refshim
    with contextlib.redirect_stdout(io.StringIO()):
        mods = refshim.load(("tools", "pack"))
    shard = None
    conn.send(("ready", refshim.REFERENCE_DIR))
    while True:
        msg = conn.recv()
        if msg[0] == "load":
            shard = msg[1]
            conn.send(("ok",))
        elif msg[0] == "run":
            with contextlib.redirect_stdout(io.StringIO()):
                res = reference_episode(mods, shard["static"], shard["dynamic"], shard["u"], shard["size"], shard["reward_type"],
                                        shard["heightmap_type"], shard["strategy"], trace=msg[1])
            conn.send(("done", res))
        elif msg[0] == "stop":
            conn.send(("bye",))
            return


class ReferencePool(object):
    """`workers` processes, each holding one shard of environments; run() = one episode on every shard in parallel."""

    def __init__(self, workers, ref_dir=None):
        ctx = mp.get_context("spawn")
        self.procs, self.conns = [], []
        for _ in range(workers):
            a, b = ctx.Pipe()
            p = ctx.Process(target=_worker_main, args=(b, ref_dir), daemon=True)
            p.start()
            self.procs.append(p); self.conns.append(a)
        self.source = None
        for c in self.conns:
            tag, src = c.recv()
            assert tag == "ready"
            self.source = src

    def load(self, shards):
        assert len(shards) == len(self.conns)
        for c, s in zip(self.conns, shards):
            c.send(("load", s))
        for c in self.conns:
            assert c.recv()[0] == "ok"

    def run(self, trace=False, only=None):
        """-> list of per-shard results.  only: indices of the workers to use (the others stay idle)."""
        idx = list(range(len(self.conns))) if only is None else list(only)
        for i in idx:
            self.conns[i].send(("run", trace))
        return [self.conns[i].recv()[1] for i in idx]

    def close(self):
        for c in self.conns:
            try:
                c.send(("stop",)); c.recv()
            except Exception:
                pass
        for p in self.procs:
            p.join(timeout=5)
            if p.is_alive():
                p.kill()


def make_shards(static, dynamic, u, size, reward_type, heightmap_type, strategy, workers, per_worker):
    """Disjoint shards of `per_worker` environments (wrapping around the batch if workers*per_worker exceeds it)."""
    B = static.shape[0]
    shards, ranges = [], []
    for w in range(workers):
        idx = (np.arange(per_worker) + w * per_worker) % B
        shards.append(dict(static=np.ascontiguousarray(static[idx]), dynamic=np.ascontiguousarray(dynamic[idx]),
                           u=np.ascontiguousarray(u[:, idx]), size=list(size), reward_type=reward_type,
                           heightmap_type=heightmap_type, strategy=strategy))
        ranges.append(idx)
    return shards, ranges
