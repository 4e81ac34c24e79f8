#!/usr/bin/env python
"""bench.py -- env-steps/sec of the TAP packing-environment step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

One bench "step" = ONE EPISODE of the hot path over one batch: K0 reset + n fused decode-step launches
(update_dynamic + update_mask + add_new_block for all B environments) + K6 reward.  It advances B*n
env-steps; reset and reward are inside the timed region but are not counted as env-steps
(SURVEY.md section 8d).  The pointer for every decode step is a recorded random-valid policy
(ptr ~ multinomial(current_mask), fixed seed) replayed from memory.

Prints ONE JSON line (rank 0).  Keys: see the contract in DESIGN.md section "Measurement".
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tap-net_b200"))

METRIC = "env-steps/sec 2D LB_GREEDY 10-block W=5 at 1/2/4/8 B200 vs CPU ref"
UNIT = "env-steps/s"

WORKLOADS = {
    # name: (fixture, container_size, reward_type, heightmap_type, packing_strategy, default batch, description)
    "c2": ("rand2d_n10.npz", [5, 50], "C+P+S-lb-soft", "diff", "LB_GREEDY", 4096,
           "2D RAND nodes=10 width=5 LB_GREEDY C+P+S-lb-soft batch=4096 (BASELINE configs[1])"),
    "c3": ("rand3d_n10.npz", [5, 5, 50], "C+P+S-lb-soft", "diff", "LB_GREEDY", 4096,
           "3D RAND nodes=10 width=5 LB_GREEDY C+P+S-lb-soft batch=4096 (BASELINE configs[2])"),
    # container height 100 (not the trainer default 50): under a RANDOM policy a few 20-block stacks pass 50, where the
    # reference raises IndexError; placements are identical for any height that is not reached
    "c4": ("ppsg2d_n20.npz", [7, 100], "C+P+S-mcs-hard", "diff", "MACS", 1024,
           "2D PPSG nodes=20 width=7 height=100 MACS C+P+S-mcs-hard batch=8192/8 per GPU (BASELINE configs[3])"),
    # rolling inference (rolling.py:575-640): ONE container [5,5,250] takes 50 blocks while the network window stays at 10;
    # the window is rebuilt before every decode step by generate.InitialContainer -> tapenv_rolling_step (see run_rolling)
    "c5": ("rolling3d_t50.npz", [5, 5, 250], "C+P+S-lb-soft", "diff", "LB_GREEDY", 8192,
           "3D rolling total=50 window=10 width=5 height=250 LB_GREEDY batch=65536/8 per GPU (BASELINE configs[4])"),
}
WINDOWS = {}
ROLLING = {"c5": (50, 10)}          # workload -> (total_blocks_num, network window)
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the step kernel at the workload's default batch, from the
# `ncu --set full` captures summarised under profiles/ (writes stay in the 126 MB L2 at these sizes)
TRAFFIC = {"c2": 11.34e6, "c3": 34.14e6, "c4": 10.53e6, "c5": 40.42e6}    # profiles/r01p_*_ncu_full_summary.csv (read + write)


def algorithmic_bytes_per_env_step(n, R, dim, W, L, macs):
    """minimum HBM traffic of one env-step under the reference's tensor contract."""
    S = n * R
    cells = W if dim == 2 else W * L
    dec_dyn = (W - 1) if dim == 2 else 2 * W * L
    hist = (16 * n + 16) if macs else 0
    return (2 * (3 * n * S * 4) + 3 * (S * 4) + 8 + 4 * (1 + dim) + 2 * 4 * cells + 2 * 16 + hist + 4 * dec_dyn + 4 * dim)


def load_workload(name, batch, rank):
    """-> static [Wn,B,1+dim,S], dynamic [Wn,B,3n,S] (Wn = windows per episode, 1 except for the rolling-style c5)."""
    from tests.golden_io import load_inputs
    fixture, size, rt, hm, strat, default_b, desc = WORKLOADS[name]
    B = batch or default_b
    static, dynamic = load_inputs(fixture)
    pool = static.shape[0]
    Wn = WINDOWS.get(name, 1)
    idx = (np.arange(Wn)[:, None] * 389 + np.arange(B)[None, :] + rank * 977) % pool   # tile the pool; ranks / windows start at different offsets
    return np.ascontiguousarray(static[idx]), np.ascontiguousarray(dynamic[idx]), size, rt, hm, strat, B, desc, pool


# ----------------------------------------------------------------------------------------------
# clocks (B200_PROFILING.md: sample DURING the timed region)
# ----------------------------------------------------------------------------------------------
class ClockSampler(object):
    """SM clock + throttle reasons sampled WHILE the GPU executes the timed region: the timed loops only enqueue work, so
    after the last enqueue the launching thread polls NVML until the closing event has completed (`sample_until`).  No
    background thread: a polling thread competes with the launching thread for the GIL and with the DMA engines for the
    PCIe link -- at a 2 ms period it cut the pinned H2D rate of the e2e leg from 55 to 32 GB/s (scripts/h2d_probe.py)."""

    NAMES = (("HwSlowdown", 0x8, "hw_slowdown"), ("HwThermalSlowdown", 0x40, "hw_thermal_slowdown"),
             ("SwThermalSlowdown", 0x20, "sw_thermal_slowdown"), ("SwPowerCap", 0x4, "sw_power_cap"))

    def __init__(self, index):
        self.samples = []
        self.reasons = set()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.bits = {getattr(pynvml, "nvmlClocksEventReason" + a, getattr(pynvml, "nvmlClocksThrottleReason" + a, d)): nm
                         for a, d, nm in self.NAMES}
            self.ok = True
        except Exception:
            self.sm_max = None

    def start(self):                      # kept for call-site compatibility: nothing runs in the background
        return self

    def sample(self):
        if not self.ok:
            return
        nv = self.nv
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            try:
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for bit, nm in self.bits.items():
                if r & bit:
                    self.reasons.add(nm)
        except Exception:
            pass

    def sample_until(self, event, max_samples=64):
        """Poll while the GPU is still working towards `event` (a recorded torch.cuda.Event); at least one sample."""
        self.sample()
        n = 1
        while not event.query() and n < max_samples:
            self.sample()
            n += 1

    def result(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores (the Python reference cannot travel to the GPU box)
# ----------------------------------------------------------------------------------------------
def host_policy(static, dynamic, size, seed):
    """Recorded random-valid policy for the CPU arm (checker-side code: uses the oracle's mask functions).
    static/dynamic carry the window axis; returns ptr_seq [Wn, n, B]."""
    from tests.rollout import random_valid_ptrs
    return np.stack([random_valid_ptrs(static[w], dynamic[w], size, seed=seed + w) for w in range(static.shape[0])])


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    static, dynamic, size, rt, hm, strat, B, desc, pool = load_workload(args.workload, args.batch, 0)
    threads = os.cpu_count() or 1
    ptr_seq = host_policy(static, dynamic, size, seed=1234)
    Wn, n = ptr_seq.shape[0], ptr_seq.shape[1]
    from oracle import oracle
    kw = dict(nthreads=threads, want=("reward",), capacity=Wn * n)
    for _ in range(max(args.warmup, 1)):
        oracle.episode_batch(static, dynamic, ptr_seq, size, rt, hm, strat, **kw)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o = oracle.episode_batch(static, dynamic, ptr_seq, size, rt, hm, strat, **kw)
    el = time.perf_counter() - t0
    value = args.steps * B * n * Wn / el
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32 state + f64 score", "data": "synthetic",
        "config": {"workload": desc, "batch": B, "blocks": n * Wn, "env_steps_per_step": B * n * Wn,
                   "inputs": "reference RAND/PPSG generator fixtures (tests/golden), pool of %d tiled" % pool},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d episodes x %d envs x %d steps, oracle/tap_oracle.c on %d pthreads" % (args.steps, B, n * Wn, threads)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference is pure Python/NumPy and /root/reference is not on the GPU box: this arm times the C "
                "restatement (oracle/) of the same per-step path on all host threads; the Python reference itself "
                "measured 4.9e3 env-steps/s on one core (BASELINE.md section 2)",
        "reward_mean": float(o["reward"].mean()),
    }
    print(json.dumps(line))
    return 0



# ----------------------------------------------------------------------------------------------
# rolling workload (BASELINE configs[4]): generate.InitialContainer window + one container per instance
# ----------------------------------------------------------------------------------------------
def rolling_bytes_per_env_step(T, n, dim, W, L):
    """Algorithmic HBM bytes of ONE fused rolling decode step (tapenv_rolling_step) per instance: the tensors the next
    network call consumes (static, dynamic, two masks) are written, the instance graphs / blocks / states are read."""
    R = 2 if dim == 2 else 6
    S = n * R
    cells = W if dim == 2 else W * L
    dec_dyn = (W - 1) if dim == 2 else 2 * W * L
    out = (1 + dim) * S * 4 + 3 * n * S * 4 + 2 * S * 4 + n * 4 + 4 + 4 * dim + 4 * dec_dyn
    graphs = T * 8 + (2 if dim == 2 else 4) * n * 8           # movement predecessors of every node + rotation graphs of the window
    blocks = S * dim * 4
    state = 2 * 64 + 2 * 4 * cells + 2 * 16 + 2 * 4 * dim + 1  # window state r/w, heightmap r/w, scalars r/w, position/block/stable
    return out + graphs + blocks + state + 8


def load_rolling_workload(name, batch, rank):
    from tests.golden_io import load_rolling
    fixture, size, rt, hm, strat, default_b, desc = WORKLOADS[name]
    T, n = ROLLING[name]
    B = batch or default_b
    z = load_rolling(fixture)
    pool = z["adj"].shape[0]
    idx = (np.arange(B) + rank * 977) % pool
    return z["adj"][idx], np.ascontiguousarray(z["blocks"][idx]), size, rt, hm, strat, B, desc, pool, T, n, z["dim"]


def rolling_host_policy(adj, blocks, T, n, dim, seed):
    """Recorded random-valid policy for the CPU arm (checker-side code: drives the oracle's window)."""
    from oracle import oracle
    B = adj.shape[0]
    R = 2 if dim == 2 else 6
    rng = np.random.RandomState(seed)
    ptr_seq = np.zeros((T, B), np.int64)
    for b in range(B):
        ic = oracle.InitialContainer(adj[b], blocks[b], T, n, dim)
        t = 0
        while t < T:
            static, dynamic = ic.convert_to_input()
            last = ic.is_last_graph()
            mask = np.ones((1, n * R), np.float32)
            cur = oracle.initial_mask(dynamic[None], n, R)
            dyn = dynamic[None]
            for _ in range(n if last else 1):
                ok = np.nonzero(cur[0] > 0)[0]
                p = int(rng.choice(ok))
                ptr_seq[t, b] = p
                t += 1
                if last:
                    pa = np.array([p], np.int64)
                    dyn = oracle.update_dynamic(dyn, static[None], pa)
                    cur, mask = oracle.update_mask(mask, dyn, static[None], pa)
            ic.remove_block(ic.sub_graph_nodes[p % n])
    return ptr_seq


def run_rolling_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle
    adj, blocks, size, rt, hm, strat, B, desc, pool, T, n, dim = load_rolling_workload(args.workload, args.batch, 0)
    threads = os.cpu_count() or 1
    Bs = min(B, 1024)                                     # bounded sample: the policy is recorded on the host (slow Python loop)
    adj, blocks = adj[:Bs], blocks[:Bs]
    ptr_seq = rolling_host_policy(adj, blocks, T, n, dim, seed=1234)
    kw = dict(nthreads=threads)
    for _ in range(max(args.warmup, 1)):
        oracle.rolling_batch(adj, blocks, ptr_seq, size, n, rt, hm, strat, **kw)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o = oracle.rolling_batch(adj, blocks, ptr_seq, size, n, rt, hm, strat, **kw)
    el = time.perf_counter() - t0
    value = args.steps * Bs * T / el
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32 state + f64 score", "data": "synthetic",
        "config": {"workload": desc, "batch": Bs, "blocks": T, "env_steps_per_step": Bs * T,
                   "inputs": "reference rolling.get_dataset fixtures (tests/golden), pool of %d tiled" % pool},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d episodes x %d instances x %d steps, oracle/win_oracle.c + tap_oracle.c on %d pthreads" % (args.steps, Bs, T, threads)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "C restatement (oracle/) of rolling.validate's loop on all host threads; the reference itself is Python + networkx",
        "reward_mean": float(o["reward"].mean()),
    }
    print(json.dumps(line))
    return 0


def run_rolling(args):
    import torch
    import torch.distributed as dist
    import tapenv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the tapenv path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    adj, blocks_h, size, rt, hm, strat, B, desc, pool, T, n, dim = load_rolling_workload(args.workload, args.batch, rank)
    graphs_h = tapenv.pack_graphs(adj)
    bytes_step = rolling_bytes_per_env_step(T, n, dim, size[0], size[1] if dim == 3 else 1)
    env = tapenv.BatchedContainers(size, T, rt, hm, packing_strategy=strat, batch_size=B, device=dev, window=n)

    exchange, reduction = None, "none (single GPU)"
    if world > 1:
        exchange = tapenv.dist.PeerExchange(dev)
        reduction = "fused one-shot exchange over NVLink peer memory (tapenv_reward_allreduce), inside the CUDA graph"

    # record the policy once (untimed): ptr ~ multinomial(current_mask of the live window)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    win0 = tapenv.BatchedInitialContainers(graphs_h, blocks_h, T, n, dim, device=dev)
    rec = tapenv.RollingRunner(env, win0)
    static, dynamic, cur = rec.begin()
    ptrs = []
    for t in range(T):
        ptr = torch.multinomial(cur, 1, generator=g).squeeze(1)
        ptrs.append(ptr)
        static, dynamic, cur, _, _ = rec.step(ptr)
    ptr_seq0 = torch.stack(ptrs)
    reward_ref = env.calc_ratio().clone()
    env.check_flags(); win0.check_flags()

    RING = 2                                               # one episode's ping-pong outputs alone exceed the 126 MB L2
    runners = []
    for i in range(RING):
        roll = (i * 131) % B
        win = tapenv.BatchedInitialContainers(torch.roll(win0.graphs, roll, 0), torch.roll(win0.blocks, roll, 0), T, n, dim, device=dev)
        runners.append(tapenv.RollingRunner(env, win, ptr_seq=torch.roll(ptr_seq0, roll, 1).contiguous(),
                                            use_graph=not args.no_graph, partial_sums=True, exchange=exchange))
    per_episode_bytes = sum(t.numel() * 4 for t in runners[0].static + runners[0].dynamic + runners[0].cur + runners[0].mask)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(max(args.warmup, 3)):
        runners[i % RING].run()
    barrier()
    assert torch.equal(runners[0].run(), reward_ref)

    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall = time.perf_counter()
    e0.record()
    for k in range(args.steps):
        runners[k % RING].run()
    e1.record()
    sampler.sample_until(e1)                               # clocks while the GPU works through the timed episodes
    barrier()
    t_wall = time.perf_counter() - t_wall
    tt = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    span_ms_max = float(tt.item())
    value = world * B * T * args.steps / (span_ms_max * 1e-3)
    launches = args.steps * runners[0].launches_per_episode

    # roofline of the dominant kernel: the T-n fused rolling steps of one episode, replayed from a graph
    rr = tapenv.RollingRunner(env, win0, ptr_seq=ptr_seq0)
    nl = T - n

    def roll_steps():
        for t in range(nl):
            rr.step(ptr_seq0[t])

    rr.begin(); roll_steps()
    torch.cuda.synchronize(dev)
    rr.begin()
    rg = torch.cuda.CUDAGraph()
    with torch.cuda.graph(rg):
        roll_steps()
    tot_ms, cnt = 0.0, 0
    for rep in range(8):
        rr.begin()
        torch.cuda.synchronize(dev)
        e0.record(); rg.replay(); e1.record()
        torch.cuda.synchronize(dev)
        if rep >= 2:
            tot_ms += e0.elapsed_time(e1); cnt += nl
    step_us = 1e3 * tot_ms / cnt
    achieved = B * bytes_step / (step_us * 1e-6) / 1e9

    # e2e: host graphs / blocks / pointers in, host rewards out
    gr_pin = torch.from_numpy(graphs_h).pin_memory()
    bl_pin = torch.from_numpy(blocks_h).pin_memory()
    pq_pin = ptr_seq0.cpu().pin_memory()
    from tapenv.rolling import rotation_structured
    pipe = tapenv.RollingHostPipeline(env, T, n, depth=2, use_graph=not args.no_graph, exchange=exchange,
                                      blocks_are_rotations=rotation_structured(blocks_h, T, dim))    # checked once per dataset, untimed

    def e2e_run(k):
        last = None
        for i in range(k):
            if pipe.inflight == pipe.depth:
                last = pipe.result()
            pipe.submit(gr_pin, bl_pin, pq_pin)
        while pipe.inflight:
            last = pipe.result()
        return last

    # warm-up of the host->device path itself: on a fresh box the first ~100 ms of pinned copies run at half rate (PCIe link /
    # IOMMU state); without this the same command measured 8.7e7 on its first run and 1.8e8 on its second (r01x)
    big = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
    big_d = torch.empty_like(big, device=dev)
    for _ in range(24):
        big_d.copy_(big, non_blocking=True)
    torch.cuda.synchronize(dev)
    e2e_run(max(args.warmup, 3) + 16)
    barrier()
    t0 = time.perf_counter()
    rw_pin, sums_pin = e2e_run(args.steps)
    barrier()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * B * T * args.steps / float(te.item())
    assert np.array_equal(rw_pin.numpy(), reward_ref.cpu().numpy())
    clocks = sampler.result()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import oracle
        threads = os.cpu_count() or 1
        ptr_h = ptr_seq0.cpu().numpy()
        oracle.rolling_batch(adj, blocks_h, ptr_h, size, n, rt, hm, strat, nthreads=threads)
        reps, t0 = 0, time.perf_counter()
        while True:
            o = oracle.rolling_batch(adj, blocks_h, ptr_h, size, n, rt, hm, strat, nthreads=threads)
            assert o["status"] == 0
            reps += 1
            el = time.perf_counter() - t0
            if el >= args.cpu_seconds or reps >= 10000:
                break
        cpu = {"value": reps * B * T / el, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "%d episodes x %d instances x %d steps = %.1f s of oracle/win_oracle.c + tap_oracle.c on %d pthreads" % (reps, B, T, el, threads),
               "reward_parity_vs_gpu": bool(np.array_equal(o["reward"], reward_ref.cpu().numpy()))}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": span_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32 state + f64 score, f32 tensors, u64 graph masks", "data": "synthetic",
            "config": {"workload": desc, "batch_per_gpu": B, "blocks": T, "window": n, "env_steps_per_step": world * B * T,
                       "step_definition": "one episode = clear + window reset + first window + %d fused rolling steps (place + remove_block + "
                                          "convert_to_input) + %d fused decode steps in the last window (the last also emits calc_ratio) + reward sums over the batch" % (T - n, n),
                       "l2": "%.0f MB of ping-pong window tensors per episode > 126 MB L2; %d instance sets alternate" % (per_episode_bytes / 1e6, RING),
                       "cuda_graph": not args.no_graph, "reward_reduction": reduction,
                       "inputs": "reference rolling.get_dataset fixtures (tests/golden), pool of %d tiled" % pool,
                       "policy": "recorded ptr ~ multinomial(current_mask), seed 1234+rank",
                       "node_order": "reference (networkx FilterAtlas / CPython set order)"},
            "gpu_launches": launches, "wall_ms_per_step": 1e3 * t_wall / args.steps,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(pipe.h2d_bytes), "d2h_bytes_per_step": int(pipe.d2h_bytes),
                    "ms_per_step": 1e3 * float(te.item()) / args.steps,
                    "api": "tapenv.RollingHostPipeline.submit/result (graphs + blocks + pointers from pinned host memory), rewards to pinned host memory"},
            "roofline": {"bound": "hbm", "kernel": "window_kernel (fused add_new_block + remove_block + convert_to_input)",
                         "achieved": achieved, "peak": peak,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "B200_PROFILING.md fallback (of fallback)",
                         "unit": "GB/s", "frac": achieved / peak, "traffic": TRAFFIC.get(args.workload),
                         "algorithmic_bytes_per_env_step": bytes_step, "bytes_per_launch": B * bytes_step,
                         "launch_us": step_us, "launches_timed": cnt},
            "clocks": clocks,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0

# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import tapenv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the tapenv path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    static_h, dynamic_h, size, rt, hm, strat, B, desc, pool = load_workload(args.workload, args.batch, rank)
    Wn = static_h.shape[0]
    dim = len(size)
    R = 2 if dim == 2 else 6
    S = static_h.shape[3]
    n = S // R
    macs = strat == "MACS" or "mcs" in rt
    bytes_step = algorithmic_bytes_per_env_step(n, R, dim, size[0], size[1] if dim == 3 else 1, macs)
    steps_per_episode = Wn * n

    # ring of RING distinct input sets (same instances, rotated) so every episode reads its inputs from HBM,
    # not from a warm L2: RING * (inputs + ping-pong outputs) >> 126 MB
    per_set = static_h.nbytes + dynamic_h.nbytes
    RING = max(1 if per_set > 300e6 else 2, int(np.ceil(400e6 / (3 * per_set))))
    env = tapenv.BatchedContainers(size, Wn * n, rt, hm, packing_strategy=strat, batch_size=B, device=dev, window=n)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)

    # record the policy once (untimed): ptr ~ multinomial(current_mask)
    st0 = torch.from_numpy(static_h).to(dev)
    dyn0 = torch.from_numpy(dynamic_h).to(dev)
    ptrs = []
    for w in range(Wn):
        cur, mask = env.reset(dyn0[w]) if w == 0 else env.initial_mask(dyn0[w])
        dyn = dyn0[w]
        for t in range(n):
            ptr = torch.multinomial(cur, 1, generator=g).squeeze(1)
            dyn, cur, mask, _, _ = env.step(ptr, st0[w], dyn, mask)
            ptrs.append(ptr)
    ptr_seq0 = torch.stack(ptrs).view(Wn, n, B)
    reward_ref = env.calc_ratio().clone()
    env.check_flags()

    # multi-GPU: the end-of-episode reduction of the reward statistics feeding the critic baseline (trainer.py:216-225).
    # Preferred: fused behind the reward kernel over NVLink peer memory (tapenv_reward_allreduce, in the CUDA graph);
    # otherwise an NCCL all-gather of the f64 triples on a side stream, overlapping the next episode.
    exchange, reducer, reduction = None, None, "none (single GPU)"
    if world > 1 and args.no_reduce:
        reduction = "none (diagnostic run: --no-reduce)"
    elif world > 1 and not args.nccl_reduce:
        try:
            exchange = tapenv.dist.PeerExchange(dev)
            reduction = "fused one-shot exchange over NVLink peer memory (tapenv_reward_allreduce), inside the CUDA graph"
        except Exception as e:                          # no symmetric memory / P2P on this box
            exchange = None
            reduction = "PeerExchange unavailable (%s); " % type(e).__name__
    if world > 1 and exchange is None and not args.no_reduce:
        reducer = tapenv.dist.RewardReducer(dev)
        reduction = (reduction if reduction.startswith("PeerExchange") else "") + "NCCL all_gather of the f64 triples on a side stream"

    runners = []
    for i in range(RING):
        roll = (i * 131) % B
        st = torch.roll(st0, roll, 1).contiguous()
        dy = torch.roll(dyn0, roll, 1).contiguous()
        pq = torch.roll(ptr_seq0, roll, 2).contiguous()
        runners.append(tapenv.EpisodeRunner(env, st, dy, pq, use_graph=not args.no_graph, partial_sums=True, exchange=exchange))

    def episode(i):
        r = runners[i % RING]
        if reducer is not None and getattr(r, "reduced", None) is not None:
            torch.cuda.current_stream().wait_event(r.reduced)     # the slot's previous sums have been consumed
        r.run()
        if reducer is not None:
            r.total, r.reduced = reducer.reduce_async(r.sums)
        return r

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(max(args.warmup, 3)):
        episode(i)
    barrier()
    # parity spot check of the replay against the recorded pass (same kernels; the oracle check lives in tests/ and smoke())
    assert torch.equal(runners[0].run(), reward_ref)

    # ---- timed region 1: `value` -- K episodes, inputs resident in HBM --------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_wall = time.perf_counter()
    for k in range(args.steps):
        ev[k][0].record()
        episode(k)
        ev[k][1].record()
    sampler.sample_until(ev[-1][1])                        # clocks while the GPU works through the timed episodes
    barrier()
    t_wall = time.perf_counter() - t_wall
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    # K episodes back to back on one stream: device time from first start to last end
    span_ms = ev[0][0].elapsed_time(ev[-1][1])
    tt = torch.tensor([span_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    span_ms_max = float(tt.item())
    value = world * B * steps_per_episode * args.steps / (span_ms_max * 1e-3)
    launches = args.steps * runners[0].launches_per_episode

    # ---- timed region 2: roofline of the dominant kernel (the fused step), cold inputs ---------------
    # graph-replayed launches, each on a different (window, ring slot) input set, events on the launching stream
    slots = [(r, w) for r in runners for w in range(Wn)]
    nl = min(len(slots), n)                                # launches per replay (k stays < capacity between clears)
    out_bufs = [(torch.empty_like(dyn0[0]), torch.empty(B, S, device=dev), torch.empty(B, S, device=dev),
                 torch.empty(B, dim, device=dev), torch.empty(B, env.enc_len, device=dev)) for _ in range(nl)]
    mask1 = torch.ones(B, S, device=dev)

    def roof_launches():
        for i in range(nl):
            r, w = slots[i]
            env.step(r.ptr_seq[w, 0], r.static[w], r.dynamic[w], mask1, out=out_bufs[i])

    env.clear_container()
    roof_launches()                                        # warm-up outside capture
    torch.cuda.synchronize(dev)
    rg = torch.cuda.CUDAGraph()                            # the launches are replayed from a graph so that the
    with torch.cuda.graph(rg):                             # host's per-call overhead is not what gets timed
        roof_launches()
    tot_ms, cnt = 0.0, 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(12):
        env.clear_container()
        torch.cuda.synchronize(dev)
        e0.record()
        rg.replay()
        e1.record()
        torch.cuda.synchronize(dev)
        if rep >= 2:
            tot_ms += e0.elapsed_time(e1)
            cnt += nl
    step_us = 1e3 * tot_ms / cnt
    achieved = B * bytes_step / (step_us * 1e-6) / 1e9

    # context for the fraction: the out-of-place copy of `dynamic` alone (the clone every update_dynamic must make,
    # pack.py:370) by torch's copy kernel, same cold ring slots, same graph-replay + event method
    def copy_launches():
        for i in range(nl):
            r, w = slots[i]
            out_bufs[i][0].copy_(r.dynamic[w])

    copy_launches()
    torch.cuda.synchronize(dev)
    cg = torch.cuda.CUDAGraph()
    with torch.cuda.graph(cg):
        copy_launches()
    ctot = 0.0
    for rep in range(12):
        env.clear_container()                              # same L2-disturbing prelude as above
        torch.cuda.synchronize(dev)
        e0.record(); cg.replay(); e1.record()
        torch.cuda.synchronize(dev)
        if rep >= 2:
            ctot += e0.elapsed_time(e1)
    copy_us = 1e3 * ctot / (10 * nl)

    # ---- timed region 3: e2e -- host buffers in, host rewards out, through the public Python API --------
    # tapenv.HostPipeline: per episode one H2D upload of (static, dynamic, ptr_seq) from pinned memory on a copy
    # stream (double-buffered, overlapping the previous episode's kernels), the episode, D2H of rewards + sums.
    pq_pin = ptr_seq0.cpu().pin_memory()
    pipe = tapenv.HostPipeline(env, n, depth=4, use_graph=not args.no_graph, windows=Wn, exchange=exchange)
    hb = pipe.new_host_batch()                             # ONE contiguous pinned batch (what a loader fills in place): one H2D copy per episode
    hb.static.copy_(torch.from_numpy(static_h)); hb.dynamic.copy_(torch.from_numpy(dynamic_h)); hb.ptr.copy_(pq_pin)
    after = (lambda r: r.sums.copy_(tapenv.dist.combine_partial_sums(r.sums))) if reducer is not None else None   # in-order here: the host reads the totals

    def e2e_run(k):
        last = None
        for i in range(k):
            if pipe.inflight == pipe.depth:
                last = pipe.result()
            pipe.submit(hb, after_episode=after)
        while pipe.inflight:
            last = pipe.result()
        return last

    # warm-up of the host->device path itself: on a fresh box the first ~100 ms of pinned copies run at half rate (PCIe link /
    # IOMMU state); without this the same command measured 8.7e7 on its first run and 1.8e8 on its second (r01x)
    big = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
    big_d = torch.empty_like(big, device=dev)
    for _ in range(24):
        big_d.copy_(big, non_blocking=True)
    torch.cuda.synchronize(dev)
    e2e_run(max(args.warmup, 3) + 16)
    barrier()
    t0 = time.perf_counter()
    rw_pin, sums_pin = e2e_run(args.steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * B * steps_per_episode * args.steps / float(te.item())
    assert np.array_equal(rw_pin.numpy(), reward_ref.cpu().numpy())

    # ---- e2e, packed upload: the same episodes fed from the loader's compact host format (u8 static, bit-row dynamic:
    # tapenv.pack_inputs / PACKDataset.packed()); the fp32 tensors are produced on the device by tapenv_reset_packed ----
    e2e_packed = None
    if Wn == 1 and strat != "LB":
        su8, bits = tapenv.pack_inputs(static_h[0], dynamic_h[0])
        pipe_p = tapenv.HostPipeline(env, n, depth=4, use_graph=not args.no_graph, windows=1, exchange=exchange, packed=True)
        hbp = pipe_p.new_host_batch()
        hbp.static.copy_(torch.from_numpy(su8)); hbp.dynamic.copy_(torch.from_numpy(bits)); hbp.ptr.copy_(pq_pin)

        def e2e_packed_run(k):
            last = None
            for i in range(k):
                if pipe_p.inflight == pipe_p.depth:
                    last = pipe_p.result()
                pipe_p.submit(hbp, after_episode=after)
            while pipe_p.inflight:
                last = pipe_p.result()
            return last

        e2e_packed_run(max(args.warmup, 3) + 16)
        barrier()
        t0 = time.perf_counter()
        rwp, _ = e2e_packed_run(args.steps)
        barrier()
        tp = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        assert np.array_equal(rwp.numpy(), reward_ref.cpu().numpy())
        e2e_packed = {"value": world * B * steps_per_episode * args.steps / float(tp.item()), "unit": UNIT,
                      "h2d_bytes_per_step": int(pipe_p.h2d_bytes), "d2h_bytes_per_step": int(pipe_p.d2h_bytes),
                      "ms_per_step": 1e3 * float(tp.item()) / args.steps,
                      "api": "tapenv.HostPipeline(packed=True): u8 static + bit-row dynamic + ptr_seq from pinned host memory, "
                             "expanded to the fp32 tensors on the device (tapenv_reset_packed)"}
    # plain pinned H2D bandwidth of this box, for reading the e2e numbers (the fp32 path is PCIe-bound)
    big_d.copy_(big, non_blocking=True)
    torch.cuda.synchronize(dev)
    eh0, eh1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eh0.record()
    for _ in range(4):
        big_d.copy_(big, non_blocking=True)
    eh1.record()
    torch.cuda.synchronize(dev)
    h2d_gbs = 4 * big.numel() / (eh0.elapsed_time(eh1) * 1e-3) / 1e9
    del big, big_d
    clocks = sampler.result()

    # ---- extra: the whole-episode kernel (K7, tapenv_episode): one launch per episode, no intermediate tensors ----
    k7 = None
    if Wn == 1:
        for _ in range(3):
            rk = env.episode(st0[0], dyn0[0], ptr_seq0[0])[0]
        assert torch.equal(rk, reward_ref)
        torch.cuda.synchronize(dev)
        e0.record()
        for i in range(20):
            r = runners[i % RING]
            env.episode(r.static[0], r.dynamic[0], r.ptr_seq[0])
        e1.record()
        torch.cuda.synchronize(dev)
        k7 = {"value": B * n * 20 / (e0.elapsed_time(e1) * 1e-3), "unit": UNIT, "ms_per_episode": e0.elapsed_time(e1) / 20,
              "note": "tapenv_episode: reset + n steps + reward in ONE launch per batch (per GPU), inputs in HBM, eager launches"}

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only) ------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import oracle
        threads = os.cpu_count() or 1
        ptr_h = ptr_seq0.cpu().numpy()
        kw = dict(nthreads=threads, want=("reward",), capacity=Wn * n)
        oracle.episode_batch(static_h, dynamic_h, ptr_h, size, rt, hm, strat, **kw)
        reps, t0 = 0, time.perf_counter()
        while True:
            o = oracle.episode_batch(static_h, dynamic_h, ptr_h, size, rt, hm, strat, **kw)
            assert o["status"] == 0
            reps += 1
            el = time.perf_counter() - t0
            if el >= args.cpu_seconds or reps >= 10000:
                break
        rate = reps * B * steps_per_episode / el
        parity = bool(np.array_equal(o["reward"], reward_ref.cpu().numpy()))
        # the reference itself is single-threaded (trainer.py:155, no multiprocessing): the same port on ONE thread, short sample
        kw1 = dict(kw, nthreads=1)
        Bs = min(B, 512)
        r1, t1 = 0, time.perf_counter()
        while time.perf_counter() - t1 < min(3.0, args.cpu_seconds):
            oracle.episode_batch(static_h[:, :Bs], dynamic_h[:, :Bs], ptr_h[:, :, :Bs], size, rt, hm, strat, **kw1)
            r1 += 1
        rate1 = r1 * Bs * steps_per_episode / (time.perf_counter() - t1)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "value_one_thread": rate1,
               "sample": "%d episodes x %d envs x %d steps = %.1f s of oracle/tap_oracle.c on %d pthreads" % (reps, B, steps_per_episode, el, threads),
               "reward_parity_vs_gpu": parity,
               "python_reference_1core": "4.9e3 env-steps/s (unmodified tools.py path, build container, BASELINE.md section 2)"}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": span_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32 state + f64 score, f32 tensors", "data": "synthetic",
            "config": {"workload": desc, "batch_per_gpu": B, "blocks": steps_per_episode, "env_steps_per_step": world * B * steps_per_episode,
                       "step_definition": "one episode = reset + %d fused decode-step launches%s (the last also emits calc_ratio) + reward sums over the batch"
                                          % (steps_per_episode, " (%d windows of %d, masks re-initialised per window)" % (Wn, n) if Wn > 1 else ""),
                       "l2": "inputs rotate over a ring of %d distinct batches (%.0f MB incl. ping-pong outputs) > 126 MB L2" % (RING, RING * 3 * per_set / 1e6),
                       "cuda_graph": not args.no_graph, "reward_reduction": reduction,
                       "inputs": "reference RAND/PPSG generator fixtures (tests/golden), pool of %d tiled" % pool,
                       "policy": "recorded ptr ~ multinomial(current_mask), seed 1234+rank"},
            "gpu_launches": launches,
            "device_ms_sum_per_step": dev_ms / args.steps, "wall_ms_per_step": 1e3 * t_wall / args.steps,
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": int(pipe.h2d_bytes), "d2h_bytes_per_step": int(pipe.d2h_bytes), "ms_per_step": 1e3 * float(te.item()) / args.steps,
                    "api": "tapenv.HostPipeline.submit/result: one contiguous pinned host batch (fp32 static + dynamic, int64 ptr_seq) -> one H2D copy per episode, "
                           "4-deep pipeline, BatchedContainers.reset/step(+reward)/reward_sums graph, rewards + sums to pinned host memory",
                    "h2d_gbs_achieved": pipe.h2d_bytes * args.steps / float(te.item()) / 1e9, "h2d_gbs_box": h2d_gbs},
            "e2e_packed": e2e_packed,
            "episode_kernel": k7,
            "roofline": {"bound": "hbm", "kernel": "step_kernel (fused update_dynamic+update_mask+add_new_block)",
                         "achieved": achieved, "peak": peak,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "B200_PROFILING.md fallback (of fallback)",
                         "unit": "GB/s", "frac": achieved / peak, "traffic": TRAFFIC.get(args.workload),
                         "algorithmic_bytes_per_env_step": bytes_step, "bytes_per_launch": B * bytes_step,
                         "launch_us": step_us, "launches_timed": cnt,
                         "torch_copy_of_dynamic_us": copy_us,
                         "note": "torch_copy_of_dynamic_us: torch's copy kernel on the dynamic tensor alone (%.0f%% of the step's bytes), "
                                 "same slots and timing method -- the practical floor for a launch of this size" % (100.0 * 2 * dynamic_h[0].nbytes / (B * bytes_step))},
            "clocks": clocks,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="environments per GPU (default: the workload's)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--nccl-reduce", action="store_true", help="multi-GPU: reduce the reward statistics with NCCL instead of the fused peer-memory exchange")
    ap.add_argument("--no-reduce", action="store_true", help="multi-GPU diagnostic: skip the reward-statistics reduction altogether")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    args = ap.parse_args()
    if args.workload in ROLLING:
        return run_rolling_reference(args) if args.impl == "reference" else run_rolling(args)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
