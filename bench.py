#!/usr/bin/env python
"""bench.py -- env-steps/sec of the TAP packing-environment step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c4|c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

One bench "step" = ONE EPISODE of the hot path over one batch: K0 reset + n fused decode-step launches
(update_dynamic + update_mask + add_new_block for all B environments, the last one also emits calc_ratio) + the reward
statistics.  It advances B*n env-steps; reset and reward are inside the timed region but are not counted as env-steps
(SURVEY.md section 8d).  The pointer of every decode step is a recorded policy -- the floor(u*count)-th accessible
candidate of the CURRENT mask, u from np.random.RandomState(1234 + rank) -- computed identically in both arms from their own
masks (`policy_pick`), so the GPU arm and the CPU reference arm walk the same trajectories.

Prints ONE JSON line (rank 0).  The default line (workload c2 = BASELINE configs[1]) also carries a `configs` block with
C3 / C4 (B = 8192/N) / C5 (B = 65536/N) measured in the same process and, at N=1, a `model_in_loop` block (the unmodified
reference network driving the environment).  Keys: DESIGN.md section "Measurement".
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
    # name: (fixture, container_size, reward_type, heightmap_type, packing_strategy, description)
    "c2": ("rand2d_n10.npz", [5, 50], "C+P+S-lb-soft", "diff", "LB_GREEDY",
           "2D RAND nodes=10 width=5 LB_GREEDY C+P+S-lb-soft batch=4096 per GPU (BASELINE configs[1])"),
    "c3": ("rand3d_n10.npz", [5, 5, 50], "C+P+S-lb-soft", "diff", "LB_GREEDY",
           "3D RAND nodes=10 width=5 LB_GREEDY C+P+S-lb-soft batch=4096 per GPU (BASELINE configs[2])"),
    # container height 100 (not the trainer default 50): under a RANDOM policy a few 20-block stacks pass 50, where the
    # reference raises IndexError; placements are identical for any height that is not reached
    "c4": ("ppsg2d_n20.npz", [7, 100], "C+P+S-mcs-hard", "diff", "MACS",
           "2D PPSG nodes=20 width=7 height=100 MACS C+P+S-mcs-hard batch=8192 sharded over the GPUs (BASELINE configs[3])"),
    # rolling inference (rolling.py:575-640): ONE container [5,5,250] takes 50 blocks while the network window stays at 10;
    # the window is rebuilt before every decode step by generate.InitialContainer -> tapenv_rolling_step (see measure_rolling)
    "c5": ("rolling3d_t50.npz", [5, 5, 250], "C+P+S-lb-soft", "diff", "LB_GREEDY",
           "3D rolling total=50 window=10 width=5 height=250 LB_GREEDY batch=65536 sharded over the GPUs (BASELINE configs[4])"),
}
# the voxel-state strategies (no BASELINE configuration uses them): C2 / C3 inputs, other packing strategy
WORKLOADS["c2lb"] = ("rand2d_n10.npz", [5, 50], "C+P+S-lb-soft", "diff", "LB",
                     "2D RAND nodes=10 width=5 LB (tools.py:1602-1754) C+P+S-lb-soft batch=4096 per GPU")
WORKLOADS["c3lb"] = ("rand3d_n10.npz", [5, 5, 50], "C+P+S-lb-soft", "diff", "LB",
                     "3D RAND nodes=10 width=5 LB (tools.py:1756-1914) C+P+S-lb-soft batch=4096 per GPU")
WORKLOADS["c3macs"] = ("rand3d_n10.npz", [5, 5, 50], "C+P+S-mcs-soft", "diff", "MACS",
                       "3D RAND nodes=10 width=5 MACS (tools.py:2751-3165) C+P+S-mcs-soft batch=4096 per GPU")
ROLLING = {"c5": (50, 10)}          # workload -> (total_blocks_num, network window)
TOTAL_BATCH = {"c4": 8192, "c5": 65536}     # BASELINE's global batch, sharded over the GPUs (strong scaling)
PER_GPU_BATCH = {"c2": 4096, "c3": 4096, "c2lb": 4096, "c3lb": 4096, "c3macs": 4096}    # single-GPU batch, kept per GPU (weak scaling)


def default_batch(name, world):
    return PER_GPU_BATCH[name] if name in PER_GPU_BATCH else max(TOTAL_BATCH[name] // world, 1)


def scaling_of(name, batch_arg):
    return "weak" if (name in PER_GPU_BATCH or batch_arg) else "strong"


def algorithmic_bytes_per_env_step(n, R, dim, W, L, macs):
    """minimum HBM traffic of one env-step under the reference's tensor contract."""
    S = n * R
    cells = W if dim == 2 else W * L
    dec_dyn = (W - 1) if dim == 2 else 2 * W * L
    hist = (16 * n + 16) if macs else 0
    return (2 * (3 * n * S * 4) + 3 * (S * 4) + 8 + 4 * (1 + dim) + 2 * 4 * cells + 2 * 16 + hist + 4 * dec_dyn + 4 * dim)


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


def measured_traffic(name):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the workload's dominant kernel from this round's
    `ncu --set full` capture: profiles/traffic.json = {workload: {"bytes": .., "batch": .., "source": "<summary csv>"}}, written
    by scripts/ncu_summary.py --traffic from the captures of scripts/gpu_round.sh.  None when no capture is on file."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[name]
        return float(t["bytes"]), "%s (B=%s)" % (t["source"], t.get("batch"))
    except Exception:
        return None, None


def policy_pick(cur_mask, u):
    """cur_mask [B,S] (0/1), u [B] uniform in [0,1) -> int64 [B]: the floor(u*count)-th accessible candidate.
    (Same function as oracle/ref_worker.py:policy_pick, which the reference arm uses -- kept here so that the GPU arm does not
    import anything under oracle/.)"""
    m = np.asarray(cur_mask) > 0
    cnt = m.sum(1)
    k = np.minimum((np.asarray(u) * cnt).astype(np.int64), np.maximum(cnt - 1, 0))
    return (np.cumsum(m, 1) > k[:, None]).argmax(1).astype(np.int64)


def policy_uniforms(steps, B, rank):
    return np.random.RandomState(1234 + rank).random_sample((steps, B))


def load_workload(name, B, rank):
    """-> static [B,1+dim,S], dynamic [B,3n,S] float32 (the fixture pool tiled; ranks start at different offsets)."""
    from tests.golden_io import load_inputs
    fixture, size, rt, hm, strat, desc = WORKLOADS[name]
    static, dynamic = load_inputs(fixture)
    pool = static.shape[0]
    idx = (np.arange(B) + rank * 977) % pool
    return np.ascontiguousarray(static[idx]), np.ascontiguousarray(dynamic[idx]), pool


def load_rolling_workload(name, B, rank):
    from tests.golden_io import load_rolling
    T, n = ROLLING[name]
    z = load_rolling(WORKLOADS[name][0])
    pool = z["adj"].shape[0]
    idx = (np.arange(B) + rank * 977) % pool
    return z["adj"][idx], np.ascontiguousarray(z["blocks"][idx]), pool, T, n, z["dim"]


def ring_size(per_set_bytes):
    """Distinct input sets the timed episodes rotate over so that every episode reads its inputs from HBM, not from a warm
    L2: RING * (inputs + ping-pong outputs) >> 126 MB."""
    return max(1 if per_set_bytes > 300e6 else 2, int(np.ceil(400e6 / (3 * per_set_bytes))))


def bench_config(name, B, world, args):
    """The `config` object of the JSON line -- a function of (workload, batch, world size, flags) only, so that both arms
    of the bench print the SAME object."""
    fixture, size, rt, hm, strat, desc = WORKLOADS[name]
    dim = len(size)
    R = 2 if dim == 2 else 6
    if name in ROLLING:
        T, n = ROLLING[name]
        steps = T
        S = n * R
        per_episode = 2 * B * ((1 + dim) * S + 3 * n * S + 2 * S) * 4
        step_def = ("one episode = clear + window reset + first window + %d fused rolling steps (place + remove_block + convert_to_input) + "
                    "%d fused decode steps in the last window (the last also emits calc_ratio) + reward sums over the batch" % (T - n, n))
        l2 = "%.0f MB of ping-pong window tensors per episode > 126 MB L2; 2 instance sets alternate" % (per_episode / 1e6)
    else:
        n = 20 if name == "c4" else 10
        steps = n
        S = n * R
        per_set = B * ((1 + dim) * S + 3 * n * S) * 4
        ring = ring_size(per_set)
        step_def = "one episode = reset + %d fused decode-step launches (the last also emits calc_ratio) + reward sums over the batch" % n
        l2 = "inputs rotate over a ring of %d distinct batches (%.0f MB incl. ping-pong outputs) > 126 MB L2" % (ring, ring * 3 * per_set / 1e6)
    if world == 1:
        red = "none (single GPU)"
    elif args.no_reduce:
        red = "none (diagnostic run: --no-reduce)"
    elif args.nccl_reduce:
        red = "NCCL all_gather of the f64 triples on a side stream"
    else:
        red = "one-shot exchange of the f64 triples over NVLink peer memory (tapenv_reward_sums) on a side stream behind the last decode step"
    return {"workload": desc, "batch_per_gpu": B, "blocks": steps, "env_steps_per_step": world * B * steps,
            "step_definition": step_def, "l2": l2, "cuda_graph": not args.no_graph, "reward_reduction": red,
            "inputs": "reference RAND/PPSG/rolling generator fixtures (tests/golden), pool tiled to the batch",
            "policy": "recorded: floor(u*count)-th accessible candidate of the current mask, u ~ RandomState(1234+rank)"}


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
# CPU reference arm: the UNMODIFIED Python reference (oracle/_ref via oracle/ref_worker.py) on the host cores;
# the C port (oracle/tap_oracle.c) as the second figure, and as the fallback when the reference is not staged
# ----------------------------------------------------------------------------------------------
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def reference_staged():
    from oracle import refshim
    return refshim.available()


class PythonReferenceArm(object):
    """`workers` processes x `per_worker` environments of the workload's batch, each running the reference's own
    pack.update_* + tools.Container path (oracle/ref_worker.py)."""

    def __init__(self, name, static, dynamic, u, workers, per_worker):
        from oracle import ref_worker
        fixture, size, rt, hm, strat, desc = WORKLOADS[name]
        self.workers, self.per_worker = workers, per_worker
        self.steps = u.shape[0]
        shards, self.ranges = ref_worker.make_shards(static, dynamic, u, size, rt, hm, strat, workers, per_worker)
        self.pool = ref_worker.ReferencePool(workers)
        self.pool.load(shards)
        self.source = self.pool.source

    def episode(self, trace=False, only=None):
        return self.pool.run(trace=trace, only=only)

    def env_steps(self, nworkers=None):
        return (self.workers if nworkers is None else nworkers) * self.per_worker * self.steps

    def close(self):
        self.pool.close()


def port_rate(name, static, dynamic, ptr_seq, threads, seconds, max_reps=10000):
    """The C restatement (oracle/tap_oracle.c) on `threads` pthreads -> (env-steps/s, reps, elapsed, last result)."""
    from oracle import oracle
    fixture, size, rt, hm, strat, desc = WORKLOADS[name]
    B, n = static.shape[0], ptr_seq.shape[0]
    kw = dict(nthreads=threads, want=("reward",), capacity=n)
    st, dy, pq = static[None], dynamic[None], ptr_seq[None]
    oracle.episode_batch(st, dy, pq, size, rt, hm, strat, **kw)
    reps, t0 = 0, time.perf_counter()
    while True:
        o = oracle.episode_batch(st, dy, pq, size, rt, hm, strat, **kw)
        assert o["status"] == 0
        reps += 1
        el = time.perf_counter() - t0
        if el >= seconds or reps >= max_reps:
            break
    return reps * B * n / el, reps, el, o


def oracle_policy_ptrs(name, static, dynamic, u):
    """The recorded policy evaluated on the oracle's masks (reference arm / checker side only) -> ptr_seq [n,B]."""
    from oracle import oracle
    size = WORKLOADS[name][1]
    dim = len(size)
    R = 2 if dim == 2 else 6
    B, _, S = static.shape
    n = S // R
    mask = np.ones((B, S), np.float32)
    cur = oracle.initial_mask(dynamic, n, R)
    dyn, seq = dynamic, []
    for t in range(n):
        ptr = policy_pick(cur, u[t])
        dyn = oracle.update_dynamic(dyn, static, ptr)
        cur, mask = oracle.update_mask(mask, dyn, static, ptr)
        seq.append(ptr)
    return np.stack(seq)


def rolling_policy_ptrs(adj, blocks, T, n, dim, u):
    """The recorded policy on the oracle's rolling window (reference arm / checker side only) -> ptr_seq [T,B]."""
    from oracle import oracle
    B = adj.shape[0]
    R = 2 if dim == 2 else 6
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
                p = int(policy_pick(cur, u[t, b:b + 1])[0])
                ptr_seq[t, b] = p
                t += 1
                if last:
                    pa = np.array([p], np.int64)
                    dyn = oracle.update_dynamic(dyn, static[None], pa)
                    cur, mask = oracle.update_mask(mask, dyn, static[None], pa)
            ic.remove_block(ic.sub_graph_nodes[p % n])
    return ptr_seq


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path on this box's host cores, on the GPU arm's
    config.  Each bench step = one episode over a bounded sample of the batch (per_worker environments on every core)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return 0
    name = args.workload
    B = args.batch or default_batch(name, world)
    cfg = bench_config(name, B, world, args)
    cores = host_cores()
    base = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": scaling_of(name, args.batch), "vs_baseline": None,
            "dtype": "int64 state + f64 score (NumPy), f32 tensors (torch CPU)", "data": "synthetic", "config": cfg}
    if name in ROLLING:
        return run_rolling_reference_port(args, base, B, cores)
    static, dynamic, pool = load_workload(name, B, 0)
    n = cfg["blocks"]
    u = policy_uniforms(n, B, 0)
    if reference_staged() and not args.port:
        per = max(1, min(args.ref_envs_per_core, B // max(cores, 1) if B >= cores else 1))
        arm = PythonReferenceArm(name, static, dynamic, u, cores, per)
        try:
            for _ in range(max(args.warmup, 1)):
                res = arm.episode()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                res = arm.episode()
            el = time.perf_counter() - t0
            value = args.steps * arm.env_steps() / el
            # the faithful single-process figure (the reference has no multiprocessing, trainer.py:155): worker 0 alone
            t1 = time.perf_counter()
            one = arm.episode(only=[0])
            one_rate = arm.env_steps(1) / (time.perf_counter() - t1)
            source = arm.source
        finally:
            arm.close()
        line = dict(base)
        line.update({
            "value": value, "ms_per_step": 1e3 * el / args.steps,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "cpu_model": cpu_model(),
                             "value_one_core": one_rate,
                             "sample": "%d episodes x %d of the %d environments (%d worker processes x %d) x %d steps of the unmodified "
                                       "pack.update_dynamic + pack.update_mask + tools.Container.add_new_block + calc_ratio (%s)"
                                       % (args.steps, cores * per, B, cores, per, n, source)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "reward_mean": float(np.mean([r["reward"].mean() for r in res])),
        })
        # second figure: the C port of the same path on all host threads (what round 1 timed)
        try:
            pq = oracle_policy_ptrs(name, static, dynamic, u)
            prate, preps, pel, _ = port_rate(name, static, dynamic, pq, cores, min(3.0, args.cpu_seconds))
            line["cpu_baseline"]["port"] = {"value": prate, "cores": cores, "kind": "port",
                                            "sample": "%d episodes x %d envs, oracle/tap_oracle.c on %d pthreads" % (preps, B, cores)}
        except Exception as e:                                # the port is a courtesy figure here
            line["cpu_baseline"]["port"] = {"error": repr(e)}
        print(json.dumps(line))
        return 0
    # fallback: the reference tree is not staged (oracle/stage_ref.py) -> the C port
    pq = oracle_policy_ptrs(name, static, dynamic, u)
    from oracle import oracle
    fixture, size, rt, hm, strat, desc = WORKLOADS[name]
    kw = dict(nthreads=cores, want=("reward",), capacity=n)
    for _ in range(max(args.warmup, 1)):
        oracle.episode_batch(static[None], dynamic[None], pq[None], size, rt, hm, strat, **kw)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o = oracle.episode_batch(static[None], dynamic[None], pq[None], size, rt, hm, strat, **kw)
    el = time.perf_counter() - t0
    value = args.steps * B * n / el
    line = dict(base)
    line.update({"value": value, "ms_per_step": 1e3 * el / args.steps,
                 "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "cpu_model": cpu_model(),
                                  "sample": "%d episodes x %d envs x %d steps, oracle/tap_oracle.c on %d pthreads (the Python reference "
                                            "is not staged under oracle/_ref)" % (args.steps, B, n, cores)},
                 "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                 "reward_mean": float(o["reward"].mean())})
    print(json.dumps(line))
    return 0


def run_rolling_reference_port(args, base, B, cores):
    """C5: rolling.validate's loop.  The Python reference runs it at batch 1 with networkx graph copies per step; the arm
    timed here is the C restatement (oracle/win_oracle.c + tap_oracle.c), kind "port", on a bounded sample."""
    from oracle import oracle
    name = args.workload
    Bs = min(B, 1024)
    adj, blocks, pool, T, n, dim = load_rolling_workload(name, Bs, 0)
    fixture, size, rt, hm, strat, desc = WORKLOADS[name]
    u = policy_uniforms(T, B, 0)[:, :Bs]
    ptr_seq = rolling_policy_ptrs(adj, blocks, T, n, dim, u)
    kw = dict(nthreads=cores)
    for _ in range(max(args.warmup, 1)):
        oracle.rolling_batch(adj, blocks, ptr_seq, size, n, rt, hm, strat, **kw)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o = oracle.rolling_batch(adj, blocks, ptr_seq, size, n, rt, hm, strat, **kw)
    el = time.perf_counter() - t0
    value = args.steps * Bs * T / el
    line = dict(base)
    line.update({"value": value, "ms_per_step": 1e3 * el / args.steps,
                 "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "cpu_model": cpu_model(),
                                  "sample": "%d episodes x %d of the %d instances x %d steps, oracle/win_oracle.c + tap_oracle.c on %d pthreads"
                                            % (args.steps, Bs, B, T, cores)},
                 "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                 "reward_mean": float(o["reward"].mean())})
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
class Ctx(object):
    """One process per GPU: rank plumbing, barrier and max-over-ranks."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- the tapenv path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.affinity = bind_cores(self.local, self.world)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        self.exchange = None
        self.peaks = {}
        try:
            self.peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def allmax(self, v):
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(self, v):
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def peak(self):
        return float(self.peaks.get("hbm_gbs", 6650.0)), ("MEASURED_PEAKS.json hbm_gbs (of measured)" if self.peaks else "B200_PROFILING.md fallback (of fallback)")

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def bind_cores(local, world):
    """Give every rank its own slice of the host cores it is allowed to run on (launcher thread, pinned-buffer first touch,
    copy-engine doorbells stay on cores no other rank uses).  On the 8-GPU boxes of this pool all GPUs report the same CPU
    affinity and ONE NUMA node (SCALE_r01 topology: 0-31 / node 0), so there is no NUMA-local placement to pick -- the slice
    only removes the cross-rank migration noise."""
    try:
        cpus = sorted(os.sched_getaffinity(0))
        if world <= 1 or len(cpus) < 2 * world:
            return None
        per = len(cpus) // world
        mine = cpus[local * per:(local + 1) * per]
        os.sched_setaffinity(0, mine)
        return "%d-%d" % (mine[0], mine[-1])
    except Exception:
        return None


def setup_exchange(ctx, args):
    """Multi-GPU: the end-of-episode reduction of the reward statistics feeding the critic baseline (trainer.py:216-225).
    Preferred: the one-shot exchange over NVLink peer memory (tapenv_reward_sums with a PeerExchange) on a side stream;
    otherwise an NCCL all-gather of the f64 triples on a side stream."""
    import tapenv
    exchange, reducer, how = None, None, "none (single GPU)"
    if ctx.world > 1 and args.no_reduce:
        how = "none (diagnostic run: --no-reduce)"
    elif ctx.world > 1 and not args.nccl_reduce:
        try:
            exchange = ctx.exchange if ctx.exchange is not None else tapenv.dist.PeerExchange(ctx.dev)
            ctx.exchange = exchange
            how = "peer-memory exchange (tapenv_reward_sums + PeerExchange), %s" % ("in the CUDA graph" if args.inline_exchange else "on a side stream")
        except Exception as e:                          # no symmetric memory / P2P on this box
            exchange = None
            how = "PeerExchange unavailable (%s); " % type(e).__name__
    if ctx.world > 1 and exchange is None and not args.no_reduce:
        reducer = tapenv.dist.RewardReducer(ctx.dev)
        how = (how if how.startswith("PeerExchange") else "") + "NCCL all_gather of the f64 triples on a side stream"
    return exchange, reducer, how


def check_exchange(ctx, runner, B, count_per_env=1):
    """The collective must be checkable: after an episode the fused totals equal the rank-order sum of the
    all-gathered per-rank triples, carry world*B environments, are identical on every rank and equal the NCCL path."""
    import tapenv
    torch, dist = ctx.torch, ctx.dist
    runner.tail.wait_total()
    torch.cuda.synchronize(ctx.dev)
    total, sums = runner.tail.total.clone(), runner.tail.sums.clone()
    gathered = torch.empty(ctx.world, 3, dtype=torch.float64, device=ctx.dev)
    dist.all_gather_into_tensor(gathered, sums.reshape(1, 3).contiguous())
    want = gathered[0].clone()
    for r in range(1, ctx.world):
        want += gathered[r]
    ok = bool(torch.equal(total, want)) and float(total[2]) == float(ctx.world * B)
    nccl = tapenv.dist.combine_partial_sums(sums)
    ok = ok and bool(torch.equal(nccl, total))
    alls = torch.empty(ctx.world, 3, dtype=torch.float64, device=ctx.dev)
    dist.all_gather_into_tensor(alls, total.reshape(1, 3).contiguous())
    ok = ok and all(bool(torch.equal(alls[r], alls[0])) for r in range(ctx.world))
    ctx.exchange.check()
    assert ok, "reward exchange mismatch on rank %d: fused %s vs gathered %s" % (ctx.rank, total.tolist(), want.tolist())
    return True


def warm_h2d(ctx):
    """Warm-up of the host->device path itself: on a fresh box the first ~100 ms of pinned copies run at half rate (PCIe
    link / IOMMU state); without this the same command measured 8.7e7 on its first run and 1.8e8 on its second (r01x)."""
    torch = ctx.torch
    big = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
    big_d = torch.empty_like(big, device=ctx.dev)
    for _ in range(24):
        big_d.copy_(big, non_blocking=True)
    torch.cuda.synchronize(ctx.dev)
    return big, big_d


def h2d_rate(ctx, big, big_d, reps=4):
    torch = ctx.torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    big_d.copy_(big, non_blocking=True)
    torch.cuda.synchronize(ctx.dev)
    e0.record()
    for _ in range(reps):
        big_d.copy_(big, non_blocking=True)
    e1.record()
    torch.cuda.synchronize(ctx.dev)
    return reps * big.numel() / (e0.elapsed_time(e1) * 1e-3) / 1e9


def h2d_ceiling(ctx, big, big_d):
    """Plain pinned H2D bandwidth of this box, for reading the e2e numbers (the fp32 path is PCIe-bound): every rank
    copying at the same time (what the e2e leg does) and every rank on its own (its link alone)."""
    ctx.barrier()
    together = h2d_rate(ctx, big, big_d)
    solo = None
    if ctx.world > 1:
        solo = 0.0
        for r in range(ctx.world):
            ctx.barrier()
            if r == ctx.rank:
                solo = h2d_rate(ctx, big, big_d)
        ctx.barrier()
    return together, solo


def measure_steps(ctx, args, name, B, full):
    """One non-rolling workload on this rank's GPU -> dict of measurements (rank 0 assembles the line)."""
    import tapenv
    torch = ctx.torch
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    fixture, size, rt, hm, strat, desc = WORKLOADS[name]
    static_h, dynamic_h, pool = load_workload(name, B, rank)
    dim = len(size)
    R = 2 if dim == 2 else 6
    S = static_h.shape[2]
    n = S // R
    macs = strat == "MACS" or "mcs" in rt
    bytes_step = algorithmic_bytes_per_env_step(n, R, dim, size[0], size[1] if dim == 3 else 1, macs)
    per_set = static_h.nbytes + dynamic_h.nbytes
    RING = ring_size(per_set)
    env = tapenv.BatchedContainers(size, n, rt, hm, packing_strategy=strat, batch_size=B, device=dev)

    # record the policy once (untimed), keeping the trajectory of the first `keep` environments for the CPU parity check
    u = policy_uniforms(n, B, rank)
    keep = min(B, args.ref_envs_per_core * host_cores()) if (full and rank == 0 and world == 1 and not args.no_cpu) else 0
    st0 = torch.from_numpy(static_h).to(dev)
    dyn0 = torch.from_numpy(dynamic_h).to(dev)
    cur, mask = env.reset(dyn0)
    dyn = dyn0
    ptrs, trace = [], dict(heightmap=[], cur_mask=[cur[:keep].cpu().numpy()], ptr=[])
    for t in range(n):
        ptr_h = policy_pick(cur.cpu().numpy(), u[t])
        ptr = torch.from_numpy(ptr_h).to(dev)
        dyn, cur, mask, _, _ = env.step(ptr, st0, dyn, mask)
        ptrs.append(ptr)
        if keep:
            trace["heightmap"].append(env.heightmap[:keep].reshape(keep, -1).cpu().numpy())
            trace["cur_mask"].append(cur[:keep].cpu().numpy())
            trace["ptr"].append(ptr_h[:keep])
    ptr_seq0 = torch.stack(ptrs)                              # [n,B]
    reward_ref = env.calc_ratio().clone()
    trace["positions"] = env.positions[:keep].cpu().numpy()
    env.check_flags()
    del dyn, cur, mask

    exchange, reducer, reduction = setup_exchange(ctx, args)
    runners = []
    for i in range(RING):
        roll = (i * 131) % B
        runners.append(tapenv.EpisodeRunner(env, torch.roll(st0, roll, 0).contiguous(), torch.roll(dyn0, roll, 0).contiguous(),
                                            torch.roll(ptr_seq0, roll, 1).contiguous(), use_graph=not args.no_graph, partial_sums=True,
                                            exchange=exchange, overlap_exchange=not args.inline_exchange))

    def episode(i):
        r = runners[i % RING]
        if reducer is not None and getattr(r, "nccl_reduced", None) is not None:
            torch.cuda.current_stream().wait_event(r.nccl_reduced)     # the slot's previous sums have been consumed
        r.run()
        if reducer is not None:
            r.tail.wait_total()                                        # the local sums come from the side stream
            r.nccl_total, r.nccl_reduced = reducer.reduce_async(r.sums)
        return r

    for i in range(max(args.warmup, 3)):
        episode(i)
    ctx.barrier()
    # parity spot check of the replay against the recorded pass (same kernels; the oracle / reference checks live in tests/,
    # smoke() and the cpu_baseline leg below)
    assert torch.equal(runners[0].run(), reward_ref)
    exchange_checked = None
    if exchange is not None:
        exchange_checked = check_exchange(ctx, runners[0], B)

    # ---- timed region 1: `value` -- K episodes, inputs resident in HBM --------------------------------
    # (the NVML handle is opened and the checks above are done BEFORE the last warm-up episodes, so that nothing but the
    # barrier + synchronize sits between warm-up and the K timed episodes: at K=20 the region is about a millisecond and
    # an idle gap of NVML-init length in front of it showed up as 5 % run-to-run spread)
    sampler = ClockSampler(ctx.local)
    for i in range(3):
        episode(i)
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall = time.perf_counter()
    e0.record()
    for k in range(args.steps):
        episode(k)
    e1.record()
    sampler.sample_until(e1)                                  # clocks while the GPU works through the timed episodes
    ctx.barrier()
    t_wall = time.perf_counter() - t_wall
    span_ms_max = ctx.allmax(e0.elapsed_time(e1))
    value = world * B * n * args.steps / (span_ms_max * 1e-3)
    launches = args.steps * runners[0].launches_per_episode
    if exchange is not None:
        exchange_checked = check_exchange(ctx, runners[(args.steps - 1) % RING], B) and exchange_checked
    # the contract's timed region is K episodes (about a millisecond at K=20): the same loop over >= 1000 episodes as a
    # steadiness check of that number (not the headline)
    sustained = None
    if full:
        ks = max(1000, 10 * args.steps)
        ctx.barrier()
        e0.record()
        for k in range(ks):
            episode(k)
        e1.record()
        sampler.sample_until(e1)                              # more clock / throttle samples under the same load
        ctx.barrier()
        sustained = {"episodes": ks, "value": world * B * n * ks / (ctx.allmax(e0.elapsed_time(e1)) * 1e-3), "unit": UNIT}
    out = {"name": name, "B": B, "n": n, "value": value, "ms_per_step": span_ms_max / args.steps, "launches": launches,
           "sustained": sustained,
           "wall_ms_per_step": 1e3 * t_wall / args.steps, "reduction": reduction, "exchange_checked": exchange_checked,
           "reward_parity_vs_recorded_pass": True, "pool": pool}

    # ---- timed region 2: roofline of the dominant kernel (the fused step), cold inputs ---------------
    # graph-replayed launches, each on a different ring slot, events on the launching stream
    # One replay = the n decode steps of an episode (container fill k = 0..n-1: the placement scans lengthen with k), step t
    # on ring slot t % RING -- consecutive launches never touch the same input set, and a slot is revisited only after
    # RING * (inputs + outputs) >> 126 MB of other traffic.
    nl = n
    nbuf = min(len(runners), n)
    out_bufs = [(torch.empty_like(dyn0), torch.empty(B, S, device=dev), torch.empty(B, S, device=dev),
                 torch.empty(B, dim, device=dev), torch.empty(B, env.enc_len, device=dev)) for _ in range(nbuf)]
    mask1 = torch.ones(B, S, device=dev)

    def roof_launches():
        for i in range(nl):
            r = runners[i % len(runners)]
            env.step(r.ptr_seq[0, i], r.static[0], r.dynamic[0], mask1, out=out_bufs[i % nbuf])

    def graph_timed(fn, reps=12):
        env.clear_container()
        fn()                                                  # warm-up outside capture
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()                            # replayed from a graph so that the host's per-call
        with torch.cuda.graph(g):                             # overhead is not what gets timed
            fn()
        tot, cnt = 0.0, 0
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for rep in range(reps):
            env.clear_container()                             # also disturbs L2 between replays
            torch.cuda.synchronize(dev)
            a.record(); g.replay(); b.record()
            torch.cuda.synchronize(dev)
            if rep >= 2:
                tot += a.elapsed_time(b); cnt += nl
        return 1e3 * tot / cnt, cnt

    step_us, cnt = graph_timed(roof_launches)
    achieved = B * bytes_step / (step_us * 1e-6) / 1e9

    # context for the fraction: the out-of-place copy of `dynamic` alone (the clone every update_dynamic must make,
    # pack.py:370) by torch's copy kernel, same cold ring slots, same graph-replay + event method
    def copy_launches():
        for i in range(nl):
            out_bufs[i % nbuf][0].copy_(runners[i % len(runners)].dynamic[0])

    copy_us, _ = graph_timed(copy_launches)
    peak, peak_src = ctx.peak()
    traffic, traffic_src = measured_traffic(name)
    out["roofline"] = {"bound": "hbm", "kernel": "step_kernel (fused update_dynamic+update_mask+add_new_block)",
                       "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                       "traffic": traffic, "traffic_source": traffic_src,
                       "algorithmic_bytes_per_env_step": bytes_step, "bytes_per_launch": B * bytes_step,
                       "launch_us": step_us, "launches_timed": cnt, "torch_copy_of_dynamic_us": copy_us,
                       "note": "torch_copy_of_dynamic_us: torch's copy kernel on the dynamic tensor alone (%.0f%% of the step's bytes), "
                               "same slots and timing method -- the practical floor for a launch of this size" % (100.0 * 2 * dynamic_h.nbytes / (B * bytes_step))}
    del out_bufs
    if not full:
        out["clocks"] = sampler.result()
        del runners
        return out

    # ---- timed region 3: e2e -- host buffers in, host rewards out, through the public Python API --------
    # tapenv.HostPipeline: per episode one H2D upload of (static, dynamic, ptr_seq) from pinned memory on a copy
    # stream (pipelined, overlapping the previous episodes' kernels), the episode, D2H of rewards + sums.
    pq_pin = ptr_seq0.cpu().pin_memory()
    pipe = tapenv.HostPipeline(env, n, depth=4, use_graph=not args.no_graph, windows=1, exchange=exchange)
    hb = pipe.new_host_batch()                                # ONE contiguous pinned batch (what a loader fills in place): one H2D copy per episode
    hb.static.copy_(torch.from_numpy(static_h).view_as(hb.static)); hb.dynamic.copy_(torch.from_numpy(dynamic_h).view_as(hb.dynamic))
    hb.ptr.copy_(pq_pin.view_as(hb.ptr))
    def nccl_after(r):                                        # NCCL fallback, in-order here: the host reads the totals
        r.tail.wait_total()
        with torch.cuda.stream(r.tail.stream):
            r.sums.copy_(tapenv.dist.combine_partial_sums(r.sums))
            r.tail.reduced.record(r.tail.stream)
    after = nccl_after if reducer is not None else None

    def drive(p, batch, k):
        last = None
        for i in range(k):
            if p.inflight == p.depth:
                last = p.result()
            p.submit(batch, after_episode=after)
        while p.inflight:
            last = p.result()
        return last

    big, big_d = warm_h2d(ctx)
    drive(pipe, hb, max(args.warmup, 3) + 16)
    ctx.barrier()
    t0 = time.perf_counter()
    rw_pin, sums_pin = drive(pipe, hb, args.steps)
    ctx.barrier()
    e2e_s = ctx.allmax(time.perf_counter() - t0)
    e2e_value = world * B * n * args.steps / e2e_s
    assert np.array_equal(rw_pin.numpy(), reward_ref.cpu().numpy())
    if exchange is not None:
        assert float(sums_pin[2]) == float(world * B)

    # ---- e2e, packed upload: the same episodes fed from the loader's compact host format (u8 static, bit-row dynamic:
    # tapenv.pack_inputs / PACKDataset.packed()); the fp32 tensors are produced on the device by tapenv_reset_packed ----
    e2e_packed = None
    if strat != "LB":
        su8, bits = tapenv.pack_inputs(static_h, dynamic_h)
        pipe_p = tapenv.HostPipeline(env, n, depth=4, use_graph=not args.no_graph, windows=1, exchange=exchange, packed=True)
        hbp = pipe_p.new_host_batch()
        hbp.static.copy_(torch.from_numpy(su8)); hbp.dynamic.copy_(torch.from_numpy(bits)); hbp.ptr.copy_(pq_pin.view_as(hbp.ptr))
        drive(pipe_p, hbp, max(args.warmup, 3) + 16)
        ctx.barrier()
        t0 = time.perf_counter()
        rwp, _ = drive(pipe_p, hbp, args.steps)
        ctx.barrier()
        tp = ctx.allmax(time.perf_counter() - t0)
        assert np.array_equal(rwp.numpy(), reward_ref.cpu().numpy())
        e2e_packed = {"value": world * B * n * args.steps / tp, "unit": UNIT,
                      "h2d_bytes_per_step": int(pipe_p.h2d_bytes), "d2h_bytes_per_step": int(pipe_p.d2h_bytes),
                      "ms_per_step": 1e3 * tp / args.steps,
                      "api": "tapenv.HostPipeline(packed=True): u8 static + bit-row dynamic + ptr_seq from pinned host memory, "
                             "expanded to the fp32 tensors on the device (tapenv_reset_packed)"}
        del pipe_p
    together, solo = h2d_ceiling(ctx, big, big_d)
    del big, big_d
    out["clocks"] = sampler.result()
    out["e2e"] = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(pipe.h2d_bytes), "d2h_bytes_per_step": int(pipe.d2h_bytes),
                  "ms_per_step": 1e3 * e2e_s / args.steps,
                  "api": "tapenv.HostPipeline.submit/result: one contiguous pinned host batch (fp32 static + dynamic, int64 ptr_seq) -> one H2D copy "
                         "per episode, 4-deep pipeline, BatchedContainers.reset/step(+reward)/reward_sums graph, rewards + sums to pinned host memory",
                  "h2d_gbs_achieved": pipe.h2d_bytes * args.steps / e2e_s / 1e9,
                  "h2d_gbs_box": together, "h2d_gbs_box_aggregate": ctx.allsum(together),
                  "h2d_gbs_link_alone": solo, "cores_bound": ctx.affinity,
                  "note": "h2d_gbs_box: plain pinned 64 MB copies with every rank copying at once (aggregate = the host's feed ceiling for "
                          "this many GPUs); h2d_gbs_link_alone: the same copy with the other ranks idle"}
    out["e2e_packed"] = e2e_packed
    del pipe

    # ---- extra: the whole-episode kernel (K7, tapenv_episode): one launch per episode, no intermediate tensors ----
    for _ in range(3):
        rk = env.episode(st0, dyn0, ptr_seq0)[0]
    assert torch.equal(rk, reward_ref)
    torch.cuda.synchronize(dev)
    e0.record()
    for i in range(20):
        r = runners[i % RING]
        env.episode(r.static[0], r.dynamic[0], r.ptr_seq[0])
    e1.record()
    torch.cuda.synchronize(dev)
    out["episode_kernel"] = {"value": B * n * 20 / (e0.elapsed_time(e1) * 1e-3), "unit": UNIT, "ms_per_episode": e0.elapsed_time(e1) / 20,
                             "note": "tapenv_episode: reset + n steps + reward in ONE launch per batch (per GPU), inputs in HBM, eager launches"}

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only) ------------------------------------
    if rank == 0 and world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_baseline_leg(args, name, static_h, dynamic_h, u, ptr_seq0.cpu().numpy(), reward_ref.cpu().numpy(), trace, keep)
    del runners
    return out


def cpu_baseline_leg(args, name, static_h, dynamic_h, u, ptr_gpu, reward_gpu, trace, keep):
    """The reference's CPU path beside the GPU numbers, same box, same inputs, same recorded policy, parity asserted:
    kind "reference" = the unmodified Python modules (oracle/_ref) in one process per host core; "port" = the C
    restatement (second figure, or the only one when the reference is not staged)."""
    B, n = static_h.shape[0], ptr_gpu.shape[0]
    cores = host_cores()
    out = None
    if reference_staged() and keep >= cores:
        per = keep // cores
        arm = PythonReferenceArm(name, static_h, dynamic_h, u, cores, per)
        try:
            first = arm.episode(trace=True)                   # also the warm-up
            ok = True
            for w, idx in enumerate(arm.ranges):              # per-step parity against the GPU pass (heightmap / mask / pointer / position bit-exact)
                r = first[w]
                ok = ok and np.array_equal(r["ptr"], np.stack(trace["ptr"])[:, idx]) and np.array_equal(r["heightmap"], np.stack(trace["heightmap"])[:, idx])
                ok = ok and np.array_equal(r["cur_mask"], np.stack(trace["cur_mask"])[:, idx]) and np.array_equal(r["positions"], trace["positions"][idx])
                ok = ok and float(np.abs(r["reward"].astype(np.float64) - reward_gpu[idx].astype(np.float64)).max()) <= 1e-6
            reps, t0 = 0, time.perf_counter()
            while True:
                arm.episode()
                reps += 1
                el = time.perf_counter() - t0
                if el >= args.cpu_seconds or reps >= 1000:
                    break
            rate = reps * arm.env_steps() / el
            t1 = time.perf_counter()
            arm.episode(only=[0])
            one = arm.env_steps(1) / (time.perf_counter() - t1)
            out = {"value": rate, "unit": UNIT, "cores": cores, "kind": "reference", "cpu_model": cpu_model(), "value_one_core": one,
                   "sample": "%d episodes x %d of the %d environments (%d worker processes x %d) x %d steps = %.1f s of the unmodified "
                             "pack.update_dynamic + pack.update_mask + tools.Container.add_new_block + calc_ratio (%s)"
                             % (reps, cores * per, B, cores, per, n, el, arm.source),
                   "parity_vs_gpu_every_step": bool(ok)}
            assert ok, "GPU path and the Python reference disagree"
        finally:
            arm.close()
    # the C port on all host threads, whole batch (round 1's figure; the fast large-batch checker)
    prate, preps, pel, o = port_rate(name, static_h, dynamic_h, ptr_gpu, cores, min(args.cpu_seconds, 5.0) if out else args.cpu_seconds)
    port = {"value": prate, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d episodes x %d envs x %d steps = %.1f s of oracle/tap_oracle.c on %d pthreads" % (preps, B, n, pel, cores),
            "reward_parity_vs_gpu": bool(np.array_equal(o["reward"], reward_gpu))}
    Bs = min(B, 512)
    r1, _, _, _ = port_rate(name, static_h[:Bs], dynamic_h[:Bs], ptr_gpu[:, :Bs], 1, 2.0)
    port["value_one_thread"] = r1
    if out is None:
        port["cpu_model"] = cpu_model()
        return port
    out["port"] = port
    return out


def measure_rolling(ctx, args, name, B, full):
    import tapenv
    torch = ctx.torch
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    fixture, size, rt, hm, strat, desc = WORKLOADS[name]
    adj, blocks_h, pool, T, n, dim = load_rolling_workload(name, B, rank)
    graphs_h = tapenv.pack_graphs(adj)
    bytes_step = rolling_bytes_per_env_step(T, n, dim, size[0], size[1] if dim == 3 else 1)
    env = tapenv.BatchedContainers(size, T, rt, hm, packing_strategy=strat, batch_size=B, device=dev, window=n)
    exchange, reducer, reduction = setup_exchange(ctx, args)

    # record the policy once (untimed) on the live windows
    u = policy_uniforms(T, B, rank)
    win0 = tapenv.BatchedInitialContainers(graphs_h, blocks_h, T, n, dim, device=dev)
    rec = tapenv.RollingRunner(env, win0)
    static, dynamic, cur = rec.begin()
    ptrs = []
    for t in range(T):
        ptr = torch.from_numpy(policy_pick(cur.cpu().numpy(), u[t])).to(dev)
        ptrs.append(ptr)
        static, dynamic, cur, _, _ = rec.step(ptr)
    ptr_seq0 = torch.stack(ptrs)
    reward_ref = env.calc_ratio().clone()
    env.check_flags(); win0.check_flags()
    del rec

    RING = 2                                               # one episode's ping-pong outputs alone exceed the 126 MB L2
    runners = []
    for i in range(RING):
        roll = (i * 131) % B
        win = tapenv.BatchedInitialContainers(torch.roll(win0.graphs, roll, 0), torch.roll(win0.blocks, roll, 0), T, n, dim, device=dev)
        runners.append(tapenv.RollingRunner(env, win, ptr_seq=torch.roll(ptr_seq0, roll, 1).contiguous(), use_graph=not args.no_graph,
                                            partial_sums=True, exchange=exchange, overlap_exchange=not args.inline_exchange))
    for i in range(max(args.warmup, 3)):
        runners[i % RING].run()
    ctx.barrier()
    assert torch.equal(runners[0].run(), reward_ref)
    exchange_checked = check_exchange(ctx, runners[0], B) if exchange is not None else None

    sampler = ClockSampler(ctx.local)
    for i in range(2):
        runners[i % RING].run()
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall = time.perf_counter()
    e0.record()
    for k in range(args.steps):
        runners[k % RING].run()
    e1.record()
    sampler.sample_until(e1)
    ctx.barrier()
    t_wall = time.perf_counter() - t_wall
    span_ms_max = ctx.allmax(e0.elapsed_time(e1))
    value = world * B * T * args.steps / (span_ms_max * 1e-3)
    if exchange is not None:
        exchange_checked = check_exchange(ctx, runners[(args.steps - 1) % RING], B) and exchange_checked
    out = {"name": name, "B": B, "n": T, "value": value, "ms_per_step": span_ms_max / args.steps,
           "launches": args.steps * runners[0].launches_per_episode, "wall_ms_per_step": 1e3 * t_wall / args.steps,
           "reduction": reduction, "exchange_checked": exchange_checked, "reward_parity_vs_recorded_pass": True, "pool": pool}

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
    peak, peak_src = ctx.peak()
    traffic, traffic_src = measured_traffic(name)
    out["roofline"] = {"bound": "hbm", "kernel": "window_kernel (fused add_new_block + remove_block + convert_to_input)",
                       "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                       "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_env_step": bytes_step,
                       "bytes_per_launch": B * bytes_step, "launch_us": step_us, "launches_timed": cnt}
    del rr, rg
    if not full:
        out["clocks"] = sampler.result()
        del runners
        return out

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

    big, big_d = warm_h2d(ctx)
    del big, big_d
    e2e_run(max(args.warmup, 3) + 16)
    ctx.barrier()
    t0 = time.perf_counter()
    rw_pin, sums_pin = e2e_run(args.steps)
    ctx.barrier()
    te = ctx.allmax(time.perf_counter() - t0)
    assert np.array_equal(rw_pin.numpy(), reward_ref.cpu().numpy())
    out["clocks"] = sampler.result()
    out["e2e"] = {"value": world * B * T * args.steps / te, "unit": UNIT, "h2d_bytes_per_step": int(pipe.h2d_bytes),
                  "d2h_bytes_per_step": int(pipe.d2h_bytes), "ms_per_step": 1e3 * te / args.steps,
                  "api": "tapenv.RollingHostPipeline.submit/result (graphs + blocks + pointers from pinned host memory), rewards to pinned host memory"}

    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import oracle
        threads = host_cores()
        Bs = min(B, 4096)
        ptr_h = ptr_seq0.cpu().numpy()[:, :Bs]
        oracle.rolling_batch(adj[:Bs], blocks_h[:Bs], ptr_h, size, n, rt, hm, strat, nthreads=threads)
        reps, t0 = 0, time.perf_counter()
        while True:
            o = oracle.rolling_batch(adj[:Bs], blocks_h[:Bs], ptr_h, size, n, rt, hm, strat, nthreads=threads)
            assert o["status"] == 0
            reps += 1
            el = time.perf_counter() - t0
            if el >= args.cpu_seconds or reps >= 10000:
                break
        out["cpu_baseline"] = {"value": reps * Bs * T / el, "unit": UNIT, "cores": threads, "kind": "port", "cpu_model": cpu_model(),
                               "sample": "%d episodes x %d of the %d instances x %d steps = %.1f s of oracle/win_oracle.c + tap_oracle.c on %d pthreads"
                                         % (reps, Bs, B, T, el, threads),
                               "reward_parity_vs_gpu": bool(np.array_equal(o["reward"], reward_ref.cpu().numpy()[:Bs]))}
    del runners, pipe
    return out


def model_in_loop(ctx, args):
    """The UNMODIFIED reference network (model.DRL + the shipped pretrained 2D actor, greedy decode) driving the environment at
    the C2 batch: (a) with the reference's own environment (bounded sample -- it is a per-environment Python loop), (b) after
    tapenv.install(pack, tools) -- the same model.py, environment on the GPU, (c) the same network modules feeding
    tapenv.DecodeLoop through tapenv.adapters.drl_actor_step (no host round trips, one CUDA graph per episode).
    env-steps/s = B*n / wall time of DRL.forward."""
    import tapenv
    from tests import ref_model
    torch = ctx.torch
    if not ref_model.available():
        return {"unavailable": "reference tree not staged (oracle/stage_ref.py)"}
    dev = ctx.dev
    B, n, dim = 4096, 10, 2
    static_h, dynamic_h, _ = load_workload("c2", B, 0)
    st, dy = torch.from_numpy(static_h).to(dev), torch.from_numpy(dynamic_h).to(dev)
    mods = ref_model.reference_modules()
    out = {"batch": B, "network": "model.DRL, pretrain_model/%s/actor.pt, eval() (greedy)" % ref_model.CHECKPOINTS[2]}

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(reps):
            res = fn()
        torch.cuda.synchronize(dev)
        return (time.perf_counter() - t0) / reps, res

    with torch.no_grad():
        Bs = 512
        actor = ref_model.make_actor(dim, True).eval()
        sec, (tour_ref, _, _, r_ref) = timed(lambda: ref_model.forward(actor, st[:Bs], dy[:Bs]), 1)
        out["reference_env"] = {"value": Bs * n / sec, "unit": UNIT, "ms_per_forward": 1e3 * sec,
                                "sample": "%d of the %d environments (tools.Container per environment on one host core, network on the GPU)" % (Bs, B)}
        tapenv.install(mods["pack"], mods["tools"])
        try:
            actor2 = ref_model.make_actor(dim, True).eval()
            sec, (tour, _, _, r) = timed(lambda: ref_model.forward(actor2, st, dy), 3)
        finally:
            tapenv.uninstall()
        out["install"] = {"value": B * n / sec, "unit": UNIT, "ms_per_forward": 1e3 * sec,
                          "tours_equal_reference_env": bool(torch.equal(tour[:Bs], tour_ref)),
                          "reward_max_abs_diff": float((r[:Bs] - r_ref).abs().max())}
        env = tapenv.BatchedContainers([5, 50], n, "C+P+S-lb-soft", "diff", batch_size=B, device=dev)
        for graph in (False, True):
            loop = tapenv.DecodeLoop(env, tapenv.adapters.drl_actor_step(actor), greedy=True, use_graph=graph)
            try:
                sec, (tour_d, _, r_d) = timed(lambda: loop.run(st, dy), 10)
                out["decode_loop_graph" if graph else "decode_loop"] = {
                    "value": B * n / sec, "unit": UNIT, "ms_per_forward": 1e3 * sec,
                    "tours_equal_install": bool(torch.equal(tour_d, tour)), "reward_max_abs_diff": float((-r_d - r).abs().max())}
            except Exception as e:                          # e.g. an actor that cannot be captured
                out["decode_loop_graph" if graph else "decode_loop"] = {"error": repr(e)[:200]}
    return out


def run_ours(args):
    ctx = Ctx()
    name = args.workload
    B = args.batch or default_batch(name, ctx.world)
    measure = measure_rolling if name in ROLLING else measure_steps
    m = measure(ctx, args, name, B, True)
    line = None
    if ctx.rank == 0:
        line = {"metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": ctx.world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": scaling_of(name, args.batch), "vs_baseline": None,
                "dtype": "int32 state + f64 score, f32 tensors" + (", u64 graph masks" if name in ROLLING else ""), "data": "synthetic",
                "config": bench_config(name, B, ctx.world, args), "gpu_launches": m["launches"], "wall_ms_per_step": m["wall_ms_per_step"],
                "reward_reduction": m["reduction"], "exchange_checked": m["exchange_checked"], "value_sustained": m.get("sustained"),
                "e2e": m.get("e2e"), "e2e_packed": m.get("e2e_packed"), "episode_kernel": m.get("episode_kernel"),
                "roofline": m["roofline"], "clocks": m["clocks"]}
        if "cpu_baseline" in m:
            line["cpu_baseline"] = m["cpu_baseline"]
    # the other BASELINE configurations under the same clock: short device-timed runs + rooflines
    if name == "c2" and not args.no_configs:
        sub = argparse.Namespace(**vars(args))
        sub.steps = max(5, min(args.steps, 20))
        blocks = {}
        for other in ("c3", "c4", "c5"):
            Bo = default_batch(other, ctx.world)
            mo = (measure_rolling if other in ROLLING else measure_steps)(ctx, sub, other, Bo, False)
            blocks[other] = {"workload": WORKLOADS[other][5], "batch_per_gpu": Bo, "n_gpus": ctx.world, "scaling": scaling_of(other, 0),
                             "value": mo["value"], "unit": UNIT, "ms_per_step": mo["ms_per_step"], "steps": sub.steps,
                             "env_steps_per_step": ctx.world * Bo * mo["n"], "roofline": mo["roofline"],
                             "reward_parity_vs_recorded_pass": mo["reward_parity_vs_recorded_pass"],
                             "exchange_checked": mo["exchange_checked"], "clocks": mo["clocks"]}
            ctx.torch.cuda.empty_cache()
        if line is not None:
            line["configs"] = blocks
    if name == "c2" and ctx.world == 1 and not args.no_model:
        try:
            mil = model_in_loop(ctx, args)
        except Exception as e:
            mil = {"error": repr(e)[:300]}
        line["model_in_loop"] = mil
    if line is not None:
        print(json.dumps(line))
    ctx.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="environments per GPU (default: the workload's BASELINE batch)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-configs", action="store_true", help="skip the c3/c4/c5 blocks of the default line")
    ap.add_argument("--no-model", action="store_true", help="skip the model_in_loop block of the default line")
    ap.add_argument("--port", action="store_true", help="reference arm: time the C port even when the Python reference is staged")
    ap.add_argument("--nccl-reduce", action="store_true", help="multi-GPU: reduce the reward statistics with NCCL instead of the peer-memory exchange")
    ap.add_argument("--inline-exchange", action="store_true", help="multi-GPU: keep the peer-memory exchange in-stream (inside the CUDA graph) instead of on a side stream")
    ap.add_argument("--no-reduce", action="store_true", help="multi-GPU diagnostic: skip the reward-statistics reduction altogether")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--ref-envs-per-core", type=int, default=128, help="reference arm: environments per worker process (C1's batch)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
