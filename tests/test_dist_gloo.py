"""CPU, world_size 2 over gloo: the host-side multi-GPU logic (sharding + deterministic combination of the
per-rank reward partial sums).  The per-rank sums themselves come from the CPU oracle here; on the GPU box
tests/test_gpu_parity.py checks the kernel that produces them."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    for p in (ROOT, os.path.join(ROOT, "tap-net_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tapenv import dist as tdist
    from tests.golden_io import load_inputs
    from tests.rollout import random_valid_ptrs
    from oracle import oracle
    static, dynamic = load_inputs("rand2d_n10.npz", 101)           # odd size: uneven shards
    ptrs = random_valid_ptrs(static, dynamic, [5, 50], seed=3)
    lo, hi = tdist.shard_range(101, world, rank)
    o = oracle.episode_batch(static[lo:hi], dynamic[lo:hi], ptrs[:, lo:hi], [5, 50], "C+P+S-lb-soft", "diff", "LB_GREEDY",
                             want=("reward",))
    r = torch.from_numpy(o["reward"]).double()
    sums = torch.stack([r.sum(), (r * r).sum(), torch.tensor(float(hi - lo), dtype=torch.float64)])
    total = tdist.combine_partial_sums(sums)
    all_sums = [torch.zeros(3, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(all_sums, sums)
    out[rank] = (lo, hi, total.tolist(), [s.tolist() for s in all_sums], tdist.reward_statistics(total))
    dist.destroy_process_group()


def test_sharded_reward_statistics_world2():
    world, port = 2, 29500 + os.getpid() % 2000
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    (lo0, hi0, t0, parts0, st0), (lo1, hi1, t1, parts1, st1) = out[0], out[1]
    assert (lo0, hi0, lo1, hi1) == (0, 51, 51, 101)
    assert t0 == t1 and st0 == st1                                    # identical on every rank, bit for bit
    assert t0 == [parts0[0][i] + parts0[1][i] for i in range(3)]      # rank-order sum
    # shard-invariance: equals the single-process statistics of the whole batch
    sys.path.insert(0, ROOT)
    from tests.golden_io import load_inputs
    from tests.rollout import random_valid_ptrs
    from oracle import oracle
    static, dynamic = load_inputs("rand2d_n10.npz", 101)
    ptrs = random_valid_ptrs(static, dynamic, [5, 50], seed=3)
    r = oracle.episode_batch(static, dynamic, ptrs, [5, 50], "C+P+S-lb-soft", "diff", "LB_GREEDY", want=("reward",))["reward"].astype(np.float64)
    assert t0[2] == 101.0 and abs(t0[0] - r.sum()) < 1e-9 and abs(st0[0] - r.mean()) < 1e-12


def test_shard_range_partitions():
    from tapenv.dist import shard_range
    for total in (0, 1, 7, 8, 4096, 65537):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)
