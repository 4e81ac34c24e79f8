"""GPU box with >= 2 GPUs: the fused reward + peer-memory exchange (tapenv_reward_allreduce) against the NCCL
reduction and the single-process sums.  Spawned as 2 ranks; skipped on boxes with one GPU."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    for p in (ROOT, os.path.join(ROOT, "tap-net_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import tapenv
    from tests.golden_io import load_inputs
    from tests.rollout import random_valid_ptrs
    static, dynamic = load_inputs("rand2d_n10.npz", 600)
    ptrs = random_valid_ptrs(static, dynamic, [5, 50], seed=3)
    lo, hi = tapenv.dist.shard_range(600, world, rank)
    B = hi - lo
    env = tapenv.BatchedContainers([5, 50], 10, "C+P+S-lb-soft", "diff", batch_size=B, device=dev)
    ex = tapenv.dist.PeerExchange(dev)
    runner = tapenv.EpisodeRunner(env, torch.from_numpy(static[lo:hi]).to(dev), torch.from_numpy(dynamic[lo:hi]).to(dev),
                                  torch.from_numpy(np.ascontiguousarray(ptrs[:, lo:hi])).to(dev), use_graph=True, exchange=ex)
    totals = []
    for _ in range(6):                                   # more calls than the exchange ring is deep
        runner.run()
        totals.append(runner.total.clone())
    torch.cuda.synchronize(dev)
    nccl_total = tapenv.dist.combine_partial_sums(runner.sums)
    out[rank] = (runner.sums.cpu().tolist(), [t.cpu().tolist() for t in totals], nccl_total.cpu().tolist(),
                 runner.reward.cpu().numpy())
    dist.destroy_process_group()


def test_fused_reward_exchange_two_ranks():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, port = 2, 29600 + os.getpid() % 1000
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    (s0, t0, n0, r0), (s1, t1, n1, r1) = out[0], out[1]
    want = [s0[i] + s1[i] for i in range(3)]             # rank-order sum
    assert all(t == want for t in t0) and all(t == want for t in t1)     # bit-identical on both ranks, every call
    assert n0 == want and n1 == want                                      # == the NCCL path
    r = np.concatenate([r0, r1]).astype(np.float64)
    assert want[2] == 600.0 and abs(want[0] - r.sum()) < 1e-9
