"""GPU box with >= 2 GPUs: the fused reward + peer-memory exchange (tapenv_reward_allreduce) against the NCCL
reduction and the single-process sums.  Spawned as 2 ranks; skipped on boxes with one GPU."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out, overlap=True, skip_last_on_rank1=False):
    for p in (ROOT, os.path.join(ROOT, "tap-net_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if skip_last_on_rank1:
        os.environ["TAPENV_EXCHANGE_TIMEOUT_MS"] = "300"
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
                                  torch.from_numpy(np.ascontiguousarray(ptrs[:, lo:hi])).to(dev), use_graph=True, exchange=ex,
                                  overlap_exchange=overlap)
    assert runner.tail.overlap == overlap
    totals = []
    for _ in range(6):                                   # more calls than the exchange ring is deep
        runner.run()
        runner.tail.wait_total()                         # the exchange may run on its side stream
        totals.append(runner.total.clone())
    torch.cuda.synchronize(dev)
    assert ex.status() == (0, 0)
    ex.check()
    nccl_total = tapenv.dist.combine_partial_sums(runner.sums)
    res = (runner.sums.cpu().tolist(), [t.cpu().tolist() for t in totals], nccl_total.cpu().tolist(), runner.reward.cpu().numpy())
    timed_out = None
    if skip_last_on_rank1:                               # rank 1 never makes the 7th call: rank 0 must time out LOUDLY
        dist.barrier()
        if rank == 0:
            runner.run()
            runner.tail.wait_total()
            torch.cuda.synchronize(dev)
            st, seq = ex.status()
            raised = False
            try:
                ex.check()
            except RuntimeError:
                raised = True
            timed_out = (st, seq, bool(torch.isnan(runner.total).all()), raised)
        dist.barrier()
    out[rank] = res + (timed_out,)
    dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [True, False], ids=["side-stream", "in-graph"])
def test_fused_reward_exchange_two_ranks(overlap):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, port = 2, 29600 + (os.getpid() + int(overlap)) % 1000
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out, overlap), nprocs=world, join=True)
    (s0, t0, n0, r0, _), (s1, t1, n1, r1, _) = out[0], out[1]
    want = [s0[i] + s1[i] for i in range(3)]             # rank-order sum
    assert all(t == want for t in t0) and all(t == want for t in t1)     # bit-identical on both ranks, every call
    assert n0 == want and n1 == want                                      # == the NCCL path
    r = np.concatenate([r0, r1]).astype(np.float64)
    assert want[2] == 600.0 and abs(want[0] - r.sum()) < 1e-9


def test_exchange_timeout_is_loud():
    """a peer that never calls must not turn the statistics into silent NaNs -- sticky status + RuntimeError."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, port = 2, 29700 + os.getpid() % 1000
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out, True, True), nprocs=world, join=True)
    st, seq, all_nan, raised = out[0][4]
    assert st == 1 and seq == 7 and all_nan and raised
