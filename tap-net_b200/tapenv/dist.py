"""Multi-GPU: environments are independent, so a batch shards across ranks with NO per-step exchange
(one process per GPU, torch.distributed; NCCL over NVLink on the GPU box, gloo in the CPU tests).

The only collective of the path is the end-of-episode reduction of the reward statistics that feed the
critic baseline / advantage normalisation (trainer.py:216-225, :238-240): every rank contributes the
fp64 triple (sum r, sum r^2, count) produced by tapenv_reward, the triples are all-gathered and summed
IN RANK ORDER on every rank, so the result is bit-identical on all ranks and independent of the
collective's internal reduction order.
"""
import torch
import torch.distributed as dist


def shard_range(total, world_size, rank):
    """Contiguous shard [lo, hi) of `total` environments owned by `rank` (remainder spread over the first ranks)."""
    if not 0 <= rank < world_size:
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, rem = divmod(int(total), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard(tensor, world_size, rank, dim=0):
    """This rank's slice of a batch-major tensor."""
    lo, hi = shard_range(tensor.shape[dim], world_size, rank)
    return tensor.narrow(dim, lo, hi - lo)


def combine_partial_sums(sums, group=None):
    """sums: f64 [3] = (sum r, sum r^2, count) of THIS rank (any device the group's backend supports).
    Returns the f64 [3] totals over all ranks, summed in rank order (deterministic, identical on every rank)."""
    if sums.dtype != torch.float64 or sums.numel() != 3:
        raise ValueError("partial sums must be a float64 [3] tensor")
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return sums.clone()
    world = dist.get_world_size(group)
    gathered = torch.empty(world, 3, dtype=torch.float64, device=sums.device)
    dist.all_gather_into_tensor(gathered, sums.reshape(1, 3).contiguous(), group=group)
    total = gathered[0].clone()
    for r in range(1, world):                      # fixed order: rank 0 + rank 1 + ...
        total += gathered[r]
    return total


def reward_statistics(sums_total):
    """(mean, variance, count) of the rewards of the GLOBAL batch from the combined sums."""
    s1, s2, cnt = [float(v) for v in sums_total.tolist()]
    if cnt == 0:
        return 0.0, 0.0, 0
    mean = s1 / cnt
    return mean, max(s2 / cnt - mean * mean, 0.0), int(cnt)


def gather_rewards(reward, group=None):
    """All-gather of the per-environment rewards (equal shard sizes) -> f32 [world*B_local], rank-major."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return reward.clone()
    world = dist.get_world_size(group)
    out = torch.empty(world * reward.numel(), dtype=reward.dtype, device=reward.device)
    dist.all_gather_into_tensor(out, reward.contiguous(), group=group)
    return out


class RewardReducer(object):
    """Runs combine_partial_sums on a side stream so the (latency-bound, 24-byte) exchange overlaps the next
    episode's kernels: nothing on the environment path waits for it -- only the trainer's loss does.

        red = RewardReducer(device)
        total, done = red.reduce_async(sums)      # sums: f64 [3] produced on the current stream
        ...                                       # keep stepping environments
        done.synchronize()  /  torch.cuda.current_stream().wait_event(done)   # before `total` (or `sums`) is reused
    """

    def __init__(self, device, group=None):
        self.device = torch.device(device)
        self.group = group
        self.stream = torch.cuda.Stream(device=self.device)

    def reduce_async(self, sums):
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        done = torch.cuda.Event()
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            total = combine_partial_sums(sums, self.group)
            done.record(self.stream)
        return total, done


class PeerExchange(object):
    """Peer-mapped exchange buffers for tapenv_reward_allreduce (the reward-statistics reduction fused behind the
    reward kernel, over NVLink peer memory).  One per process group; needs torch symmetric memory (CUDA P2P between
    the ranks' GPUs on one node).  Raises if that is unavailable -- callers then use combine_partial_sums (NCCL).

    `stream`: a high-priority side stream the runners launch the exchange on, so that the poll for the slowest rank sits
    BESIDE the next episode's kernels instead of in front of them (nothing on the environment path consumes the totals;
    the trainer's loss does, trainer.py:216-225).
    Rank skew: a call polls for its peers for at most TAPENV_EXCHANGE_TIMEOUT_MS (default 10 s); a timeout yields NaN totals
    and a sticky status that `check()` (called by the runners' / pipelines' result paths) turns into a RuntimeError."""

    def __init__(self, device, group=None):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        from . import _capi
        group = group if group is not None else dist.group.WORLD
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world > 8:
            raise ValueError("PeerExchange supports up to 8 ranks (one node)")
        nbytes = int(_capi.lib.tapenv_comm_bytes())
        self.buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, group)
        torch.cuda.synchronize(device)
        dist.barrier(group)                                # every rank's buffer is zeroed before anyone posts
        comm = _capi.PeerComm()
        comm.world, comm.rank = world, rank
        ptrs = list(self.handle.buffer_ptrs)
        for r in range(world):
            comm.peer[r] = C.c_void_p(int(ptrs[r]))
        self.comm = comm
        self.world, self.rank = world, rank
        self.device = torch.device(device)
        lo, hi = torch.cuda.Stream.priority_range()            # (lowest, highest): numerically larger = lower priority
        self.stream = torch.cuda.Stream(device=self.device, priority=hi)
        off = int(_capi.lib.tapenv_comm_status_offset())
        self._status = self.buf[off: off + 16].view(torch.int64)   # [status, first failed call]

    def status(self):
        """(status, first_failed_call) of this rank's buffer; synchronises with the device."""
        st = self._status.cpu()
        return int(st[0]), int(st[1])

    def check(self):
        st, seq = self.status()
        if st != 0:
            raise RuntimeError("tapenv: the NVLink reward exchange timed out waiting for a peer (rank %d, first failed call %d): "
                               "its totals are NaN.  Raise TAPENV_EXCHANGE_TIMEOUT_MS above the job's worst rank skew, or "
                               "use tapenv.dist.combine_partial_sums (NCCL)." % (self.rank, seq))
