"""Whole-episode entries with the reference's signatures: tools.calc_positions_lb_greedy (tools.py:2393-2449),
tools.calc_positions_mcs (:3213-3315) and pack.reward (pack.py:378-473), batched over environments.

The placement runs in the CUDA kernels (one add_new_block launch per block for the whole batch); the
C/P/S arithmetic on the resulting integer state is IEEE fp64 division in torch, the same operations the
reference performs in NumPy."""
import numpy as np
import torch

from .containers import BatchedContainers
from .config import rotate_types


def _pack_sequence(blocks, container_size, reward_type, packing_strategy):
    blocks = torch.as_tensor(blocks)
    single = blocks.dim() == 2
    if single:
        blocks = blocks.unsqueeze(0)
    if not blocks.is_cuda:
        blocks = blocks.cuda()
    blocks = blocks.to(torch.float32)
    B, n, dim = blocks.shape
    env = BatchedContainers(container_size, n, reward_type, "full", packing_strategy=packing_strategy, batch_size=B,
                            device=blocks.device)
    for i in range(n):
        env.add_new_blocks(blocks[:, i].contiguous())
    env.check_flags()
    return env, single


def voxel_container(positions, blocks, container_size):
    """The reference's voxel grid (0 empty / -1 "empty under a block" / k+1 block id) that calc_positions_* return in
    second place, rebuilt on the host from the placements the kernels recorded: block k writes k+1 into its extent and -1
    into the still-empty cells below it (tools.py:2168-2169, :2661-2665, :3045-3049), in arrival order.  A block that was
    not placed (its position stays at the origin, tools.py:2084-2087) is recognised by replaying the heightmap: a placed
    block always sits exactly on the highest column under its footprint."""
    size = [int(v) for v in container_size]
    dim = len(size)
    grid = np.zeros(size, dtype=int)
    h = np.zeros(size[:-1], dtype=int)
    for k in range(len(blocks)):
        b = [int(v) for v in blocks[k]]
        p = [int(v) for v in positions[k]]
        foot = tuple(slice(p[d], p[d] + b[d]) for d in range(dim - 1))
        if any(b[d] < 1 or p[d] + b[d] > size[d] for d in range(dim - 1)) or b[-1] < 1:
            continue                                   # no EMS for this block: not placed
        if int(h[foot].max()) != p[-1]:
            continue                                   # not resting on its footprint: not placed
        z = p[-1]
        under = grid[foot + (slice(0, z),)]
        under[under == 0] = -1
        grid[foot + (slice(z, z + b[-1]),)] = k + 1
        h[foot] = z + b[-1]
    return grid


def _result(env, single):
    sc = env.scalars.to(torch.float64)
    valid, empty, nstable = sc[:, 0], sc[:, 1], sc[:, 2]
    hmax = env.heightmap.reshape(env.batch_size, -1).max(dim=1).values
    cells = env._cells
    box = hmax.to(torch.float64) * cells
    n = env.blocks_num
    ratio = valid / box + valid / (empty + valid) + nstable / n          # C + P + S, NOT divided by 3 (tools.py:2438-2446)
    scores = torch.stack([sc[:, 0], box, sc[:, 1], sc[:, 2], hmax.to(torch.float64)], 1).to(torch.int64)
    positions, stable, heightmap = env.positions.clone(), env.stable.bool(), env.heightmap.clone()
    if single:                                         # the reference's call form: the voxel `container` in second place
        pos = positions[0].cpu().numpy().astype(np.int64)
        grid = voxel_container(pos, env.blocks[0].cpu().numpy(), env.container_size)
        return (pos, grid, [bool(v) for v in stable[0].tolist()], float(ratio[0].item()), [int(v) for v in scores[0].tolist()])
    return positions, heightmap, stable, ratio, scores


def calc_positions_lb_greedy(blocks, container_size, reward_type):
    """blocks [n,dim] (reference form) -> (positions, container, stable, ratio, scores) exactly as tools.py:2393-2449:
    `container` is the voxel grid (generate.calc_dependent indexes it, generate.py:112,:908), ratio = C+P+S and
    scores = [valid_size, box_size, empty_size, stable_num, packing_height].
    blocks [B,n,dim] (batched extension, device tensors out) -> the heightmap [B,W(,L)] stands in second place."""
    env, single = _pack_sequence(blocks, container_size, reward_type, "LB_GREEDY")
    return _result(env, single)


def calc_positions_mcs(blocks, container_size, reward_type):
    env, single = _pack_sequence(blocks, container_size, reward_type, "MACS")
    return _result(env, single)


def reward(static, tour_indices, reward_type, input_type, allow_rot, container_width, container_height,
           packing_strategy="LB_GREEDY"):
    """pack.reward: re-pack a finished tour and return -(C+P+S) as f32 [B] (pack.py:378-473)."""
    static = static.detach()
    if not static.is_cuda:
        static = static.cuda()
    if input_type in ("mul", "mul-with"):
        return _reward_two_containers(static, tour_indices, reward_type, input_type, allow_rot, container_width,
                                      container_height, packing_strategy)
    dim = static.shape[1] - 1
    R = rotate_types(dim, allow_rot)
    n = static.shape[2] // R
    size = [container_width, container_height] if dim == 2 else [container_width, container_width, container_height]
    idx = tour_indices.to(static.device).long()[:, :n]
    seq = torch.gather(static[:, 1:1 + dim], 2, idx.unsqueeze(1).expand(-1, dim, -1)).transpose(1, 2).contiguous()   # [B,n,dim]
    strat = "MACS" if packing_strategy in ("MACS", "MUL") else "LB_GREEDY"
    env, _ = _pack_sequence(seq, size, reward_type, strat)
    ratio = _result(env, False)[3]
    return -ratio.to(torch.float32)


def _reward_two_containers(static, tour_indices, reward_type, input_type, allow_rot, container_width, container_height,
                           packing_strategy):
    """pack.reward for 'mul' / 'mul-with' (pack.py:455-468): the tour's blocks are split by their target-container id,
    each part is packed on its own (calc_positions_* on blocks_a / blocks_b) and the two C+P+S sums are averaged; an
    empty part scores 0."""
    from .containers import BatchedContainerPairs
    dim = static.shape[1] - 2
    R = rotate_types(dim, allow_rot)
    n = static.shape[2] // R
    B = static.shape[0]
    size = [container_width, container_height] if dim == 2 else [container_width, container_width, container_height]
    idx = tour_indices.to(static.device).long()[:, :n]
    seq = torch.gather(static[:, 1:1 + dim], 2, idx.unsqueeze(1).expand(-1, dim, -1)).transpose(1, 2).contiguous()   # [B,n,dim]
    tgt = torch.gather(static[:, -1], 1, idx).contiguous()                                                               # [B,n]
    strat = "MACS" if packing_strategy in ("MACS", "MUL") else "LB_GREEDY"
    pairs = BatchedContainerPairs(size, n, reward_type, "full", packing_strategy=strat, batch_size=B, device=static.device,
                                  input_type=input_type, allow_rot=allow_rot)
    for i in range(n):
        pairs.add_new_blocks(seq[:, i].contiguous(), tgt[:, i].contiguous())
    pairs.check_flags()
    total = torch.zeros(B, dtype=torch.float64, device=static.device)
    for env in (pairs.a, pairs.b):
        sc = env.scalars.to(torch.float64)
        valid, empty, nstable, k = sc[:, 0], sc[:, 1], sc[:, 2], sc[:, 3]
        hmax = env.heightmap.reshape(B, -1).max(dim=1).values.to(torch.float64)
        ratio = valid / (hmax * env._cells) + valid / (empty + valid) + nstable / k      # tools.py:2438-2446 with n = len(part)
        total += torch.where(k > 0, ratio, torch.zeros_like(ratio))                       # `scores_a = 0` for an empty part
    return -(total / 2).to(torch.float32)
