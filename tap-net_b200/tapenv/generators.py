"""generate.generate_blocks (generate.py:773-971) with the rejection loop batched through the kernels (SURVEY.md section 8f
N2, second half).

The RAND generator draws block sets until `tools.calc_positions_lb_greedy(blocks, initial_container, 'C+P+S-lb-hard')` places
every block (generate.py:896-910) -- one environment at a time on the host, ~2 ms per candidate.  Here candidates are
evaluated SPECULATIVELY: K consecutive draws of the reference's own NumPy stream are packed in ONE batch on the GPU, the
first accepted one wins, and the global `np.random` state is rewound and re-advanced so that it ends exactly where the
sequential loop would have left it.  Same signature, same return values, same random stream => `pack.create_dataset` writes
byte-identical files (tests/test_gpu_model_in_loop.py), with the placement work on the device.

    tapenv.install(pack, tools, generate)       # also patches generate.generate_blocks
    pack.create_dataset(10, 64000, 10000, 2, 7, 50, 1, [1, 5], seed=12345)

Branches without a container (container_width < 0: random dependencies, generate.py:912-923) and containers beyond the
compiled limits (3D 7x7 = 49 cells > tapenv_limits.max_cells_3d) run the saved reference function.
"""
import itertools

import numpy as np

from . import _capi
from .episode import calc_positions_lb_greedy, voxel_container
from .rolling import calc_dependent

_original = None          # the reference's generate.generate_blocks, saved by tapenv.install(generate=...)
_original_gt = None       # ... generate.generate_blocks_with_GT
_generate = None          # ... the reference's `generate` module itself (its BPP_Generator_* stay the samplers of the PPSG path)


def _block_probabilities(size_list):
    """generate.py:881-891."""
    if len(size_list) == 4:
        return [0.15, 0.35, 0.35, 0.15]
    if len(size_list) == 5:
        return [0.08, 0.26, 0.32, 0.26, 0.08]
    mu, sigma = 0.5, 0.16
    prob_x = np.linspace(mu - 3 * sigma, mu + 3 * sigma, len(size_list))
    prob = np.exp(-(prob_x - mu) ** 2 / (2 * sigma ** 2)) / (np.sqrt(2 * np.pi) * sigma)
    return prob / np.sum(prob)


def rejection_sample(blocks_num, container_size, size_list, prob_blocks, speculate=8, max_rounds=100000):
    """The `while True` loop of generate.py:896-910 for a FIXED container: returns (blocks int [n,dim], positions int [n,dim])
    of the first draw of the global NumPy stream whose blocks all find a stable place, leaving np.random exactly where the
    sequential loop leaves it.  K = `speculate` draws are evaluated per GPU batch (K grows while nothing is accepted)."""
    n, dim = int(blocks_num), len(container_size)
    K = max(1, int(speculate))
    for _ in range(max_rounds):
        state = np.random.get_state()
        cands = np.stack([np.random.choice(size_list, (n, dim), p=prob_blocks) for _ in range(K)])
        positions, _, stable, _, _ = calc_positions_lb_greedy(cands, container_size, "C+P+S-lb-hard")
        ok = stable.all(dim=1).cpu().numpy()
        if ok.any():
            j = int(np.argmax(ok))
            np.random.set_state(state)                                   # rewind, then consume exactly j+1 draws
            for _ in range(j + 1):
                blocks = np.random.choice(size_list, (n, dim), p=prob_blocks)
            assert np.array_equal(blocks, cands[j])
            return blocks, positions[j].cpu().numpy().astype(int)
        K = min(2 * K, 256)
    raise RuntimeError("tapenv.generators: no feasible block set found")


def generate_blocks(blocks_num, container_size, arm_size, size_range, random_distribution=None, random_num=False, speculate=8):
    """generate.generate_blocks: -> (rotate_blocks, positions, deps_move, rotate_deps_small, rotate_deps_large)."""
    blocks_num = int(blocks_num)
    block_dim = len(container_size)
    container_width = container_size[0]
    size_list = [i for i in range(size_range[0], size_range[1])]
    max_box = size_range[1]
    prob_blocks = _block_probabilities(size_list)
    lim = _capi.limits()

    def on_device(size):
        return size[0] <= lim.max_width_2d if block_dim == 2 else size[0] * size[1] <= lim.max_cells_3d

    if container_width < 0 or (container_width > 0 and not on_device(container_size)) or blocks_num > lim.max_blocks:
        if _original is None:
            raise _capi.TapEnvError(_capi.ELIMIT, "generate_blocks: this configuration needs the reference function (tapenv.install(generate=...))")
        return _original(blocks_num, container_size, arm_size, size_range, random_distribution, random_num)

    container_size = list(container_size)
    first = None
    if container_width == 0:
        # the width is drawn AFTER the first block set (generate.py:899-904): that draw is evaluated on its own
        first = np.random.choice(size_list, (blocks_num, block_dim), p=prob_blocks)
        container_width = np.random.randint(5, 11)
        container_size = ([container_width, blocks_num * max_box + 10] if block_dim == 2
                          else [container_width, container_width, blocks_num * max_box + 10])
        if not on_device(container_size):
            raise _capi.TapEnvError(_capi.ELIMIT, "generate_blocks: random container %s beyond the compiled limits" % container_size)
    blocks = positions = None
    if first is not None:
        pos, _, stable, _, _ = calc_positions_lb_greedy(first, container_size, "C+P+S-lb-hard")
        if int(np.sum(stable)) == blocks_num:
            blocks, positions = first, pos
    if blocks is None:
        blocks, positions = rejection_sample(blocks_num, container_size, size_list, prob_blocks, speculate)
    return _emit(blocks, positions, container_size, arm_size)


def _emit(blocks, positions, container_size, arm_size):
    """The common tail of generate_blocks (generate.py:926-971) and generate_blocks_with_GT (:184-229)."""
    block_dim = len(container_size)
    # precedence of the accepted packing (generate.calc_dependent, generate.py:575-771) from the block intervals
    adj = calc_dependent(blocks, positions, container_size, arm_size).astype(np.float64)
    deps_move, deps_left, deps_right, deps_forward, deps_backward = adj
    deps_up = np.zeros_like(deps_move)
    deps_down = np.zeros_like(deps_move)

    # every rotation of the blocks with the dependencies its last axis needs
    rotate_blocks, rotate_deps_small, rotate_deps_large = [], [], []
    bt = blocks.transpose()
    flat = {0: (deps_left.flatten(), deps_right.flatten()),
            1: ((deps_forward.flatten(), deps_backward.flatten()) if block_dim == 3 else (deps_down.flatten(), deps_up.flatten())),
            2: (deps_down.flatten(), deps_up.flatten())}
    for p in itertools.permutations(range(block_dim)):
        small, large = flat[p[-1]]
        rotate_deps_small.append(small)
        rotate_deps_large.append(large)
        rotate_blocks.append(bt[list(p)].flatten())
    positions = np.array(positions).transpose().flatten()
    return (np.array(rotate_blocks), positions, deps_move.transpose().flatten(),
            np.array(rotate_deps_small), np.array(rotate_deps_large))


def generate_blocks_with_GT(blocks_num, gt_packing_size, initial_container_size, arm_size, size_range, input_type, data_index,
                            allow_rot=True):
    """generate.generate_blocks_with_GT (generate.py:17-229, the PPSG generator behind pack.create_dataset_gt / BASELINE C4):
    draw a perfect packing of the target container (the reference's own BPP_Generator_*), then up to 20 random unpacking
    orders + rotations of its blocks, each packed into the INITIAL container with
    `tools.calc_positions_lb_greedy(blocks, initial_container_size, 'C+P+S-lb-hard')` (:112) and accepted if every block
    is stable and the reversed order can be unpacked.  The reference evaluates those 20 tries one after the other on the
    host (~3 ms of NumPy each at n=20, plus a voxel-scanning calc_dependent per stable try; SURVEY: median 3.7 s per sample).
    The random draws of a try do not depend on the outcome of the tries before it, so here all 20 are drawn first (the
    global NumPy stream position after each is remembered), packed in ONE batch on the GPU, checked in order with the
    interval-arithmetic calc_dependent, and np.random is put back to where the sequential loop would have stopped.  Same
    signature, same return values, same stream => byte-identical datasets.  Needs tapenv.install(pack, tools, generate);
    3D (7x7 initial container = 49 cells) and other shapes beyond the compiled limits run the saved reference function."""
    if _generate is None or _original_gt is None:
        raise _capi.TapEnvError(_capi.EINVAL, "generate_blocks_with_GT: the reference's samplers are needed (tapenv.install(generate=...))")
    blocks_num = int(blocks_num)
    block_dim = len(gt_packing_size)
    lim = _capi.limits()
    ics = list(initial_container_size)
    ok_shape = (block_dim == 2 and 0 < ics[0] <= lim.max_width_2d) or (block_dim == 3 and ics[0] > 0 and ics[0] * ics[1] <= lim.max_cells_3d)
    if not ok_shape or blocks_num > lim.max_blocks:
        return _original_gt(blocks_num, gt_packing_size, initial_container_size, arm_size, size_range, input_type, data_index, allow_rot)
    lo, hi = size_range[0], size_range[1]
    rotates = np.array([list(p) for p in itertools.permutations(range(block_dim))])
    TRIES = 20                                                            # loop_time = 20 (:85-89)
    while True:
        while True:                                                       # candidate perfect packing (:68-75)
            if block_dim == 3:
                gt_blocks, gt_positions, gt_container = _generate.BPP_Generator_3D(blocks_num, gt_packing_size, size_range)
            else:
                gt_blocks, gt_positions, gt_container = _generate.BPP_Generator_2D_easy(blocks_num, gt_packing_size, size_range)
            if ((gt_blocks >= lo) & (gt_blocks < hi)).all():
                break
        gt_deps = calc_dependent(gt_blocks, gt_positions, list(gt_container.shape), arm_size)[0].astype(np.float64)
        # the 20 tries' draws: a random topological order of the unpacking graph + a random rotation per block (:91-108)
        # The reference keeps an n x n matrix and draws np.random.choice(candidate_idx) -- which consumes exactly one
        # randint(0, len(candidate_idx)) of the legacy stream -- from ALL rows without open dependencies, the already chosen
        # ones included (a re-pick is discarded, :97-101: ~125 draws per order at n=20).  Same draws here on one bit mask per
        # row (bit j of need[i]: i still waits for j).
        cands, states = [], []
        need0 = [int(sum(1 << j for j in range(blocks_num) if gt_deps[j, i] != 0)) for i in range(blocks_num)]
        randint = np.random.randint
        for _ in range(TRIES):
            order, chosen = [], 0
            need = list(need0)
            free = [i for i in range(blocks_num) if need[i] == 0]         # ascending, like np.where
            while len(order) < blocks_num:                                # (np.sum(my_deps) > 0 implies an unchosen row)
                idx = free[randint(0, len(free))]
                if (chosen >> idx) & 1:
                    continue
                order.append(idx)
                chosen |= 1 << idx
                clear = ~(1 << idx)
                need = [m & clear for m in need]
                free = [i for i in range(blocks_num) if need[i] == 0]
            blocks = gt_blocks[order]
            if allow_rot:
                for i in range(len(blocks)):
                    blocks[i] = blocks[i][rotates[np.random.randint(0, len(rotates))]]
            cands.append(blocks)
            states.append(np.random.get_state())
        batch = np.stack(cands)
        positions, _, stable, _, _ = calc_positions_lb_greedy(batch, ics, "C+P+S-lb-hard")      # ONE launch for the 20 tries
        stable = np.asarray(stable.cpu()) if hasattr(stable, "cpu") else np.asarray(stable)
        positions = np.asarray(positions.cpu()) if hasattr(positions, "cpu") else np.asarray(positions)
        for j in range(TRIES):
            if int(np.sum(stable[j])) < blocks_num:                       # :114
                continue
            blocks, pos = cands[j], positions[j].astype(int)
            adj = calc_dependent(blocks, pos, ics, arm_size).astype(np.float64)
            # :117-160 -- walk the packing order backwards: block s can leave iff nothing rests on it and one side of every
            # rotation axis is free; then it no longer blocks anybody
            dm, dl, dr, df, db = [a.transpose().copy() for a in adj]
            work = True
            for sidx in reversed(range(blocks_num)):
                x = np.sum(df[sidx]) * np.sum(db[sidx])
                y = np.sum(dl[sidx]) * np.sum(dr[sidx])
                if input_type == "simple":
                    x = y = 0
                if np.sum(dm[sidx]) == 0 and x == 0 and y == 0:
                    for a in (dm, dl, dr, df, db):
                        a[:, sidx] = 0
                else:
                    work = False
                    break
            if work:
                np.random.set_state(states[j])                            # where the sequential loop stops drawing
                return _emit(blocks, pos, ics, arm_size)


# kept for callers that want the voxel grid of the accepted packing (generate.py:908 hands it to calc_dependent)
def packing_container(blocks, positions, container_size):
    return voxel_container(positions, blocks, container_size)
