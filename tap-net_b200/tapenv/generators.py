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
    # precedence of the accepted packing (generate.calc_dependent, generate.py:575-771) from the block intervals
    adj = calc_dependent(blocks, positions, container_size, arm_size).astype(np.float64)
    deps_move, deps_left, deps_right, deps_forward, deps_backward = adj
    deps_up = np.zeros_like(deps_move)
    deps_down = np.zeros_like(deps_move)

    # generate.py:926-971: every rotation of the blocks with the dependencies its last axis needs
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


# kept for callers that want the voxel grid of the accepted packing (generate.py:908 hands it to calc_dependent)
def packing_container(blocks, positions, container_size):
    return voxel_container(positions, blocks, container_size)
