"""PACKDataset with the reference's constructor and item layout (pack.py:25-273), built from the same six text
files a reference dataset directory holds (blocks / pos / container / dep_move / dep_small / dep_large .txt).

Host-side, one-off work: plain NumPy index arithmetic, tensors stay on the CPU exactly like the reference's
(the trainer moves each batch with .cuda(), trainer.py:189-192).

Layouts produced (R = dim! rotations, candidates rotation-major, column j = r*n + i):
    static  f32 [N, 1+dim, n*R]   row 0 = block id i, rows 1..dim = edge lengths of block i in rotation r
    dynamic f32 [N, 3n, n*R]      rows [0,n) move, [n,2n) rot-small, [2n,3n) rot-large precedence ('bot')
"""
import math

import numpy as np
import torch
from torch.utils.data import Dataset


def _load(path):
    return np.atleast_2d(np.loadtxt(path).astype("float32"))


def packed_words(dyn_rows, S):
    """u32 words per environment of the bit-row `dynamic` (rows padded to 16 bytes; == tapenv_packed_words)."""
    return ((dyn_rows * S + 31) // 32 + 3) // 4 * 4


def pack_inputs(static, dynamic):
    """Host-side compaction of a batch for upload: static f32 [B,rows,S] (small non-negative integers) -> u8,
    dynamic f32 [B,3n,S] (0/1, pack.py:101-223) -> int32 [B, packed_words] bit rows, bit (row*S + col) little-endian.
    Expanded on the device by BatchedContainers.reset_packed (tapenv_reset_packed)."""
    static = np.asarray(static)
    dynamic = np.asarray(dynamic)
    if not ((static == np.round(static)).all() and static.min() >= 0 and static.max() < 256):
        raise ValueError("static is not u8-representable")
    if not np.isin(dynamic, (0.0, 1.0)).all():
        raise ValueError("dynamic must hold 0/1")
    B, rows, S = dynamic.shape
    words = packed_words(rows, S)
    bits = np.zeros((B, words * 4), np.uint8)
    pk = np.packbits(dynamic.reshape(B, -1).astype(np.uint8), axis=1, bitorder="little")
    bits[:, :pk.shape[1]] = pk
    return np.ascontiguousarray(static.astype(np.uint8)), np.ascontiguousarray(bits).view("<i4")


class PACKDataset(Dataset):
    def __init__(self, data_file, blocks_num, num_samples, seed, input_type, heightmap_type, allow_rot,
                 container_width, mix_data_file=None, unit=1, no_precedence=False):
        super(PACKDataset, self).__init__()
        if seed is None:
            seed = np.random.randint(123456)
        np.random.seed(seed)
        torch.manual_seed(seed)
        n, N = int(blocks_num), int(num_samples)
        move = _load(data_file + "dep_move.txt")
        small = _load(data_file + "dep_small.txt")
        large = _load(data_file + "dep_large.txt")
        blocks = _load(data_file + "blocks.txt")
        positions = _load(data_file + "pos.txt")
        container = _load(data_file + "container.txt")
        if mix_data_file is not None:
            # pack.py:67-97: the first half of the samples from data_file, the second from mix_data_file.  Per-sample files are
            # cut at num_samples/2 rows, per-rotation files at HALF THE ROWS OF data_file's blocks.txt (the reference's own rule)
            num_mid, rot_mid = int(N / 2), int(len(blocks) / 2)
            mix = lambda name: _load(mix_data_file + name)
            move = np.vstack((move[:num_mid], mix("dep_move.txt")[:num_mid]))
            positions = np.vstack((positions[:num_mid], mix("pos.txt")[:num_mid]))
            container = np.vstack((container[:num_mid], mix("container.txt")[:num_mid]))
            small = np.vstack((small[:rot_mid], mix("dep_small.txt")[:rot_mid]))
            large = np.vstack((large[:rot_mid], mix("dep_large.txt")[:rot_mid]))
            blocks = np.vstack((blocks[:rot_mid], mix("blocks.txt")[:rot_mid]))

        dim = positions.reshape(N, -1, n).shape[1]                      # pack.py:108
        R_file = math.factorial(dim)
        # blocks.txt: per sample R lines, each line [dim][n] row-major -> static[b, 1+d, r*n+i]   (pack.py:113-121)
        edges = blocks.reshape(N, R_file, dim, n).transpose(0, 2, 1, 3).reshape(N, dim, R_file * n)
        edges = np.ceil(edges * unit).astype(np.float32)               # pack.py:123-125
        # dep_small / dep_large: per sample R lines, each an n x n matrix (row i, column j) -> [b, i, r*n+j]
        def rot_dep(a):
            return np.ascontiguousarray(a.reshape(N, R_file, n, n).transpose(0, 2, 1, 3).reshape(N, n, R_file * n))
        small_t, large_t = rot_dep(small), rot_dep(large)
        # dep_move: one n x n matrix per sample, stored transposed (pack.py:104-106), identical for every rotation
        move_t = move.reshape(N, n, n).transpose(0, 2, 1)

        R = R_file if allow_rot else 1                                  # pack.py:139-141
        if not allow_rot:
            edges = edges[:, :, :n]
        ids = np.tile(np.arange(n, dtype=np.float32), R)[None, None, :].repeat(N, axis=0)
        move_t = np.tile(move_t, (1, 1, R))
        cont = np.tile(container.reshape(N, 1, n), (1, 1, R)).astype(np.float32)
        if no_precedence:
            move_t, small_t, large_t = np.zeros_like(move_t), np.zeros_like(small_t), np.zeros_like(large_t)

        if input_type in ("simple", "rot"):
            static, dynamic = np.concatenate([ids, edges], 1), move_t
        elif input_type == "bot":
            static, dynamic = np.concatenate([ids, edges], 1), np.concatenate([move_t, small_t, large_t], 1)
        elif input_type in ("bot-rot", "use-static", "use-pnet"):
            static = np.concatenate([ids, edges], 1)
            dynamic = np.concatenate([move_t, np.zeros_like(small_t), np.zeros_like(large_t)], 1)
        elif input_type in ("mul", "mul-with"):
            static = np.concatenate([ids, edges, cont], 1)
            dynamic = np.concatenate([move_t, small_t, large_t], 1)
        elif input_type == "rot-old":                                   # pack.py:218-223: n movement rows + one zero rotate-state row
            static, dynamic = np.concatenate([ids, edges], 1), np.concatenate([move_t, np.zeros_like(ids)], 1)
        else:
            raise ValueError("unknown input_type %r" % (input_type,))   # the reference prints 'Dataset OHHHHH' and dies later
        self.static = torch.from_numpy(np.ascontiguousarray(static, dtype=np.float32))
        self.dynamic = torch.from_numpy(np.ascontiguousarray(dynamic, dtype=np.float32))

        # decoder inputs: zeros shaped like the encoded heightmap (pack.py:228-266)
        static_dim = dim + (1 if input_type == "mul-with" else 0)
        hm_num = 1
        if heightmap_type == "diff":
            hm_w = container_width * unit - 1 if dim == 2 else container_width * unit
            if dim == 3:
                hm_num = 2
        else:
            hm_w = container_width * unit
        hm_w = int(np.ceil(hm_w))
        hm_l = int(np.ceil(container_width * unit))
        if input_type in ("mul", "mul-with"):
            if dim == 2:
                hm_w *= 2
            else:
                hm_num *= 2
        self.decoder_static = torch.zeros(N, static_dim, 1, requires_grad=True)
        if dim == 2:
            self.decoder_dynamic = torch.zeros(N, hm_w, 1, requires_grad=True)
        else:
            self.decoder_dynamic = torch.zeros(N, hm_num, hm_w, hm_l, requires_grad=True)
        self.num_samples = N

    def __len__(self):
        return self.num_samples

    def packed(self):
        """The whole set in the compact upload format -> (static_u8 [N,rows,S] uint8, dynamic_bits [N,words] int32)
        CPU tensors (pin them once; BatchedContainers.reset_packed expands a batch on the device)."""
        if getattr(self, "_packed", None) is None:
            su8, bits = pack_inputs(self.static.numpy(), self.dynamic.numpy())
            self._packed = (torch.from_numpy(su8), torch.from_numpy(bits))
        return self._packed

    def __getitem__(self, idx):
        return (self.static[idx], self.dynamic[idx], self.decoder_static[idx], self.decoder_dynamic[idx])
