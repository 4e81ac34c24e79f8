"""CPU: the MACS 3D placement the kernels run (tap-net_b200/csrc/place_macs3d.cuh: level masks, warp-uniform walk, recorded
ties) compiled for the HOST with one emulated lane, against the oracle's literal restatement of calc_one_position_mcs_3d
(tools.py:2751-3165) -- every step: heightmap, valid / empty / #stable; at the end: positions, stability flags, the voxel
grid and the incrementally edited interval lists.  The lane-parallel halves of the same functions are covered on the GPU
(tests/test_gpu_parity.py: macs3d trajectories, fuzz, mul, rolling, whole-episode)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hostlib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hostchk") / "libhost_checks.so")
    src = os.path.join(ROOT, "tap-net_b200", "csrc", "host_checks.cpp")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", out, src])
    lib = C.CDLL(out)
    lib.tapenv_host_macs3d_episode.argtypes = [C.c_int] * 7 + [C.c_void_p] * 7
    lib.tapenv_host_macs3d_episode.restype = C.c_int
    return lib


def reward_flags(rt):
    """tapenv_config_init's substring tests (tools.py:2113, :2135, :2138, :2718, :2709)."""
    return (1 if rt.endswith("hard") else 0) | (2 if "P" in rt else 0) | (4 if "S" in rt else 0) | (8 if "mcs" in rt else 0) | \
           (16 if rt.startswith("mcs") else 0)


def host_episode(lib, W, L, H, n, rt, blocks):
    cap, lcap = n, max(n + 2, W + 4)
    cells = W * L
    blocks = np.ascontiguousarray(blocks, dtype=np.int32)
    hm = np.zeros((n, cells), np.int32); sc = np.zeros((n, 4), np.int32)
    pos = np.zeros((cap, 3), np.int32); stb = np.zeros(cap, np.uint8)
    vox = np.zeros(cells * H, np.int16); lists = np.zeros(H * L * lcap, np.int8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    anomaly = lib.tapenv_host_macs3d_episode(W, L, H, cap, lcap, reward_flags(rt), n, p(blocks), p(hm), p(sc), p(pos), p(stb), p(vox), p(lists))
    return dict(anomaly=anomaly, heightmap=hm, scal=sc, positions=pos, stable=stb, voxels=vox.reshape(W, L, H),
                lists=lists.reshape(H, L, lcap))


CASES = [(5, 5, 50, 10, "C+P+S-mcs-soft"), (5, 5, 50, 10, "C+P+S-mcs-hard"), (5, 5, 60, 12, "mcs-soft"), (5, 5, 60, 12, "mcs-hard"),
         (4, 6, 80, 16, "C+P-mcs-soft"), (6, 4, 80, 16, "C+P+S-mul-hard"), (3, 3, 40, 10, "C+P+S-mcs-soft"), (7, 4, 120, 24, "C+P+S-mcs-hard"),
         (2, 2, 60, 12, "C+P+S-mul-soft"), (8, 4, 100, 20, "C+P+S-mcs-soft"), (5, 5, 250, 50, "C+P+S-mcs-soft")]


@pytest.mark.parametrize("W,L,H,n,rt", CASES)
def test_host_build_matches_oracle(hostlib, W, L, H, n, rt):
    rng = np.random.RandomState(W * 1000 + L * 100 + n)
    episodes = 60 if n <= 24 else 12
    for ep in range(episodes):
        # block edges never wider than the container (MACS 3D walks the phantom (0,0,0) extents of unplaced blocks and the
        # reference raises IndexError when one sticks out, tools.py:2915-2922)
        hi = min(5, W, L) + 1
        blocks = np.stack([rng.randint(1, hi, size=n), rng.randint(1, hi, size=n), rng.randint(1, 6, size=n)], 1)
        env = oracle.Container([W, L, H], n, rt, "full", packing_strategy="MACS")
        want_h, want_s = [], []
        failed = False
        for t in range(n):
            try:
                env.add_new_block(blocks[t].astype(np.float32))
            except IndexError:
                failed = True                       # the reference raises: the kernel flags instead; nothing to compare
                break
            want_h.append(env.heightmap.reshape(-1).copy())
            want_s.append((env.valid_size, env.empty_size, sum(env.stable[:t + 1]), t + 1))
        got = host_episode(hostlib, W, L, H, n, rt, blocks)
        if failed:
            assert got["anomaly"] != 0
            continue
        assert got["anomaly"] == 0, (ep, got["anomaly"])
        assert np.array_equal(got["heightmap"], np.stack(want_h)), ep
        assert np.array_equal(got["scal"], np.asarray(want_s)), ep
        assert np.array_equal(got["positions"], env.positions), ep
        assert [bool(v) for v in got["stable"]] == env.stable, ep
        assert np.array_equal(got["voxels"], env.container), ep
        lfs = env.level_free_space
        for z in range(H):
            for y in range(L):
                row = got["lists"][z, y]
                assert row[1:1 + row[0]].tolist() == lfs[z][y], (ep, z, y)
