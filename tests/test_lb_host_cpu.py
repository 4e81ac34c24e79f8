"""CPU: the LB placement (packing_strategy='LB', tools.py:1602-1914) as the kernels run it -- tap-net_b200/csrc/place_lb.cuh,
warp form (level masks, one ballot per EMS scan) compiled for the HOST with one emulated lane -- against the oracle's literal
restatement after every step, and against the one-thread walk of the same header (state compared byte for byte, lists
included).  The lane-parallel halves are covered on the GPU (tests/test_gpu_parity.py: lb trajectories, fuzz, mul, rolling)."""
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
    lib.tapenv_host_lb_episode.argtypes = [C.c_int] * 8 + [C.c_void_p, C.c_int] + [C.c_void_p] * 6
    lib.tapenv_host_lb_episode.restype = C.c_int
    return lib


def reward_flags(rt):
    return (1 if rt.endswith("hard") else 0) | (2 if "P" in rt else 0) | (4 if "S" in rt else 0) | (8 if "mcs" in rt else 0) | \
           (16 if rt.startswith("mcs") else 0)


def host_episode(lib, dim, W, L, H, n, rt, blocks, warp_form):
    cap, lcap = n, max(n + 2, W + 4)
    cells = W * L
    blocks = np.ascontiguousarray(blocks, dtype=np.int32)
    hm = np.zeros((n, cells), np.int32); sc = np.zeros((n, 4), np.int32)
    pos = np.zeros((cap, dim), np.int32); stb = np.zeros(cap, np.uint8)
    vox = np.zeros(cells * H, np.int16); lists = np.zeros(H * L * lcap, np.uint8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    anomaly = lib.tapenv_host_lb_episode(dim, W, L, H, cap, lcap, reward_flags(rt), n, p(blocks), warp_form, p(hm), p(sc), p(pos), p(stb), p(vox), p(lists))
    return dict(anomaly=anomaly, heightmap=hm, scal=sc, positions=pos, stable=stb, voxels=vox, lists=lists)


CASES = [(2, 5, 1, 50, 10, "C+P+S-lb-soft"), (2, 5, 1, 50, 10, "C+P+S-lb-hard"), (2, 16, 1, 120, 24, "C+P-lb-soft"), (2, 32, 1, 200, 40, "C+P+S-lb-hard"),
         (2, 2, 1, 60, 12, "C+P+S-lb-soft"), (3, 5, 5, 50, 10, "C+P+S-lb-soft"), (3, 5, 5, 50, 10, "C+P+S-lb-hard"), (3, 4, 6, 80, 16, "C+P-lb-hard"),
         (3, 6, 4, 80, 16, "C+P+S-lb-soft"), (3, 3, 3, 40, 10, "C+P+S-lb-hard"), (3, 8, 4, 120, 24, "C+P+S-lb-soft"), (3, 5, 5, 250, 50, "C+P+S-lb-soft")]


@pytest.mark.parametrize("dim,W,L,H,n,rt", CASES)
def test_host_build_matches_oracle(hostlib, dim, W, L, H, n, rt):
    rng = np.random.RandomState(dim * 10000 + W * 100 + L * 10 + n)
    size = [W, H] if dim == 2 else [W, L, H]
    for ep in range(50 if n <= 24 else 10):
        hi = min(5, W) + 1
        if dim == 2:
            blocks = np.stack([rng.randint(1, hi, size=n), rng.randint(1, 6, size=n)], 1)
        else:
            blocks = np.stack([rng.randint(1, hi, size=n), rng.randint(1, min(5, L) + 1, size=n), rng.randint(1, 6, size=n)], 1)
        env = oracle.Container(size, n, rt, "full", packing_strategy="LB")
        want_h, want_s = [], []
        failed = False
        for t in range(n):
            try:
                env.add_new_block(blocks[t].astype(np.float32))
            except IndexError:
                failed = True
                break
            want_h.append(env.heightmap.reshape(-1).copy())
            want_s.append((env.valid_size, env.empty_size, sum(env.stable[:t + 1]), t + 1))
        got = host_episode(hostlib, dim, W, L, H, n, rt, blocks, 1)
        walk = host_episode(hostlib, dim, W, L, H, n, rt, blocks, 0)
        for key in ("anomaly", "heightmap", "scal", "positions", "stable", "voxels", "lists"):
            assert np.array_equal(got[key], walk[key]), (ep, key)
        if failed:
            assert got["anomaly"] != 0
            continue
        assert got["anomaly"] == 0, (ep, got["anomaly"])
        assert np.array_equal(got["heightmap"], np.stack(want_h)), ep
        assert np.array_equal(got["scal"], np.asarray(want_s)), ep
        assert np.array_equal(got["positions"], env.positions), ep
        assert [bool(v) for v in got["stable"]] == env.stable, ep
        assert np.array_equal(got["voxels"].reshape(env.container.shape), env.container), ep
