"""CPU: the integer restatement of tools.is_stable that the 3D kernel uses (tap-net_b200/csrc/stable3d.cuh,
compiled for the host by g++) against the oracle's literal hull + crossing-test version, exhaustively for every
footprint up to 20 cells and on a large sample above (compiled limit: 32 cells)."""
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
    subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-o", out, src])
    return C.CDLL(out)


def ours(lib, bx, by, masks):
    masks = np.ascontiguousarray(masks, dtype=np.uint32)
    out = np.zeros(len(masks), np.uint8)
    lib.tapenv_host_stable3d_masks(bx, by, masks.ctypes.data_as(C.c_void_p), len(masks), out.ctypes.data_as(C.c_void_p))
    return out


def test_stable3d_matches_oracle_everywhere(hostlib):
    rng = np.random.RandomState(0)
    checked = 0
    for bx in range(1, 17):
        for by in range(1, 17):
            nb = bx * by
            if nb > 32:
                continue
            if nb <= 18:
                m = np.arange(1 << nb, dtype=np.uint32)
            else:
                a = rng.randint(0, 2 ** 32, size=200000, dtype=np.uint64)
                b = rng.randint(0, 2 ** 32, size=200000, dtype=np.uint64)
                c = rng.randint(0, 2 ** 32, size=200000, dtype=np.uint64)
                dens = rng.randint(0, 3, size=200000)
                m = np.where(dens == 0, a, np.where(dens == 1, a & b, a & b & c))
                m = (m & np.uint64((1 << nb) - 1)).astype(np.uint32)
            got, want = ours(hostlib, bx, by, m), oracle.is_stable_3d_masks(bx, by, m)
            bad = np.nonzero(got != want)[0]
            assert len(bad) == 0, (bx, by, [bin(int(v)) for v in m[bad[:4]]])
            checked += len(m)
    assert checked > 5_000_000


def test_two_point_and_collinear_quirks(hostlib):
    """SURVEY Q10 [probe]: under a 3x3 block, supports (1,0),(1,2) -> False; (0,1),(2,1) -> True; (0,0),(2,2) -> True;
    (0,0),(1,0) -> False; three collinear supports sharing x -> False."""
    def mask(pts, by=3):
        return sum(1 << (x * by + y) for x, y in pts)
    cases = [([(1, 0), (1, 2)], 0), ([(0, 1), (2, 1)], 1), ([(0, 0), (2, 2)], 1), ([(0, 0), (1, 0)], 0),
             ([(1, 0), (1, 1), (1, 2)], 0), ([(0, 1), (1, 1), (2, 1)], 1)]
    m = np.array([mask(p) for p, _ in cases], np.uint32)
    want = np.array([w for _, w in cases], np.uint8)
    assert np.array_equal(oracle.is_stable_3d_masks(3, 3, m), want)
    assert np.array_equal(ours(hostlib, 3, 3, m), want)
