// stable3d.cuh -- the 3D support test tools.is_stable (tools.py:710-765) as a pure function of the
// block footprint (bx, by) and the set of SUPPORTING footprint cells, usable from device and host code
// (the host build exists only so the CPU test-suite can check it exhaustively against the oracle).
//
// sup: bit (i*by + j) set  <=>  footprint cell (x+i, y+j) touches the block's bottom face
//      (container[x+i][y+j][z-1] > 0, tools.py:722-728).  x-major, the enumeration order of the reference.
// All geometry is evaluated in DOUBLED integer coordinates (cell (i,j) -> (2i,2j), block centre ->
// (bx-1, by-1)), so every comparison the reference makes in floating point is exact here.
//
//   z == 0                           -> stable                       (:715)   [handled by the caller]
//   #sup > bx*by/2                   -> stable                       (:730)
//   #sup <= 1                        -> unstable                     (:732)
//   #sup == 2                        -> two-point rule               (:736-744)
//   all collinear (Qhull raises)     -> two-point rule on the first point of minimal x and the first
//                                       point of maximal x           (:750-762)
//   otherwise                        -> centre inside ConvexHull, judged by matplotlib's
//                                       Path.contains_point          (:749, :764-765)
//
// The last rule is evaluated without building the hull.  For the counter-clockwise hull Qhull returns,
// the crossing test of matplotlib's point_in_path_impl (half-open rule `vy >= ty`) reduces to
//      ymin < ty <= ymax   and   xl(ty) <= tx <= xr(ty)
// where [xl, xr] is the hull's cross-section at height ty.  xl / xr are attained on segments joining a
// support point with y >= ty to one with y < ty, so
//      inside  <=>  exists (a,b): x_ab <= tx   and   exists (a,b): x_ab >= tx,
//      a in {y >= ty}, b in {y < ty}, x_ab = intersection of segment ab with the line y = ty,
// each a sign test of an integer cross product.  (This hull branch is the one place where the
// reference's behaviour comes from un-vendored third-party code: see DESIGN.md "parity unpinned".)
#pragma once
#if defined(__CUDACC__)
#define TAPENV_HD __host__ __device__ __forceinline__
#else
#define TAPENV_HD inline
#endif

namespace tapenv {

TAPENV_HD int tap_popc(unsigned v) {
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
TAPENV_HD int tap_ctz(unsigned v) {   // v != 0
#if defined(__CUDA_ARCH__)
    return __ffs((int)v) - 1;
#else
    return __builtin_ctz(v);
#endif
}
TAPENV_HD int tap_fls(unsigned v) {   // index of the highest set bit, v != 0
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int)v);
#else
    return 31 - __builtin_clz(v);
#endif
}

// tools.py:736-744 / :754-762 with a = cx-p0x, b = cy-p0y, c = cx-p1x, d = cy-p1y (doubled, signs and
// ratios unchanged).  a/b == c/d in IEEE double <=> a*d == c*b for these small integers.
TAPENV_HD bool two_point_rule(int a, int b, int c, int d) {
    if (b == 0 || d == 0) {
        if (b != d) return false;
        return (a < 0) != (c < 0);
    }
    return a * d == c * b && ((a < 0) != (c < 0)) && ((b < 0) != (d < 0));
}

// q / by for 0 <= q < 32, 1 <= by <= 32 with one multiply: inv = ceil(65536 / by) is exact in that range
#define TAPENV_DIVBY(q) ((int)(((unsigned)(q) * inv) >> 16))

TAPENV_HD bool stable3d_from_support(int bx, int by, unsigned sup) {
    const int cnt = tap_popc(sup);
    if (2 * cnt > bx * by) return true;
    if (cnt <= 1) return false;
    const unsigned inv = (65536u + (unsigned)by - 1u) / (unsigned)by;
    const int cx = bx - 1, cy = by - 1;                       // doubled centre
    const int b0 = tap_ctz(sup);
    const int i0 = TAPENV_DIVBY(b0), j0 = b0 - i0 * by;       // first point (x-major order)
    if (cnt == 2) {
        const int b1 = tap_fls(sup);
        const int i1 = TAPENV_DIVBY(b1), j1 = b1 - i1 * by;
        return two_point_rule(cx - 2 * i0, cy - 2 * j0, cx - 2 * i1, cy - 2 * j1);
    }
    // >= 3 points: collinear?
    unsigned rest = sup & (sup - 1u);
    const int bs = tap_ctz(rest);
    const int is = TAPENV_DIVBY(bs), js = bs - is * by;       // second point
    rest &= rest - 1u;
    bool collinear = true;
    for (unsigned m = rest; m; m &= m - 1u) {
        const int bb = tap_ctz(m);
        const int ii = TAPENV_DIVBY(bb), jj = bb - ii * by;
        if ((is - i0) * (jj - j0) - (js - j0) * (ii - i0) != 0) { collinear = false; break; }
    }
    if (collinear) {
        // np.argmin / np.argmax over x: first occurrence.  x-major order: the first point has minimal x;
        // the first point of maximal x is the lowest set bit of the highest occupied x-row.
        const int imax = TAPENV_DIVBY(tap_fls(sup));
        const unsigned rowmask = (by >= 32 ? 0xffffffffu : ((1u << by) - 1u)) << (imax * by);
        const int bm = tap_ctz(sup & rowmask);
        const int jm = bm - imax * by;
        return two_point_rule(cx - 2 * i0, cy - 2 * j0, cx - 2 * imax, cy - 2 * jm);
    }
    // hull cross-section at ty = cy (doubled): A = points with 2j >= cy, B = points with 2j < cy
    bool le = false, ge = false, anyA = false, anyB = false;
    for (unsigned ma = sup; ma; ma &= ma - 1u) {
        const int ba = tap_ctz(ma);
        const int ia = TAPENV_DIVBY(ba), ja = ba - ia * by;
        if (2 * ja < cy) continue;
        anyA = true;
        for (unsigned mb = sup; mb; mb &= mb - 1u) {
            const int bb = tap_ctz(mb);
            const int ib = TAPENV_DIVBY(bb), jb = bb - ib * by;
            if (2 * jb >= cy) continue;
            anyB = true;
            // x_ab <= tx  <=>  (ty-ay)(bx-ax) >= (tx-ax)(by-ay)   [by-ay < 0], doubled coordinates
            const int lhs = (cy - 2 * ja) * (2 * ib - 2 * ia);
            const int rhs = (cx - 2 * ia) * (2 * jb - 2 * ja);
            le |= lhs >= rhs;
            ge |= lhs <= rhs;
        }
    }
    return anyA && anyB && le && ge;
}

}  // namespace tapenv
