// place_lbg2d.cuh -- LB_GREEDY placement step, 2D, one warp per environment, lane = column.
//
// Replaces tools.calc_one_position_lb_greedy_2d (tools.py:2027-2176) together with is_stable_2d
// (tools.py:839-868).  The reference walks a voxel grid with Python lists; this is a from-scratch,
// position-centric formulation on the heightmap h[W] only (the voxel grid obeys
// cell(x,z) != 0  <=>  z < h[x]  after every step, SURVEY section 8a "state reduction"):
//
//   M[p]   = max(h[p .. p+bx-1])                (the only level at which position p can ever be a
//                                                candidate: at z < M[p] the row is blocked, at z > M[p]
//                                                the block would float, tools.py:2106-2109)
//   sup[p] = { c in footprint : h[c] == M[p] }  supporting columns
//   l, r   = leading / trailing footprint columns not in sup  -> is_stable_2d:
//            stable <=> M[p]==0  or  (2l < bx and 2r < bx)     (tools.py:862-866)
//   ok[p]  = stable[p] or not hard              (tools.py:2112-2114)
//   EMS    = { x : x==0 or h[x] != h[x-1] } with x+bx <= W, level M[x] (tools.py:2067-2078)
//
// Each EMS, in (level, x) order, settles on the first not-yet-settled ok position p >= x with
// M[p] == M[x] (the shared `visited` list of tools.py:2100 only ever matters for settled positions:
// every other visited position is rejected for a reason that does not depend on the EMS).  Without the
// hard constraint every EMS settles on its own corner.  The winner is the first maximum of (C+P)+S in
// EMS order.
#pragma once
#include "tapenv_common.cuh"

namespace tapenv {

// Footprint scan shared by the 2D strategies: window max M, first/last supporting offset, sum of heights.
struct Foot2D { int M, sumh, first, last; };

__device__ __forceinline__ Foot2D foot2d_scan(int lane, int bx, int h) {
    Foot2D f; f.M = h; f.sumh = h; f.first = 0; f.last = 0;
    for (int d = 1; d < bx; ++d) {
        const int t = __shfl_down_sync(TAPENV_FULL_MASK, h, d);   // lanes past the wall read junk: masked by posvalid
        f.sumh += t;
        if (t > f.M) { f.M = t; f.first = d; f.last = d; }
        else if (t == f.M) f.last = d;
    }
    return f;
}

// h: this lane's column height (lanes >= W must hold 0).  Updates h and sc in place.
__device__ __forceinline__ PlaceOut lbg2d_place(const DevCfg &c, int lane, int bx, int bz, int &h, Scal &sc) {
    const int W = c.W;
    const bool hard = (c.flags & TAPENV_RF_HARD) != 0;
    const bool posvalid = (bx >= 1) && (lane + bx <= W);
    const int bxc = min(bx, W);

    const Foot2D f = foot2d_scan(lane, bxc, h);
    const int M = f.M;
    const int l = f.first, r = bx - 1 - f.last;
    const bool stable_p = (M == 0) || (2 * l < bx && 2 * r < bx);
    const bool ok_p = posvalid && (stable_p || !hard);
    const int add_p = bx * M - f.sumh;               // empty cells created under the block (tools.py:2131-2133)

    const int hprev = __shfl_up_sync(TAPENV_FULL_MASK, h, 1);
    const bool ems_p = posvalid && (lane == 0 || h != hprev);
    const unsigned ems_mask = __ballot_sync(TAPENV_FULL_MASK, ems_p);

    unsigned taken = ems_mask;                       // soft: every EMS settles on its own corner
    if (hard) {
        const unsigned ok_mask = __ballot_sync(TAPENV_FULL_MASK, ok_p);
        unsigned em = ems_mask, consumed = 0;
        // EMS order is (level, x); within a level the settled positions are increasing in x, and EMS of
        // different levels never compete, so processing in x order per level == processing in list order.
        while (em) {                                 // warp-uniform, <= W iterations
            const int x = __ffs(em) - 1;
            em &= em - 1;
            const int z = __shfl_sync(TAPENV_FULL_MASK, M, x);
            const unsigned same = __ballot_sync(TAPENV_FULL_MASK, posvalid && M == z);
            const unsigned avail = ok_mask & same & ~consumed & ~((1u << x) - 1u);
            if (avail) consumed |= 1u << (__ffs(avail) - 1);
        }
        taken = consumed;
    }

    PlaceOut res;
    res.placed = 0; res.x = 0; res.y = 0; res.z = 0; res.stable = 0; res.top = 0;
    if (taken == 0) return res;                      // uniform

    const int valid_new = sc.valid + bx * bz;        // tools.py:2061
    const int hmax = warp_max(h);
    const bool mine = (taken >> lane) & 1u;
    const int top = M + bz;
    const int height = max(hmax, top);
    const double score = cps_score(c.flags, valid_new, height * W, sc.empty + add_p,
                                   sc.nstable + (stable_p ? 1 : 0), sc.k);
    const unsigned key = warp_argmax_first(mine, score, ((unsigned)M << 5) | (unsigned)lane);
    const int best = (int)(key & 31u);
    const int zb = (int)(key >> 5);
    const int stb = __shfl_sync(TAPENV_FULL_MASK, stable_p ? 1 : 0, best);
    const int addb = __shfl_sync(TAPENV_FULL_MASK, add_p, best);

    if (lane >= best && lane < best + bx) h = zb + bz;   // tools.py:2119, :2174
    sc.valid = valid_new;
    sc.empty += addb;
    sc.nstable += stb;
    res.placed = 1; res.x = best; res.z = zb; res.stable = stb; res.top = zb + bz;
    return res;
}

}  // namespace tapenv
