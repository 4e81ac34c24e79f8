// place_lbg3d.cuh -- LB_GREEDY placement step, 3D, one warp per environment, lane = heightmap cell
// (x = lane / L, y = lane % L, W*L <= 32).
//
// Replaces tools.calc_one_position_lb_greedy_3d (tools.py:2178-2351) and tools.is_stable
// (tools.py:710-765) on the heightmap h[W][L] only (voxel invariant: cell != 0 <=> z < h).
//
//   dx[x,y] = h[x,y]-h[x-1,y] (x>0), dy[x,y] = h[x,y]-h[x,y-1] (y>0)                     (:2219-2225)
//   EMS corners: (0,0); cells with dx != 0 unless (y>0, dx[x,y-1] != 0, h[x,y]==h[x,y-1],
//   dx[x,y]==dx[x,y-1]); cells with dy != 0 unless (x>0, dy[x-1,y] != 0, h[x,y]==h[x-1,y],
//   dx[x,y]==dx[x-1,y]  -- sic, hm_diff_x, :2243) and not already listed                 (:2228-2246)
//   order: stable sort by y, drop corners whose block would cross a wall (continue, not break),
//   z = max(h over footprint), stable sort by z                                          (:2249-2262)
//   => a corner's rank is (z, y, list group [origin | x-list | y-list], x).
//
// A position p is feasible only at level M[p] = max(h over its footprint); the shared `visited` list
// (:2283) matters only for positions some earlier EMS settled on.  Every EMS, in rank order, settles on
// the first free position of its scan rectangle [X0, W-bx] x [Y0, L-by] (x outer, y inner == lane
// order, :2324) with M[p] == z that is stable or (not hard).  Without `hard` that is always its own
// corner.  Winner = first maximum of (C+P)+S in EMS rank order (:2336-2337).
#pragma once
#include "tapenv_common.cuh"
#include "stable3d.cuh"

namespace tapenv {

__device__ __forceinline__ PlaceOut lbg3d_place(const DevCfg &c, int lane, int x, int y, int bx, int by, int bz,
                                                int &h, Scal &sc) {
    const int W = c.W, L = c.L, cells = W * L;
    const bool hard = (c.flags & TAPENV_RF_HARD) != 0;
    const bool cell = lane < cells;
    const bool posvalid = cell && bx >= 1 && by >= 1 && x + bx <= W && y + by <= L;
    const int bxc = min(bx, W), byc = min(by, L);

    // ---- EMS corner flags (neighbour values through shuffles) ----
    const int hu = __shfl_up_sync(TAPENV_FULL_MASK, h, L);         // (x-1, y)
    const int hl = __shfl_up_sync(TAPENV_FULL_MASK, h, 1);         // (x, y-1)
    const int dxv = (cell && x > 0) ? h - hu : 0;
    const int dyv = (cell && y > 0) ? h - hl : 0;
    const int dx_l = __shfl_up_sync(TAPENV_FULL_MASK, dxv, 1);     // dx[x, y-1]
    const int dx_u = __shfl_up_sync(TAPENV_FULL_MASK, dxv, L);     // dx[x-1, y]
    const int dy_u = __shfl_up_sync(TAPENV_FULL_MASK, dyv, L);     // dy[x-1, y]
    const bool keptx = dxv != 0 && !(y != 0 && dx_l != 0 && h == hl && dxv == dx_l);
    const bool kepty = dyv != 0 && !(x != 0 && dy_u != 0 && h == hu && dxv == dx_u);
    const bool is_ems = posvalid && (lane == 0 || keptx || kepty);
    const unsigned group = lane == 0 ? 0u : (keptx ? 1u : 2u);

    // ---- footprint scan: level M, supporting cells (x-major bit i*by+j), sum of heights ----
    int M = -1, sumh = 0;
    unsigned sup = 0;
    for (int i = 0; i < bxc; ++i) {
        for (int j = 0; j < byc; ++j) {
            const int v = __shfl_sync(TAPENV_FULL_MASK, h, (lane + i * L + j) & 31);
            const unsigned bit = 1u << ((i * by + j) & 31);
            sumh += v;
            if (v > M) { M = v; sup = bit; }
            else if (v == M) sup |= bit;
        }
    }
    const unsigned ems_mask = __ballot_sync(TAPENV_FULL_MASK, is_ems);
    // is_stable is needed for every position a hard scan may visit, otherwise only for the corners
    const bool need_st = hard ? posvalid : is_ems;
    bool stable_p = false;
    if (need_st) stable_p = (M == 0) || stable3d_from_support(bx, by, sup);
    const unsigned key_own = ((unsigned)M << 12) | ((unsigned)y << 7) | (group << 5) | (unsigned)x;

    unsigned taken = ems_mask;                       // soft: every EMS settles on its own corner
    unsigned key = key_own;
    if (hard) {
        const unsigned ok_mask = __ballot_sync(TAPENV_FULL_MASK, posvalid && stable_p);
        unsigned left = ems_mask, settled = 0;
        key = 0xffffffffu;
        while (left) {                               // warp-uniform, <= W*L iterations, EMS in rank order
            const unsigned kmin = __reduce_min_sync(TAPENV_FULL_MASK, ((left >> lane) & 1u) ? key_own : 0xffffffffu);
            const int e = __ffs(__ballot_sync(TAPENV_FULL_MASK, ((left >> lane) & 1u) && key_own == kmin)) - 1;
            left &= ~(1u << e);
            const int z = (int)(kmin >> 12);
            const int X0 = __shfl_sync(TAPENV_FULL_MASK, x, e), Y0 = __shfl_sync(TAPENV_FULL_MASK, y, e);
            const unsigned avail = __ballot_sync(TAPENV_FULL_MASK, posvalid && x >= X0 && y >= Y0 && M == z) & ok_mask & ~settled;
            if (avail) {
                const int p = __ffs(avail) - 1;      // x outer, y inner == ascending lane
                settled |= 1u << p;
                if (lane == p) key = kmin;
            }
        }
        taken = settled;
    }

    PlaceOut res;
    res.placed = 0; res.x = 0; res.y = 0; res.z = 0; res.stable = 0; res.top = 0;
    if (taken == 0) return res;                      // uniform (tools.py:2265-2268, :2330-2333)

    const int valid_new = sc.valid + bx * by * bz;   // tools.py:2212
    const int hmax = warp_max(cell ? h : 0);
    const bool mine = (taken >> lane) & 1u;
    const int top = M + bz;
    const int height = max(hmax, top);
    const int add_p = bx * by * M - sumh;            // empty cells created under the block (tools.py:2307-2309)
    const double score = cps_score(c.flags, valid_new, height * W * L, sc.empty + add_p,
                                   sc.nstable + (stable_p ? 1 : 0), sc.k);
    const unsigned kbest = warp_argmax_first(mine, score, key);
    const int best = __ffs(__ballot_sync(TAPENV_FULL_MASK, mine && key == kbest)) - 1;
    const int zb = __shfl_sync(TAPENV_FULL_MASK, M, best);
    const int stb = __shfl_sync(TAPENV_FULL_MASK, stable_p ? 1 : 0, best);
    const int addb = __shfl_sync(TAPENV_FULL_MASK, add_p, best);
    const int px = __shfl_sync(TAPENV_FULL_MASK, x, best), py = __shfl_sync(TAPENV_FULL_MASK, y, best);

    if (cell && x >= px && x < px + bx && y >= py && y < py + by) h = zb + bz;   // tools.py:2342-2343
    sc.valid = valid_new;
    sc.empty += addb;
    sc.nstable += stb;
    res.placed = 1; res.x = px; res.y = py; res.z = zb; res.stable = stb; res.top = zb + bz;
    return res;
}

}  // namespace tapenv
