// place_macs2d.cuh -- MACS ("maximal accessible convex space") placement step, 2D, one warp per
// environment, lane = column (and lane = candidate position, and lane = level for the tie-break).
//
// Replaces tools.calc_one_position_mcs_2d (tools.py:2456-2749).  State: heightmap h[W] plus the history
// (x, z, bx, bz) of every previous block INCLUDING unplaced ones, whose position stays (0,0)
// (tools.py:2531-2533).  The reference's per-level free-interval lists `level_free_space[z]` are, under
// the voxel invariant, the maximal runs of columns with h <= z, so they are derived, not stored.
//
//   EMS list A (:2518-2529): for each level z (z+bz <= H) every maximal run [x1,x2] of {h <= z} that is
//     not already a run of level z-1 (<=> contains a column with h == z), x1+bx <= W; order (z, x1).
//   EMS list B (:2531-2555): for every previous block i (x,z,xx,zz), t = z+zz < H: if row t is free over
//     the block's top -> [x, x+xx-1] unless an identical EMS exists; otherwise the "left" piece (cell x
//     free, x>0, cell x-1 free: extend right inside the top) and the "right" piece (symmetric); no
//     duplicate check for those.
//   Two candidates per EMS e (:2681-2700): 2e scans x = X1 .. W-bx upward, 2e+1 scans x = X2-bx+1 .. 0
//     downward; `visited` is shared, so a position settles at most one candidate.  A position p is
//     feasible only at level M[p] = max(h over footprint) (whole extent free + not floating, :2572-2580).
//   Score (:2590-2604) with the height typo `_z + block_x > height -> height = _z + block_z` kept.
//   Selection (:2708-2736): candidates with score == max (exact fp64); if several and 'mcs' in
//     reward_type: maximal usable space  sum_{lvl < maxH} (longest free run at lvl after the placement,
//     measured x2-x1), first maximum in candidate order; otherwise the first.
#pragma once
#include "tapenv_common.cuh"
#include "place_lbg2d.cuh"

namespace tapenv {

struct MacsHist {            // lane i holds previous block i (slot 0) and i+32 (slot 1)
    int x[2], z[2], xx[2], zz[2];
};

// longest run of set bits in m
__device__ __forceinline__ int longest_run(unsigned m) {
    int r = 0;
    while (m) { m &= m >> 1; ++r; }
    return r;
}

__device__ __forceinline__ PlaceOut macs2d_place(const DevCfg &c, int lane, int bx, int bz, int &h, Scal &sc,
                                                 const MacsHist &hist, unsigned *ems_keys, int &anomaly) {
    const int W = c.W, H = c.H, k = sc.k;
    const bool hard = (c.flags & TAPENV_RF_HARD) != 0;
    const bool posvalid = (bx >= 1) && (lane + bx <= W);
    const bool col = lane < W;
    const int bxc = min(bx, W);
    const unsigned wmask = W >= 32 ? 0xffffffffu : ((1u << W) - 1u);

    // ---- per-position feasibility (as LB_GREEDY 2D) ----
    const Foot2D f = foot2d_scan(lane, bxc, h);
    const int M = f.M;
    const bool stable_p = (M == 0) || (2 * f.first < bx && 2 * (bx - 1 - f.last) < bx);
    const bool ok_p = posvalid && (stable_p || !hard);
    const int add_p = bx * M - f.sumh;
    const unsigned ok_mask = __ballot_sync(TAPENV_FULL_MASK, ok_p);

    // ---- list A representatives: lane c stands for (z = h[c], the run of {h <= z} around c) if it is the
    //      leftmost column of that run with h == z ----
    unsigned le_mask = 0, eq_mask = 0;               // columns with h <= / == this lane's h
    for (int d = 0; d < W; ++d) {
        const int t = __shfl_sync(TAPENV_FULL_MASK, h, d);
        le_mask |= (t <= h ? 1u : 0u) << d;
        eq_mask |= (t == h ? 1u : 0u) << d;
    }
    // run of set bits of le_mask containing bit `lane`
    const unsigned below = ~le_mask & ((1u << lane) - 1u);                 // blocked columns left of lane
    const int a_x1 = below ? 32 - __clz(below) : 0;
    const unsigned above = (~le_mask & wmask) >> lane;                     // bit 0 is lane itself (clear)
    const int a_x2 = above ? lane + __ffs(above) - 2 : W - 1;
    const unsigned span_left = lane > a_x1 ? (((1u << lane) - 1u) & ~((1u << a_x1) - 1u)) : 0u;
    const bool a_rep = col && (eq_mask & span_left) == 0u && a_x1 + bx <= W && h + bz <= H;
    const unsigned a_key = ((unsigned)h << 5) | (unsigned)a_x1;            // order (z, x1)

    // ---- sequential candidate assignment ----
    unsigned settled = 0;
    int cidx = 0x7fffffff;                           // candidate index settled on this lane's position
    int ne = 0;                                      // EMS count so far
    const int X = W - bx + 1;

    auto process_ems = [&](int X1, int Z, int X2) {
        const int e = ne++;
        if (lane == 0 && e < kMaxEms) ems_keys[e] = ((unsigned)Z << 10) | ((unsigned)(X2 & 31) << 5) | (unsigned)(X1 & 31);
        const unsigned level = __ballot_sync(TAPENV_FULL_MASK, posvalid && M == Z) & ok_mask;
        if (X1 < X) {                                // candidate 2e: upward from X1 (:2683-2689)
            const unsigned a = level & ~settled & ~((1u << X1) - 1u);
            if (a) { const int p = __ffs(a) - 1; settled |= 1u << p; if (lane == p) cidx = 2 * e; }
        }
        if (X2 - bx + 2 > 0) {                       // candidate 2e+1: downward from X2-bx+1 (:2691-2698)
            const int hi = X2 - bx + 1;
            const unsigned d = level & ~settled & (hi >= 31 ? 0xffffffffu : ((2u << hi) - 1u));
            if (d) { const int p = 31 - __clz(d); settled |= 1u << p; if (lane == p) cidx = 2 * e + 1; }
        }
    };

    {   // list A in (z, x1) order
        unsigned left = __ballot_sync(TAPENV_FULL_MASK, a_rep);
        while (left) {
            const unsigned kmin = __reduce_min_sync(TAPENV_FULL_MASK, ((left >> lane) & 1u) ? a_key : 0xffffffffu);
            const int e = __ffs(__ballot_sync(TAPENV_FULL_MASK, ((left >> lane) & 1u) && a_key == kmin)) - 1;
            left &= ~(1u << e);
            const int X2 = __shfl_sync(TAPENV_FULL_MASK, a_x2, e);
            process_ems((int)(kmin & 31u), (int)(kmin >> 5), X2);
        }
    }
    __syncwarp();
    for (int i = 0; i < k && i < kMaxBlocks; ++i) {  // list B in block order (warp-uniform loop)
        const int slot = i >> 5, src = i & 31;
        const int x = __shfl_sync(TAPENV_FULL_MASK, slot ? hist.x[1] : hist.x[0], src);
        const int z = __shfl_sync(TAPENV_FULL_MASK, slot ? hist.z[1] : hist.z[0], src);
        const int xx = __shfl_sync(TAPENV_FULL_MASK, slot ? hist.xx[1] : hist.xx[0], src);
        const int zz = __shfl_sync(TAPENV_FULL_MASK, slot ? hist.zz[1] : hist.zz[0], src);
        const int t = z + zz;
        if (t >= H) continue;
        if (xx < 1 || x < 0 || x >= W) { anomaly |= 1; continue; }
        const unsigned freet = __ballot_sync(TAPENV_FULL_MASK, col && h <= t);
        const unsigned top = ((xx >= 32 ? 0xffffffffu : ((1u << xx) - 1u)) << x) & wmask;   // numpy clips the slice at the wall
        if ((freet & top) == top) {
            const unsigned key = ((unsigned)t << 10) | ((unsigned)((x + xx - 1) & 31) << 5) | (unsigned)x;
            bool dup = false;
            __syncwarp();
            for (int q = lane; q < ne && q < kMaxEms; q += 32) dup |= ems_keys[q] == key;
            if (!__any_sync(TAPENV_FULL_MASK, dup)) process_ems(x, t, x + xx - 1);
            __syncwarp();
        } else {
            if (x + xx > W) { anomaly |= 1; continue; }   // the reference indexes past the wall here (IndexError)
            if (((freet >> x) & 1u) && x > 0 && ((freet >> (x - 1)) & 1u)) {          // left piece (:2541-2548)
                const unsigned occ = x + 1 >= 32 ? 1u : (~freet >> (x + 1));          // first blocked column right of x
                const int run = occ ? __ffs(occ) - 1 : 31;
                process_ems(x, t, min(x + run, x + xx - 1));
                __syncwarp();
            }
            const int xr = x + xx - 1;
            if (((freet >> xr) & 1u) && x + xx < W && ((freet >> (xr + 1)) & 1u)) {   // right piece (:2549-2555)
                const unsigned occ = ~freet & ((1u << xr) - 1u);                      // blocked columns left of xr
                const int lo = occ ? 32 - __clz(occ) : 0;
                process_ems(max(lo, x), t, xr);
                __syncwarp();
            }
        }
    }
    if (ne > kMaxEms) anomaly |= 4;

    PlaceOut res;
    res.placed = 0; res.x = 0; res.y = 0; res.z = 0; res.stable = 0; res.top = 0;
    if (settled == 0) return res;                    // tools.py:2703-2706

    // ---- scores (:2590-2604) ----
    const int valid_new = sc.valid + bx * bz;
    const int hmax = warp_max(col ? h : 0);
    const bool mine = (settled >> lane) & 1u;
    const int top = M + bz;
    const int hnew = max(hmax, top);                 // np.max(heightmap_ems[index])
    int height = hnew;
    if (M + bx > height) height = top;               // sic (:2594)
    double score = cps_score(c.flags, valid_new, height * W, sc.empty + add_p, sc.nstable + (stable_p ? 1 : 0), k);
    if (c.flags & TAPENV_RF_MCS_START) score = 0.0;  // reward_type.startswith('mcs') (:2709-2710)
    unsigned bhi, blo;
    unsigned kbest = warp_argmax_first(mine, score, (unsigned)cidx, &bhi, &blo);
    const unsigned long long sbits = (unsigned long long)__double_as_longlong(score);
    const bool tied = mine && (unsigned)(sbits >> 32) == bhi && (unsigned)sbits == blo;
    unsigned tmask = __ballot_sync(TAPENV_FULL_MASK, tied);
    if ((c.flags & TAPENV_RF_MCS_IN) && __popc(tmask) > 1) {          // tie-break (:2718-2731)
        const int maxH = warp_max(mine ? hnew : 0);                   // np.max(heightmap_ems)
        if (maxH > H) anomaly |= 1;
        int my_mus = -1;
        for (unsigned tm = tmask; tm; tm &= tm - 1u) {                // one tied candidate at a time
            const int p = __ffs(tm) - 1;
            const int ptop = __shfl_sync(TAPENV_FULL_MASK, top, p);
            int total = 0;
            for (int base = 0; base < maxH; base += 32) {             // lane = level
                const int lvl = base + lane;
                unsigned fm = 0;
                for (int d = 0; d < W; ++d) {
                    int t = __shfl_sync(TAPENV_FULL_MASK, h, d);
                    if (d >= p && d < p + bx) t = ptop;
                    fm |= (t <= lvl ? 1u : 0u) << d;
                }
                const int run = longest_run(fm);
                total += (lvl < maxH && run > 0) ? run - 1 : 0;
            }
            total = warp_add(total);
            if (lane == p) my_mus = total;
        }
        // first maximum of mus in candidate order
        const int best_mus = warp_max(tied ? my_mus : -1);
        kbest = __reduce_min_sync(TAPENV_FULL_MASK, (tied && my_mus == best_mus) ? (unsigned)cidx : 0xffffffffu);
    }
    const int best = __ffs(__ballot_sync(TAPENV_FULL_MASK, mine && (unsigned)cidx == kbest)) - 1;
    const int zb = __shfl_sync(TAPENV_FULL_MASK, M, best);
    const int stb = __shfl_sync(TAPENV_FULL_MASK, stable_p ? 1 : 0, best);
    const int addb = __shfl_sync(TAPENV_FULL_MASK, add_p, best);

    if (lane >= best && lane < best + bx) h = zb + bz;               // :2747
    sc.valid = valid_new;
    sc.empty += addb;
    sc.nstable += stb;
    res.placed = 1; res.x = best; res.z = zb; res.stable = stb; res.top = zb + bz;
    return res;
}

}  // namespace tapenv
