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

// ONE_SLOT: the caller guarantees at most 32 previous blocks (compile-time shape with n <= 32): the second history slot and
// everything derived from it drop out of the generated code.
template <bool ONE_SLOT = false>
__device__ __forceinline__ PlaceOut macs2d_place(const DevCfg &c, int lane, int bx, int bz, int &h, Scal &sc,
                                                 const MacsHist &hist, unsigned *ems_keys, int &anomaly) {
    (void)ems_keys;                                  // (the EMS key list of the r01 formulation; no longer used)
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

    // ---- ONE sweep over the columns gives every lane all it needs from "the other columns", in three roles at once:
    //   lane = column c (list A):          le / eq masks of { h[d] <= / == h[c] }
    //   lane = history entry i (list B):   free[s] = { d : h[d] <= t_i }                      (t_i = top level of block i)
    //   lane = owner of an EMS at level Z: lv*  = { positions p : ok[p] and M[p] == Z }        (the `level` set of :2683-2698)
    // (r01 evaluated the list-B and level sets with one ballot per EMS inside the sequential loop.)
    const bool two = !ONE_SLOT && k > 32;            // second history slot in use (warp-uniform)
    const int t0 = hist.z[0] + hist.zz[0], t1 = hist.z[1] + hist.zz[1];
    unsigned le_mask = 0, eq_mask = 0, fr0 = 0, fr1 = 0, lvA = 0, lv0 = 0, lv1 = 0;
    for (int d = 0; d < W; ++d) {
        const int hd = __shfl_sync(TAPENV_FULL_MASK, h, d);
        const int Md = __shfl_sync(TAPENV_FULL_MASK, M, d);
        const unsigned bit = 1u << d;
        const bool okd = (ok_mask & bit) != 0u;
        if (hd <= h) le_mask |= bit;
        if (hd == h) eq_mask |= bit;
        if (hd <= t0) fr0 |= bit;
        if (okd && Md == h) lvA |= bit;
        if (okd && Md == t0) lv0 |= bit;
        if (two) {
            if (hd <= t1) fr1 |= bit;
            if (okd && Md == t1) lv1 |= bit;
        }
    }

    // ---- list A representatives: lane c stands for (z = h[c], the run of {h <= z} around c) if it is the
    //      leftmost column of that run with h == z ----
    const unsigned below = ~le_mask & ((1u << lane) - 1u);                 // blocked columns left of lane
    const int a_x1 = below ? 32 - __clz(below) : 0;
    const unsigned above = (~le_mask & wmask) >> lane;                     // bit 0 is lane itself (clear)
    const int a_x2 = above ? lane + __ffs(above) - 2 : W - 1;
    const unsigned span_left = lane > a_x1 ? (((1u << lane) - 1u) & ~((1u << a_x1) - 1u)) : 0u;
    const bool a_rep = col && (eq_mask & span_left) == 0u && a_x1 + bx <= W && h + bz <= H;
    const unsigned a_key = ((unsigned)h << 5) | (unsigned)a_x1;            // order (z, x1)
    // identity of an EMS for the duplicate test of list B (:2538): (level, X2, X1)
    const unsigned a_id = a_rep ? (((unsigned)h << 10) | ((unsigned)(a_x2 & 31) << 5) | (unsigned)(a_x1 & 31)) : 0xffffffffu;

    // ---- list B, lane-parallel (tools.py:2531-2555): entry i = previous block i, INCLUDING unplaced ones at (0,0) ----
    // per slot: up to two EMS (a = "full top" or "left piece", b = "right piece"), packed X1:5 | X2:6 | valid:1 each
    unsigned pk[2] = {0u, 0u}, ida[2] = {0xffffffffu, 0xffffffffu}, idb[2] = {0xffffffffu, 0xffffffffu};
    bool full[2] = {false, false};
    bool badent = false;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        if (s == 1 && !two) break;
        const int i = lane + 32 * s;
        const int x = hist.x[s], xx = hist.xx[s], t = s ? t1 : t0;
        const unsigned fr = s ? fr1 : fr0;
        const bool active = i < k && i < kMaxBlocks && t < H;
        const bool bad = active && (xx < 1 || x < 0 || x >= W);
        badent |= bad;
        if (!active || bad) continue;
        const unsigned top = ((xx >= 32 ? 0xffffffffu : ((1u << xx) - 1u)) << x) & wmask;   // numpy clips the slice at the wall
        if ((fr & top) == top) {                                         // whole top free -> [x, x+xx-1] unless already listed
            full[s] = true;
            const int X2 = x + xx - 1;
            pk[s] = (unsigned)x | ((unsigned)(X2 & 63) << 5) | (1u << 11);
            ida[s] = ((unsigned)t << 10) | ((unsigned)(X2 & 31) << 5) | (unsigned)x;
        } else if (x + xx > W) {
            badent = true;                                               // the reference indexes past the wall here (IndexError)
        } else {
            if (((fr >> x) & 1u) && x > 0 && ((fr >> (x - 1)) & 1u)) {   // left piece (:2541-2548)
                const unsigned occ = x + 1 >= 32 ? 1u : (~fr >> (x + 1));   // first blocked column right of x
                const int run = occ ? __ffs(occ) - 1 : 31;
                const int X2 = min(x + run, x + xx - 1);
                pk[s] = (unsigned)x | ((unsigned)(X2 & 63) << 5) | (1u << 11);
                ida[s] = ((unsigned)t << 10) | ((unsigned)(X2 & 31) << 5) | (unsigned)x;
            }
            const int xr = x + xx - 1;
            if (((fr >> xr) & 1u) && x + xx < W && ((fr >> (xr + 1)) & 1u)) {   // right piece (:2549-2555)
                const unsigned occ = ~fr & ((1u << xr) - 1u);            // blocked columns left of xr
                const int lo = occ ? 32 - __clz(occ) : 0;
                const int X1 = max(lo, x);
                pk[s] |= ((unsigned)X1 << 12) | ((unsigned)(xr & 63) << 17) | (1u << 23);
                idb[s] = ((unsigned)t << 10) | ((unsigned)(xr & 31) << 5) | (unsigned)(X1 & 31);
            }
        }
    }
    if (__any_sync(TAPENV_FULL_MASK, badent)) anomaly |= 1;
    // duplicate test of the "full top" case: an identical EMS listed EARLIER -- in list A, or from an earlier block (any piece;
    // an earlier identical full top that was itself dropped as a duplicate implies a still earlier identical one).
    const unsigned anyfull = __ballot_sync(TAPENV_FULL_MASK, full[0] || full[1]);
    if (anyfull) {
        const unsigned lower = (1u << lane) - 1u;
        // a-pieces of earlier blocks (full tops and left pieces share `ida`): one MATCH instead of a loop
        bool dup0 = (__match_any_sync(TAPENV_FULL_MASK, ida[0]) & lower) != 0u, dup1 = false;
        for (unsigned m = __ballot_sync(TAPENV_FULL_MASK, a_rep); m; m &= m - 1u) {          // list A
            const unsigned q = __shfl_sync(TAPENV_FULL_MASK, a_id, __ffs(m) - 1);
            dup0 |= q == ida[0]; dup1 |= q == ida[1];
        }
        for (unsigned m = __ballot_sync(TAPENV_FULL_MASK, idb[0] != 0xffffffffu); m; m &= m - 1u) {   // right pieces of earlier blocks
            const int j = __ffs(m) - 1;
            const unsigned q = __shfl_sync(TAPENV_FULL_MASK, idb[0], j);
            if (j < lane) dup0 |= q == ida[0];
            dup1 |= q == ida[1];                     // every slot-0 entry precedes every slot-1 entry
        }
        if (two) {
            dup1 |= (__match_any_sync(TAPENV_FULL_MASK, ida[1]) & lower) != 0u;
            for (int j = 0; j < 32; ++j) dup1 |= __shfl_sync(TAPENV_FULL_MASK, ida[0], j) == ida[1];
            for (unsigned m = __ballot_sync(TAPENV_FULL_MASK, idb[1] != 0xffffffffu); m; m &= m - 1u) {
                const int j = __ffs(m) - 1;
                const unsigned q = __shfl_sync(TAPENV_FULL_MASK, idb[1], j);
                if (j < lane) dup1 |= q == ida[1];
            }
        }
        if (full[0] && dup0) pk[0] = 0u;             // ida stays: later identical entries are duplicates of the same EMS
        if (full[1] && dup1) pk[1] = 0u;
    }

    // ---- sequential candidate assignment (shared `visited`): a few bit operations per EMS on warp-uniform values ----
    unsigned settled = 0;
    int cidx = 0x7fffffff;                           // candidate index settled on this lane's position
    const int X = W - bx + 1;

    auto process_ems = [&](int X1, unsigned level, int X2, int e) {
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

    int nA = 0;
    {   // list A in (z, x1) order
        unsigned left = __ballot_sync(TAPENV_FULL_MASK, a_rep);
        while (left) {
            const unsigned kmin = __reduce_min_sync(TAPENV_FULL_MASK, ((left >> lane) & 1u) ? a_key : 0xffffffffu);
            const int e = __ffs(__ballot_sync(TAPENV_FULL_MASK, ((left >> lane) & 1u) && a_key == kmin)) - 1;
            left &= ~(1u << e);
            const int X2 = __shfl_sync(TAPENV_FULL_MASK, a_x2, e);
            const unsigned level = __shfl_sync(TAPENV_FULL_MASK, lvA, e);
            process_ems((int)(kmin & 31u), level, X2, nA++);
        }
    }
    // list B in block order: only the entries that contribute an EMS
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        if (s == 1 && !two) break;
        for (unsigned m = __ballot_sync(TAPENV_FULL_MASK, pk[s] != 0u); m; m &= m - 1u) {
            const int src = __ffs(m) - 1;
            const unsigned w = __shfl_sync(TAPENV_FULL_MASK, pk[s], src);
            const unsigned level = __shfl_sync(TAPENV_FULL_MASK, s ? lv1 : lv0, src);
            const int e = 32 + 2 * (src + 32 * s);   // any numbering that is monotone in list order (list A: e < 32)
            if (w & (1u << 11)) process_ems((int)(w & 31u), level, (int)((w >> 5) & 63u), e);
            if (w & (1u << 23)) process_ems((int)((w >> 12) & 31u), level, (int)((w >> 17) & 63u), e + 1);
        }
    }

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
        // maximal usable space after the placement (:2667-2678): at level lvl the free columns are those of the CURRENT
        // heightmap outside the footprint plus -- iff the new top is <= lvl -- the footprint itself; lane = level
        int my_mus = tied ? 0 : -1;
        for (int base = 0; base < maxH; base += 32) {
            const int lvl = base + lane;
            unsigned fmb = 0;
            for (int d = 0; d < W; ++d) fmb |= (__shfl_sync(TAPENV_FULL_MASK, h, d) <= lvl ? 1u : 0u) << d;
            for (unsigned tm = tmask; tm; tm &= tm - 1u) {            // one tied candidate at a time
                const int p = __ffs(tm) - 1;
                const int ptop = __shfl_sync(TAPENV_FULL_MASK, top, p);
                const unsigned F = (((bx >= 32 ? 0xffffffffu : ((1u << bx) - 1u)) << p)) & wmask;
                const unsigned fm = (fmb & ~F) | (ptop <= lvl ? F : 0u);
                const int run = longest_run(fm);
                const int total = warp_add((lvl < maxH && run > 0) ? run - 1 : 0);
                if (lane == p) my_mus += total;
            }
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
