// place_macs3d.cuh -- the MACS placement strategy in 3D (packing_strategy='MACS'/'MUL' or a 'C+P+S-mcs-*' / 'C+P+S-mul-*'
// reward with a 3D container; calc_one_position_mcs_3d, tools.py:2751-3165), one WARP per environment.
//
// Like LB (place_lb.cuh) and unlike LB_GREEDY / MACS-2D this strategy cannot be reduced to the heightmap: its EMS
// scan compares voxel VALUES (block ids) between neighbouring rows (:2926, :2933) and it keeps per-(level, row)
// interval lists that are edited incrementally and drift away from the true empty runs (probe: 322 of 2400 steps).
// The state therefore carries the voxel grid (int16: 0 empty, -1 empty under a block, k+1 block id) and the lists
// (int8 rows: byte 0 = length).
//
// r02 (second form).  The first r02 form let lane 0 of the warp walk the grid literally (97 000 warp-instructions per
// environment and step, every voxel test a dependent global load: 2.1 ms per decode step at B=4096).  Here the warp first
// turns the grid into LEVEL MASKS in shared memory -- emp[z] / pos[z]: bit (y*W + x) <=> voxel (x,y,z) == 0 / > 0, lane =
// level, coalesced reads; W*L <= 32 cells, so a whole level is one word -- and every test the reference makes on voxels
// becomes a bit operation on one or two words:
//   (container[xa:xb, y, z] == 0).all()          (emp[z] & rowmask) == rowmask
//   floating / free / support of a footprint     emp[Z-1], AND of emp[Z..Z+bz), pos[Z-1] against the footprint mask
//   empty cells under the footprint (calc_C_P_S) sum over levels of popc(emp[t] & foot)          -- lane = level
//   np.max(heightmap) with the block written in  warp max over the cells outside the footprint   -- lane = cell
//   calc_maximal_usable_spaces                   per level a run-length scan of L bit rows        -- lane = level
// The control flow (EMS construction, the corner scans with their shared `visited` list) stays the reference's, executed
// warp-UNIFORMLY on values every lane holds identically; the lanes split the reductions above.  The candidates tying for
// the best score are recorded during the walk, so the 'mcs' tie-break no longer re-walks the EMS list.  Only the commit
// (voxel / list edits) is left to lane 0.
//
// Reference behaviour kept on purpose: the read of the stale loop variable `x1` (:2869), `z` instead of `z+zz` in the
// EMS found on top of a partly covered block (:2939-2940), the `_z + block_x` height typo (:2976), phantom (0,0,0)
// positions of unplaced blocks in the neighbour scan, first-maximum choices, Python slice clipping (an empty slice is
// "all zero").
//
// The header also compiles for the host (g++, one "lane"): tests/test_macs3d_host_cpu.py drives it against the oracle.
#pragma once
#include <stdint.h>
#include "../../include/tapenv.h"
#include "stable3d.cuh"
#if defined(__CUDACC__)
#include "tapenv_common.cuh"
#endif

namespace tapenv {

constexpr int kM3MaxEms = 192;       // EMS per step (observed <= 44 with 50 blocks in 5x5x250); overflow -> flag 8
constexpr int kM3MaxLevels = 64;     // distinct z levels carrying visited positions (observed <= 23); overflow -> flag 8
constexpr int kM3MaxH = 256;         // container height (tapenv_config_check: <= 255 for this strategy)
constexpr int kM3MaxTies = 256;      // candidates sharing the best score; overflow -> flag 8

struct M3Ems { unsigned char x1, y1, z, x2, y2, z2; };

struct M3Scratch {                   // per warp, shared memory (4.9 kB)
    unsigned emp[kM3MaxH];           // bit (y*W + x): voxel (x, y, z) == 0
    unsigned pos[kM3MaxH];           // bit (y*W + x): voxel (x, y, z) > 0 (a block; -1 = empty under a block is neither)
    unsigned ties[kM3MaxTies];       // x:5 | y:5 | z:8 | stable:1 | empty cells under the footprint:13
    unsigned vmask[kM3MaxLevels];    // the shared `visited` list (:2952): per level a mask over start positions
    unsigned fmask[kM3MaxLevels];    // AND of emp[Z .. Z+bz): cells free over the whole block height
    unsigned chg[kM3MaxH / 32];      // bit z: level_free_space[z] differs from level z-1 (:2813-2815)
    short zkey[kM3MaxLevels];
    M3Ems ems[kM3MaxEms];
};

// ---- the warp vocabulary (device: 32 lanes; host build: one lane that loops over everything) ----
TAPENV_HD int m3_wsum(int v) {
#if defined(__CUDA_ARCH__)
    return __reduce_add_sync(0xffffffffu, v);
#else
    return v;
#endif
}
TAPENV_HD int m3_wmax(int v) {
#if defined(__CUDA_ARCH__)
    return __reduce_max_sync(0xffffffffu, v);
#else
    return v;
#endif
}
TAPENV_HD void m3_sync() {
#if defined(__CUDA_ARCH__)
    __syncwarp();
#endif
}
TAPENV_HD unsigned m3_bits(int n) { return n >= 32 ? 0xffffffffu : (n <= 0 ? 0u : ((1u << n) - 1u)); }

struct M3State {
    short *vox;            // [cells][H]
    signed char *lists;    // [H][L][lcap]: byte 0 = length, then x1,x2,x1,x2,...
    int *h;                // [cells]
    int W, L, H, cells, lcap;
    M3Scratch *sm;
    int lane, nl;          // this lane, lanes cooperating (32 / 1)
    TAPENV_HD short &v(int x, int y, int z) const { return vox[(x * L + y) * H + z]; }
    TAPENV_HD signed char *list(int z, int y) const { return lists + (size_t)(z * L + y) * lcap; }
    TAPENV_HD bool isz(int x, int y, int z) const { return (sm->emp[z] >> (y * W + x)) & 1u; }
};

// ---- Python list semantics on an int8 row ----
TAPENV_HD int m3_index(const signed char *l, int v) { for (int i = 1; i <= l[0]; ++i) if (l[i] == v) return i - 1; return -1; }
TAPENV_HD void m3_remove(signed char *l, int v) {
    for (int i = 1; i <= l[0]; ++i) if (l[i] == v) { for (int q = i; q < l[0]; ++q) l[q] = l[q + 1]; --l[0]; return; }
}
TAPENV_HD bool m3_eq(const signed char *a, const signed char *b) {
    if (a[0] != b[0]) return false;
    for (int i = 1; i <= a[0]; ++i) if (a[i] != b[i]) return false;
    return true;
}
TAPENV_HD void m3_sort(signed char *l) {
    for (int i = 2; i <= l[0]; ++i) { const signed char t = l[i]; int j = i - 1; while (j >= 1 && l[j] > t) { l[j + 1] = l[j]; --j; } l[j + 1] = t; }
}

// cells [xa, xb) of row y as a level mask; the x slice follows Python: negative bounds count from the wall, the stop is clipped
TAPENV_HD unsigned m3_rowmask(const M3State &s, int xa, int xb, int y) {
    if (xa < 0) { xa += s.W; if (xa < 0) xa = 0; }
    if (xb < 0) { xb += s.W; if (xb < 0) xb = 0; }
    if (xb > s.W) xb = s.W;
    if (xb <= xa) return 0u;
    return (m3_bits(xb - xa) << xa) << (y * s.W);
}
// (container[xa:xb, y, z] == 0).all()
TAPENV_HD bool m3_row_free(const M3State &s, int xa, int xb, int y, int z) {
    const unsigned m = m3_rowmask(s, xa, xb, y);
    return (s.sm->emp[z] & m) == m;
}
// footprint of a bx x by block at (x, y); x + bx <= W, y + by <= L
TAPENV_HD unsigned m3_foot(const M3State &s, int x, int y, int bx, int by) {
    const unsigned row = m3_bits(bx) << x;
    unsigned m = 0u;
    for (int j = 0; j < by; ++j) m |= row << ((y + j) * s.W);
    return m;
}

// emp / pos for the levels below zlim (nothing above the pile is ever examined), and the levels whose interval lists differ
// from the level below (above the pile every list is pristine): lane = level
TAPENV_HD void m3_build_masks(const M3State &s, int zlim) {
    M3Scratch &sm = *s.sm;
    for (int i = s.lane; i < kM3MaxH / 32; i += s.nl) sm.chg[i] = 0u;
    m3_sync();
    for (int z0 = 0; z0 < zlim; z0 += s.nl) {
        const int z = z0 + s.lane;
        bool changed = false;
        if (z < zlim) {
            unsigned e = 0u, p = 0u;
            for (int x = 0; x < s.W; ++x)
                for (int y = 0; y < s.L; ++y) {
                    const int val = s.v(x, y, z);
                    const unsigned bit = 1u << (y * s.W + x);
                    if (val == 0) e |= bit;
                    if (val > 0) p |= bit;
                }
            sm.emp[z] = e; sm.pos[z] = p;
            changed = z == 0;
            if (z > 0) for (int y = 0; y < s.L && !changed; ++y) changed = !m3_eq(s.list(z - 1, y), s.list(z, y));
        }
#if defined(__CUDA_ARCH__)
        const unsigned w = __ballot_sync(0xffffffffu, changed);
        if (s.lane == 0) sm.chg[z0 >> 5] = w;
#else
        if (changed) sm.chg[z >> 5] |= 1u << (z & 31);
#endif
    }
    m3_sync();
}

// EMS list :2810-2940.  Returns the number of entries (entries beyond kM3MaxEms are dropped and flagged).
TAPENV_HD int m3_build_ems(const M3State &s, int k, const int *positions, const int *blocks, int bx, int by, int bz, int &anomaly) {
    const int W = s.W, L = s.L, H = s.H;
    M3Ems *ems = s.sm->ems;
    int ne = 0;
    // the list lives in shared memory: lane 0 writes an entry, every lane counts it; readers synchronise first
    auto push = [&](int a, int b, int c, int d, int e, int f) {
        if (ne < kM3MaxEms) {
            if (s.lane == 0) { ems[ne].x1 = (unsigned char)a; ems[ne].y1 = (unsigned char)b; ems[ne].z = (unsigned char)c;
                               ems[ne].x2 = (unsigned char)d; ems[ne].y2 = (unsigned char)e; ems[ne].z2 = (unsigned char)f; }
            ++ne;
        } else anomaly |= 8;
    };
    auto listed = [&](int a, int b, int c, int d, int e, int f) {
        m3_sync();
        for (int i = 0; i < ne; ++i)
            if (ems[i].x1 == a && ems[i].y1 == b && ems[i].z == c && ems[i].x2 == d && ems[i].y2 == e && ems[i].z2 == f) return true;
        return false;
    };
    // Python function-scope loop variables that outlive their loops
    int x1 = 0, x2 = 0, y1 = 0, y2 = 0;
    bool x1_bound = false;

    // ---- from level_free_space (:2810-2838): only the levels whose lists differ from the level below ----
    for (int zw = 0; zw * 32 < H; ++zw) {
        bool stop = false;
        for (unsigned m = s.sm->chg[zw]; m; m &= m - 1u) {
            const int z = zw * 32 + tap_ctz(m);
            if (z + bz > H) { stop = true; break; }
            for (int y = 0; y < L; ++y) {
                const signed char *fs = s.list(z, y);
                if (y + by > L) break;
                if (y > 0 && m3_eq(s.list(z, y - 1), fs)) continue;
                for (int sidx = 1; sidx + 1 <= fs[0]; sidx += 2) {
                    x1 = fs[sidx]; x2 = fs[sidx + 1]; x1_bound = true;
                    if (x1 + bx > W) break;
                    if (y > 0) {
                        const signed char *lo = s.list(z, y - 1);
                        const int idx = m3_index(lo, x1) + 1;                      // 0 when absent
                        if (idx > 0 && idx % 2 == 1 && idx < lo[0] && x2 == lo[idx + 1]) continue;
                    }
                    if (z > 0) {
                        const signed char *lo = s.list(z - 1, y);
                        const int idx = m3_index(lo, x1) + 1;
                        if (idx > 0 && idx % 2 == 1 && idx < lo[0] && x2 == lo[idx + 1]) continue;
                    }
                    bool xspace = true;
                    for (y2 = y; y2 < L; ++y2) {
                        if (y2 == L - 1) break;
                        if (!m3_row_free(s, x1, x2 + 1, y2 + 1, z)) break;
                        const signed char *nx = s.list(z, y2 + 1);
                        if (xspace && !(m3_index(nx, x1) >= 0 && m3_index(nx, x2) >= 0)) { xspace = false; push(x1, y, z, x2, y2, z); }
                    }
                    push(x1, y, z, x2, y2, z);
                }
            }
        }
        if (stop) break;                                   // z + bz > H holds for every higher level too
    }
    // ---- next to the settled blocks (:2841-2940); unplaced blocks sit at their phantom (0,0,0) ----
    for (int b = 0; b < k; ++b) {
        const int x = positions[b * 3], y = positions[b * 3 + 1], z = positions[b * 3 + 2];
        const int xx = blocks[b * 3], yy = blocks[b * 3 + 1], zz = blocks[b * 3 + 2];
        if (z >= H || x + xx > W || y + yy > L) { anomaly |= 1; return ne; }         // the reference raises IndexError
        if (y + yy < L) {                                                           // upon along the y axis
            const int ya = y + yy;
            if (m3_row_free(s, x, x + xx, ya, z)) {
                if ((x > 0 && s.isz(x - 1, ya, z)) || (x + xx < W && s.isz(x + xx, ya, z))) {
                    for (y2 = ya; y2 < L; ++y2) { if (y2 == L - 1) break; if (!m3_row_free(s, x, x + xx, y2 + 1, z)) break; }
                    push(x, ya, z, x + xx - 1, y2, z);
                }
            } else {
                if (s.isz(x, ya, z) && x > 0 && s.isz(x - 1, ya, z)) {               // left
                    for (x2 = x; x2 < x + xx; ++x2) { if (x2 == W - 1) break; if (!s.isz(x2 + 1, ya, z)) break; }
                    if (x2 == x + xx) x2 = x + xx - 1;
                    if (!x1_bound) { anomaly |= 1; return ne; }                      // UnboundLocalError in the reference
                    for (y2 = ya; y2 < L; ++y2) { if (y2 == L - 1) break; if (!m3_row_free(s, x1 /* stale, :2869 */, x2 + 1, y2 + 1, z)) break; }
                    push(x, ya, z, x2, y2, z);
                }
                if (s.isz(x + xx - 1, ya, z) && x + xx < W && s.isz(x + xx, ya, z)) {  // right
                    for (x1 = x + xx - 1; x1 >= x; --x1) { if (x1 == 0) break; if (!s.isz(x1 - 1, ya, z)) break; }
                    if (x1 < x) x1 = x;
                    x1_bound = true;
                    for (y2 = ya; y2 < L; ++y2) { if (y2 == L - 1) break; if (!m3_row_free(s, x1, x + xx, y2 + 1, z)) break; }
                    push(x1, ya, z, x + xx - 1, y2, z);
                }
            }
        }
        if (y > 0) {                                                                // under along the y axis
            const int yb = y - 1;
            if (m3_row_free(s, x, x + xx, yb, z)) {
                if ((x > 0 && s.isz(x - 1, yb, z)) || (x + xx < W && s.isz(x + xx, yb, z))) {
                    for (y1 = yb; y1 >= 0; --y1) { if (y1 == 0) break; if (!m3_row_free(s, x, x + xx, y1 - 1, z)) break; }
                    push(x, y1, z, x + xx - 1, yb, z);
                }
            } else {
                if (s.isz(x, yb, z) && x > 0 && s.isz(x - 1, yb, z)) {               // left
                    for (x2 = x; x2 < x + xx; ++x2) { if (x2 == W - 1) break; if (!s.isz(x2 + 1, yb, z)) break; }
                    if (x2 == x + xx) x2 = x + xx - 1;
                    for (y1 = yb; y1 >= 0; --y1) { if (y1 == 0) break; if (!m3_row_free(s, x, x2 + 1, y1 - 1, z)) break; }
                    push(x, y1, z, x2, yb, z);
                }
                if (s.isz(x + xx - 1, yb, z) && x + xx < W && s.isz(x + xx, yb, z)) {  // right
                    for (x1 = x + xx - 1; x1 >= x; --x1) { if (x1 == 0) break; if (!s.isz(x1 - 1, yb, z)) break; }
                    if (x1 < x) x1 = x;
                    x1_bound = true;
                    for (y1 = yb; y1 >= 0; --y1) { if (y1 == 0) break; if (!m3_row_free(s, x1, x + xx, y1 - 1, z)) break; }
                    push(x1, y1, z, x + xx - 1, yb, z);
                }
            }
        }
        if (z + zz < H) {                                                           // on top
            const int t = z + zz;
            const unsigned top = m3_foot(s, x, y, xx, yy);
            if ((s.sm->emp[t] & top) == top) {
                if (!listed(x, y, t, x + xx - 1, y + yy - 1, t)) push(x, y, t, x + xx - 1, y + yy - 1, t);
            } else {
                // histogram of free run lengths along +x over the block's top face (:2915-2922): hist(i,j) = number of
                // consecutive free cells x+i, x+i+1, ... (inside the face) of row y+j = trailing ones of the row's bits
                const unsigned et = s.sm->emp[t], rb = m3_bits(xx);
                auto hist = [&](int i, int j) { const unsigned r = ~(((et >> ((y + j) * W + x)) & rb) >> i); return r ? tap_ctz(r) : 32; };
                for (int i = 0; i < xx; ++i)
                    for (int j = 0; j < yy; ++j) {
                        const int hij = hist(i, j);
                        if (hij == 0) continue;
                        if (j > 0 && hij == hist(i, j - 1)) continue;
                        if (i > 0) { bool eq = true; for (int r = y + j; r < y + yy && eq; ++r) eq = s.v(x + i, r, t) == s.v(x + i - 1, r, t); if (eq) continue; }
                        const int i2 = i + hij - 1;
                        int j2, j1;
                        for (j2 = j; j2 < yy; ++j2) { if (j2 == yy - 1) break; if (hist(i, j2 + 1) < hij) break; }
                        if (i > 0) { bool eq = true; for (int r = y + j; r < y + j2 && eq; ++r) eq = s.v(x + i, r, t) == s.v(x + i - 1, r, t); if (eq) continue; }   // empty slice when j2 == j
                        for (j1 = j; j1 >= 0; --j1) { if (j1 == 0) break; if (hist(i, j1 - 1) < hij) break; }
                        if (!listed(x + i, y + j1, z, x + i2, y + j2, z)) push(x + i, y + j1, z /* sic :2939 */, x + i2, y + j2, z);
                    }
            }
        }
    }
    return ne;
}

// calc_maximal_usable_spaces (:3052-3080) of the container with the candidate block written in (update_container
// :3046-3050): a cell is empty iff it is empty now and not inside / under the candidate.  Per level the reference builds
// hist[i][j] = length of the run of empty cells starting at (i,j) along x and takes the largest "hist value x maximal
// y-extent with hist >= that value"; a level is L bit rows of the level mask, hist a count of trailing ones.  lane = level.
TAPENV_HD int m3_usable(const M3State &s, unsigned foot, int ctop, int hlim) {
    const int W = s.W, L = s.L;
    const unsigned rb = m3_bits(W);
    int acc = 0;
    for (int hh = s.lane; hh < hlim; hh += s.nl) {
        unsigned lvl = s.sm->emp[hh];
        if (hh < ctop) lvl &= ~foot;
        auto hist = [&](int i, int j) { const unsigned r = ~(((lvl >> (j * W)) & rb) >> i); return r ? tap_ctz(r) : 32; };
        int level_max = 0;
        for (int i = 0; i < W; ++i)
            for (int j = 0; j < L; ++j) {
                const int hij = hist(i, j);
                if (hij == 0) continue;
                if (j > 0 && hij == hist(i, j - 1)) continue;
                int j2, j1;
                for (j2 = j; j2 < L; ++j2) { if (j2 == L - 1) break; if (hist(i, j2 + 1) < hij) break; }
                for (j1 = j; j1 >= 0; --j1) { if (j1 == 0) break; if (hist(i, j1 - 1) < hij) break; }
                const int area = hij * (j2 - j1 + 1);
                if (area > level_max) level_max = area;
            }
        acc += level_max;
    }
    return m3_wsum(acc);
}

// IEEE fp64 C+P+S exactly as the reference evaluates it (tapenv_common.cuh cps_score; plain divisions on the host)
TAPENV_HD double m3_score(int flags, int valid_new, int bbox, int empty_new, int stable_cnt, int k) {
#if defined(__CUDA_ARCH__)
    return cps_score(flags, valid_new, bbox, empty_new, stable_cnt, k);
#else
    const double vd = (double)valid_new;
    const double c = vd / (double)bbox;
    const double p = (flags & TAPENV_RF_P) ? vd / (double)(empty_new + valid_new) : 0.0;
    const double sq = (flags & TAPENV_RF_S) ? (double)stable_cnt / (double)(k + 1) : 0.0;
    return (c + p) + sq;
#endif
}

struct M3Best { bool any; int x, y, z, stable, add; };

// One block for one environment, the whole warp.  hc: the heightmap cell this lane owns (lane = x*L + y; 0 beyond the
// cells; ignored by the host build, which reads s.h).  Returns the chosen placement (any == false: not placed).
TAPENV_HD M3Best macs3d_place(int flags, const M3State &s, int k, const int *positions, const int *blocks,
                              int bx, int by, int bz, int valid_new, int empty, int nstable, int hc, int &anomaly) {
    const int W = s.W, L = s.L, H = s.H;
    M3Scratch &sm = *s.sm;
    const bool hard = (flags & TAPENV_RF_HARD) != 0;
    const bool mcs_start = (flags & TAPENV_RF_MCS_START) != 0, mcs_in = (flags & TAPENV_RF_MCS_IN) != 0;
    M3Best best; best.any = false; best.x = best.y = best.z = best.stable = best.add = 0;
#if defined(__CUDA_ARCH__)
    const int cx = s.lane / L, cy = s.lane - cx * L;     // this lane's heightmap cell
    const int hmax0 = m3_wmax(hc);
#else
    int hmax0 = 0;
    for (int i = 0; i < s.cells; ++i) hmax0 = hmax0 > s.h[i] ? hmax0 : s.h[i];
    (void)hc;
#endif
    // the masks are needed up to the highest level any scan can touch: the pile (or the phantom top of an unplaced block,
    // which sits at the origin with its own height) plus the new block
    int top = hmax0;
    for (int i = 0; i < k; ++i) { const int t = positions[i * 3 + 2] + blocks[i * 3 + 2]; top = top > t ? top : t; }
    m3_build_masks(s, top + bz + 1 < H ? top + bz + 1 : H);
    const int ne = m3_build_ems(s, k, positions, blocks, bx, by, bz, anomaly);
    if (anomaly & 1) return best;
    m3_sync();                                           // the EMS list is complete and visible to every lane
    const int X = W - bx + 1, Y = L - by + 1;
    const unsigned fpat = m3_foot(s, 0, 0, bx, by);     // footprint at the origin: at (x, y) it is fpat << (y*W + x)
    const unsigned invW = (65536u + (unsigned)W - 1u) / (unsigned)W;   // c / W for c < 32 with one multiply
    int nlev = 0;

    double best_score = 0.0;     // np.max(ratio_ems): never-settled entries are 0.0
    int nties = 0, nsettled = 0, max_height = 0;
    M3Best first; first.any = false; first.x = first.y = first.z = first.stable = first.add = 0;

    for (int ei = 0; ei < ne; ++ei) {
        const int X1 = sm.ems[ei].x1, Y1 = sm.ems[ei].y1, Z = sm.ems[ei].z, X2 = sm.ems[ei].x2, Y2 = sm.ems[ei].y2;
        const int xr = X2 - bx + 2, yr = Y2 - by + 2;
        // the level's slot of the shared `visited` list + the cells free over the whole block height at this level
        // (slots are written by lane 0 at the end of an EMS and read by every lane at the start of a later one)
        m3_sync();
        int vslot = -1;
        for (int i = 0; i < nlev; ++i) if (sm.zkey[i] == Z) { vslot = i; break; }
        unsigned vis, freeZ;
        if (vslot < 0) {
            if (nlev >= kM3MaxLevels) { anomaly |= 8; vslot = kM3MaxLevels - 1; }
            else vslot = nlev++;
            freeZ = 0xffffffffu;
            for (int t = Z; t < Z + bz && t < H; ++t) freeZ &= sm.emp[t];
            vis = 0u;
        } else { vis = sm.vmask[vslot]; freeZ = sm.fmask[vslot]; }
        const unsigned below = Z > 0 ? sm.emp[Z - 1] : 0u, sup_lvl = Z > 0 ? sm.pos[Z - 1] : 0u;
        for (int corner = 0; corner < 4; ++corner) {
            bool ok;
            if (corner == 0) ok = X1 < X && Y1 < Y;
            else if (corner == 1) ok = xr > 0 && Y1 < Y;
            else if (corner == 2) ok = xr > 0 && yr > 0;
            else ok = X1 < X && yr > 0;
            if (!ok) continue;
            max_height = max_height > hmax0 ? max_height : hmax0;                  // heightmap.copy() (:3091)
            // itertools.product order: corner 0 (x asc, y asc), 1 (y asc, x desc), 2 (x desc, y desc), 3 (y desc, x asc)
            const int na = corner == 0 ? X - X1 : (corner == 1 ? Y - Y1 : (corner == 2 ? xr : yr));
            const int nb = corner == 0 ? Y - Y1 : (corner == 1 ? xr : (corner == 2 ? yr : X - X1));
            bool settled = false, st = false;
            int px = 0, py = 0;
            unsigned pfoot = 0u;
            for (int ia = 0; ia < na && !settled; ++ia)
                for (int ib = 0; ib < nb && !settled; ++ib) {
                    int _x, _y;
                    if (corner == 0) { _x = X1 + ia; _y = Y1 + ib; }
                    else if (corner == 1) { _y = Y1 + ia; _x = xr - 1 - ib; }
                    else if (corner == 2) { _x = xr - 1 - ia; _y = yr - 1 - ib; }
                    else { _y = yr - 1 - ia; _x = X1 + ib; }
                    if (_x < 0 || _y < 0 || _x + bx > W || _y + by > L) { anomaly |= 1; continue; }
                    const int org = _y * W + _x;
                    const unsigned bit = 1u << org;
                    if (vis & bit) continue;                                               // :2956
                    const unsigned foot = fpat << org;
                    if (Z > 0 && (below & foot) == foot) continue;                          // floating: skipped, NOT marked (:2957)
                    vis |= bit;
                    if ((freeZ & foot) != foot) continue;
                    bool stable = true;
                    if (Z > 0) {
                        // is_stable (:710-765): the count rules need no geometry; otherwise the supporting cells in the
                        // footprint's own x-major numbering
                        const unsigned under = sup_lvl & foot;
                        const int cnt = tap_popc(under);
                        if (2 * cnt > bx * by) stable = true;
                        else if (cnt <= 1) stable = false;
                        else {
                            unsigned sup = 0u;
                            for (unsigned m = under; m; m &= m - 1u) {
                                const int cbit = tap_ctz(m);
                                const int yy = (int)(((unsigned)cbit * invW) >> 16), xx = cbit - yy * W;
                                sup |= 1u << (((xx - _x) * by + (yy - _y)) & 31);
                            }
                            stable = stable3d_from_support(bx, by, sup);
                        }
                    }
                    if (!stable && hard) continue;
                    settled = true; st = stable; px = _x; py = _y; pfoot = foot;
                }
            if (!settled) continue;
            // calc_C_P_S (:2971-2987)
#if defined(__CUDA_ARCH__)
            const bool infoot = cx >= px && cx < px + bx && cy >= py && cy < py + by;
            int height = m3_wmax(infoot ? 0 : hc);
#else
            int height = 0;
            for (int q = 0; q < W; ++q) for (int r = 0; r < L; ++r)
                if (!(q >= px && q < px + bx && r >= py && r < py + by)) height = height > s.h[q * L + r] ? height : s.h[q * L + r];
#endif
            height = height > Z + bz ? height : Z + bz;
            const int hm_max = height;
            if (Z + bx > height) height = Z + bz;                                          // sic :2976
            int part = 0;
            for (int t = s.lane; t < Z && t < H; t += s.nl) part += tap_popc(sm.emp[t] & pfoot);
            const int cnt = m3_wsum(part);
            const double ratio = mcs_start ? 0.0 : m3_score(flags, valid_new, height * W * L, empty + cnt, nstable + (st ? 1 : 0), k);
            ++nsettled;
            max_height = max_height > hm_max ? max_height : hm_max;
            if (!first.any || ratio > best_score) {
                nties = 0;
                best_score = ratio;
                first.any = true; first.x = px; first.y = py; first.z = Z; first.stable = st ? 1 : 0; first.add = cnt;
            }
            if (ratio == best_score) {
                if (nties >= kM3MaxTies) anomaly |= 8;
                else if (s.lane == 0) sm.ties[nties] = (unsigned)px | ((unsigned)py << 5) | ((unsigned)Z << 10) | ((st ? 1u : 0u) << 18) | ((unsigned)cnt << 19);
                ++nties;
            }
        }
        m3_sync();                                       // every lane has read this level's slot
        if (s.lane == 0) { sm.zkey[vslot] = (short)Z; sm.vmask[vslot] = vis; sm.fmask[vslot] = freeZ; }
    }
    if (nsettled == 0) return best;                                                    // :3129-3132
    const int count_best = mcs_start ? ne * 4 : nties;                                 // mcs*: every entry of ratio_ems is 0.0
    const bool tie = count_best > 1 && mcs_in;                                         // :3143
    if (!tie) return first;
    if (max_height > H) { anomaly |= 1; return best; }                                 // ctn[:, :, h] raises
    m3_sync();
    long long best_mus = -1;
    const int nt = nties < kM3MaxTies ? nties : kM3MaxTies;
    for (int i = 0; i < nt; ++i) {                                                     // first maximum in walk order (:3143-3154)
        const unsigned t = sm.ties[i];
        const int px = (int)(t & 31u), py = (int)((t >> 5) & 31u), Z = (int)((t >> 10) & 255u);
        const int mus = m3_usable(s, m3_foot(s, px, py, bx, by), Z + bz, max_height);
        if (!best.any || mus > best_mus) {
            best.any = true; best.x = px; best.y = py; best.z = Z; best.stable = (int)((t >> 18) & 1u); best.add = (int)(t >> 19);
            best_mus = mus;
        }
    }
    return best;
}

// commit (:3161-3171): update_container, update_level_free_space, heightmap -- ONE lane
TAPENV_HD void macs3d_commit(const M3State &s, int k, const M3Best &b, int bx, int by, int bz, int &anomaly) {
    if (b.z + bz > s.H) { anomaly |= 1; return; }               // level_free_space[_z+bz-1] raises IndexError
    const int _x = b.x, xx = b.x + bx - 1;
    for (int q = b.x; q < b.x + bx; ++q) for (int r = b.y; r < b.y + by; ++r) {
        for (int t = b.z; t < b.z + bz; ++t) s.v(q, r, t) = (short)(k + 1);
        for (int t = 0; t < b.z; ++t) if (s.v(q, r, t) == 0) s.v(q, r, t) = -1;
        s.h[q * s.L + r] = b.z + bz;
    }
    for (int t = b.z; t < b.z + bz; ++t) for (int r = b.y; r < b.y + by; ++r) {                // :3000-3024
        signed char *fs = s.list(t, r);
        const int idx = m3_index(fs, _x);
        if (idx >= 0) {
            if ((idx + 1) % 2 == 1) {
                if (m3_index(fs, xx) >= 0) {
                    if (bx == 1) {
                        if (idx + 1 < fs[0] && fs[idx + 2] == _x) { m3_remove(fs, _x); m3_remove(fs, _x); }
                        else fs[idx + 1] = (signed char)(_x + 1);
                    } else { m3_remove(fs, _x); m3_remove(fs, xx); }
                } else fs[idx + 1] = (signed char)(xx + 1);
            } else fs[idx + 1] = (signed char)(_x - 1);
        } else {
            const int ix = m3_index(fs, xx);
            if (ix >= 0) fs[ix + 1] = (signed char)(_x - 1);
            else if (fs[0] + 2 < s.lcap) { fs[fs[0] + 1] = (signed char)(_x - 1); fs[fs[0] + 2] = (signed char)(xx + 1); fs[0] += 2; m3_sort(fs); }
            else anomaly |= 8;
        }
    }
    for (int t = 0; t < b.z; ++t) for (int r = b.y; r < b.y + by; ++r) {                        // :3026-3042
        signed char *fs = s.list(t, r);
        signed char snap[40];
        const int n = fs[0] < 39 ? fs[0] : 39;
        for (int i = 0; i <= n; ++i) snap[i] = fs[i];
        for (int sidx = 1; sidx + 1 <= n; sidx += 2) {
            const int x1 = snap[sidx], x2 = snap[sidx + 1];
            if (x1 == x2) {
                if (x1 >= _x && x1 <= xx) { m3_remove(fs, x1); m3_remove(fs, x1); }
            } else if (bx == 1) {
                if (_x == x1) { const int i = m3_index(fs, x1); if (i >= 0) fs[i + 1] = (signed char)(_x + 1); }
                else if (_x == x2) { const int i = m3_index(fs, x2); if (i >= 0) fs[i + 1] = (signed char)(xx - 1); }
            } else if (_x <= x1 && x2 <= xx) { m3_remove(fs, x1); m3_remove(fs, x2); }
            else if (_x <= x1 && x1 <= xx) { const int i = m3_index(fs, x1); if (i >= 0) fs[i + 1] = (signed char)(xx + 1); }
            else if (_x <= x2 && x2 <= xx) { const int i = m3_index(fs, x2); if (i >= 0) fs[i + 1] = (signed char)(_x - 1); }
        }
    }
}

// Container.add_new_block for one environment (tools.py:3663-3744 with the MACS strategy in 3D), the whole warp.
// scal: (valid, empty, #stable, k) of the environment; positions / blks / stable: its [cap] rows.  Returns the anomaly bits.
TAPENV_HD int macs3d_env_add_block(int flags, int cap, const M3State &s, int *scal, int *positions, int *blks,
                                   unsigned char *stable_out, int bx, int by, int bz) {
    const int s0 = scal[0], s1 = scal[1], s2 = scal[2], k = scal[3];
    int anomaly = 0;
    if (k >= cap) return 2;
#if defined(__CUDA_ARCH__)
    const int hc = s.lane < s.cells ? s.h[s.lane] : 0;
#else
    const int hc = 0;
#endif
    m3_sync();                                           // every lane has read the state before lane 0 edits it
    int o0 = s0, o1 = s1, o2 = s2;
    unsigned char stable = 0;
    M3Best best; best.any = false;
    if (bx >= 1 && by >= 1 && bz >= 1 && bx * by <= 32) {
        const int vol = bx * by * bz;
        best = macs3d_place(flags, s, k, positions, blks, bx, by, bz, s0 + vol, s1, s2, hc, anomaly);
        if (best.any && !(anomaly & 1)) {
            m3_sync();
            if (s.lane == 0) {
                int a2 = 0;
                macs3d_commit(s, k, best, bx, by, bz, a2);
                if (!(a2 & 1)) { positions[k * 3] = best.x; positions[k * 3 + 1] = best.y; positions[k * 3 + 2] = best.z; }
                s.sm->chg[0] = (unsigned)a2;             // hand lane 0's verdict to the warp
            }
            m3_sync();
            const int a2 = (int)s.sm->chg[0];
            anomaly |= a2;
            if (!(a2 & 1)) { stable = (unsigned char)best.stable; o0 = s0 + vol; o1 = s1 + best.add; o2 = s2 + best.stable; }
        }
    }
    if (s.lane == 0) {
        blks[k * 3] = bx; blks[k * 3 + 1] = by; blks[k * 3 + 2] = bz;
        stable_out[k] = stable;
        scal[0] = o0; scal[1] = o1; scal[2] = o2; scal[3] = k + 1;
    }
    return anomaly;
}

}  // namespace tapenv
