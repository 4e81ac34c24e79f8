// place_macs3d.cuh -- the MACS placement strategy in 3D (packing_strategy='MACS'/'MUL' or a 'C+P+S-mcs-*' / 'C+P+S-mul-*'
// reward with a 3D container; calc_one_position_mcs_3d, tools.py:2751-3165).
//
// Like LB (place_lb.cuh) and unlike LB_GREEDY / MACS-2D this strategy cannot be reduced to the heightmap: its EMS
// scan compares voxel VALUES (block ids) between neighbouring rows (:2926, :2933) and it keeps per-(level, row)
// interval lists that are edited incrementally and drift away from the true empty runs (probe: 322 of 2400 steps).
// The state therefore carries the voxel grid (int16: 0 empty, -1 empty under a block, k+1 block id) and the lists
// (int8 rows: byte 0 = length), and the algorithm -- a sequential walk over levels, rows, previous blocks and EMS
// corners with a shared `visited` list -- runs ONE THREAD per environment.  No BASELINE configuration uses it.
//
// Reference behaviour kept on purpose: the read of the stale loop variable `x1` (:2869), `z` instead of `z+zz` in the
// EMS found on top of a partly covered block (:2939-2940), the `_z + block_x` height typo (:2976), phantom (0,0,0)
// positions of unplaced blocks in the neighbour scan, first-maximum choices, Python slice clipping (an empty slice is
// "all zero").  Candidates are evaluated in two passes instead of being stored: pass 1 finds the best score, the number
// of candidates sharing it and np.max(heightmap_ems); pass 2 (only when the 'mcs' tie-break applies, :3143-3154)
// re-walks them -- the walk is deterministic -- and scores the tied ones by calc_maximal_usable_spaces.
#pragma once
#include "tapenv_common.cuh"
#include "stable3d.cuh"

namespace tapenv {

constexpr int kM3MaxEms = 192;       // EMS per step (observed <= 44 with 50 blocks in 5x5x250); overflow -> flag 8
constexpr int kM3MaxLevels = 64;     // distinct z levels carrying visited positions (observed <= 23); overflow -> flag 8

struct M3State {
    short *vox;            // [cells][H]
    signed char *lists;    // [H][L][lcap]: byte 0 = length, then x1,x2,x1,x2,...
    int *h;                // [cells]
    int W, L, H, cells, lcap;
    __device__ __forceinline__ short &v(int x, int y, int z) const { return vox[(x * L + y) * H + z]; }
    __device__ __forceinline__ signed char *list(int z, int y) const { return lists + (size_t)(z * L + y) * lcap; }
};

// ---- Python list semantics on an int8 row ----
__device__ __forceinline__ int m3_index(const signed char *l, int v) { for (int i = 1; i <= l[0]; ++i) if (l[i] == v) return i - 1; return -1; }
__device__ __forceinline__ void m3_remove(signed char *l, int v) {
    for (int i = 1; i <= l[0]; ++i) if (l[i] == v) { for (int q = i; q < l[0]; ++q) l[q] = l[q + 1]; --l[0]; return; }
}
__device__ __forceinline__ bool m3_eq(const signed char *a, const signed char *b) {
    if (a[0] != b[0]) return false;
    for (int i = 1; i <= a[0]; ++i) if (a[i] != b[i]) return false;
    return true;
}
__device__ __forceinline__ void m3_sort(signed char *l) {
    for (int i = 2; i <= l[0]; ++i) { const signed char t = l[i]; int j = i - 1; while (j >= 1 && l[j] > t) { l[j + 1] = l[j]; --j; } l[j + 1] = t; }
}

// (container[xa:xb, y, z] == 0).all(), x slice clipped like NumPy does
__device__ __forceinline__ bool m3_row_free(const M3State &s, int xa, int xb, int y, int z) {
    if (xb > s.W) xb = s.W;
    for (int x = xa; x < xb; ++x) if (s.v(x, y, z) != 0) return false;
    return true;
}

struct M3Ems { unsigned char x1, y1, z, x2, y2, z2; };

struct M3Best { bool any; int x, y, z, stable, add; double score; long long mus; };

struct M3Visited {           // the shared `visited` list (:2952): per level a bit mask over (x, y) start positions
    short zkey[kM3MaxLevels];
    unsigned mask[kM3MaxLevels];
    int n;
    __device__ __forceinline__ void clear() { n = 0; }
    // returns true if (x,y,z) was already visited; otherwise leaves it unvisited (the caller marks it later)
    __device__ __forceinline__ int slot(int z, int &anomaly) {
        for (int i = 0; i < n; ++i) if (zkey[i] == z) return i;
        if (n >= kM3MaxLevels) { anomaly |= 8; return kM3MaxLevels - 1; }
        zkey[n] = (short)z; mask[n] = 0u;
        return n++;
    }
};

// EMS list :2810-2940.  Returns the number of entries (entries beyond kM3MaxEms are dropped and flagged).
__device__ __forceinline__ int m3_build_ems(const M3State &s, int k, const int *positions, const int *blocks, int bx, int by, int bz,
                                            M3Ems *ems, int &anomaly) {
    const int W = s.W, L = s.L, H = s.H;
    int ne = 0;
    auto push = [&](int a, int b, int c, int d, int e, int f) {
        if (ne < kM3MaxEms) { ems[ne].x1 = (unsigned char)a; ems[ne].y1 = (unsigned char)b; ems[ne].z = (unsigned char)c;
                              ems[ne].x2 = (unsigned char)d; ems[ne].y2 = (unsigned char)e; ems[ne].z2 = (unsigned char)f; ++ne; }
        else anomaly |= 8;
    };
    auto listed = [&](int a, int b, int c, int d, int e, int f) {
        for (int i = 0; i < ne; ++i)
            if (ems[i].x1 == a && ems[i].y1 == b && ems[i].z == c && ems[i].x2 == d && ems[i].y2 == e && ems[i].z2 == f) return true;
        return false;
    };
    // Python function-scope loop variables that outlive their loops
    int x1 = 0, x2 = 0, y1 = 0, y2 = 0;
    bool x1_bound = false;

    // ---- from level_free_space (:2810-2838) ----
    for (int z = 0; z < H; ++z) {
        if (z + bz > H) break;
        if (z > 0) {
            bool same = true;
            for (int y = 0; y < L && same; ++y) same = m3_eq(s.list(z - 1, y), s.list(z, y));
            if (same) continue;
        }
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
    // ---- next to the settled blocks (:2841-2940); unplaced blocks sit at their phantom (0,0,0) ----
    for (int b = 0; b < k; ++b) {
        const int x = positions[b * 3], y = positions[b * 3 + 1], z = positions[b * 3 + 2];
        const int xx = blocks[b * 3], yy = blocks[b * 3 + 1], zz = blocks[b * 3 + 2];
        if (z >= H || x + xx > W || y + yy > L) { anomaly |= 1; return ne; }         // the reference raises IndexError
        if (y + yy < L) {                                                           // upon along the y axis
            const int ya = y + yy;
            if (m3_row_free(s, x, x + xx, ya, z)) {
                if ((x > 0 && s.v(x - 1, ya, z) == 0) || (x + xx < W && s.v(x + xx, ya, z) == 0)) {
                    for (y2 = ya; y2 < L; ++y2) { if (y2 == L - 1) break; if (!m3_row_free(s, x, x + xx, y2 + 1, z)) break; }
                    push(x, ya, z, x + xx - 1, y2, z);
                }
            } else {
                if (s.v(x, ya, z) == 0 && x > 0 && s.v(x - 1, ya, z) == 0) {         // left
                    for (x2 = x; x2 < x + xx; ++x2) { if (x2 == W - 1) break; if (s.v(x2 + 1, ya, z) != 0) break; }
                    if (x2 == x + xx) x2 = x + xx - 1;
                    if (!x1_bound) { anomaly |= 1; return ne; }                      // UnboundLocalError in the reference
                    for (y2 = ya; y2 < L; ++y2) { if (y2 == L - 1) break; if (!m3_row_free(s, x1 /* stale, :2869 */, x2 + 1, y2 + 1, z)) break; }
                    push(x, ya, z, x2, y2, z);
                }
                if (s.v(x + xx - 1, ya, z) == 0 && x + xx < W && s.v(x + xx, ya, z) == 0) {   // right
                    for (x1 = x + xx - 1; x1 >= x; --x1) { if (x1 == 0) break; if (s.v(x1 - 1, ya, z) != 0) break; }
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
                if ((x > 0 && s.v(x - 1, yb, z) == 0) || (x + xx < W && s.v(x + xx, yb, z) == 0)) {
                    for (y1 = yb; y1 >= 0; --y1) { if (y1 == 0) break; if (!m3_row_free(s, x, x + xx, y1 - 1, z)) break; }
                    push(x, y1, z, x + xx - 1, yb, z);
                }
            } else {
                if (s.v(x, yb, z) == 0 && x > 0 && s.v(x - 1, yb, z) == 0) {         // left
                    for (x2 = x; x2 < x + xx; ++x2) { if (x2 == W - 1) break; if (s.v(x2 + 1, yb, z) != 0) break; }
                    if (x2 == x + xx) x2 = x + xx - 1;
                    for (y1 = yb; y1 >= 0; --y1) { if (y1 == 0) break; if (!m3_row_free(s, x, x2 + 1, y1 - 1, z)) break; }
                    push(x, y1, z, x2, yb, z);
                }
                if (s.v(x + xx - 1, yb, z) == 0 && x + xx < W && s.v(x + xx, yb, z) == 0) {   // right
                    for (x1 = x + xx - 1; x1 >= x; --x1) { if (x1 == 0) break; if (s.v(x1 - 1, yb, z) != 0) break; }
                    if (x1 < x) x1 = x;
                    x1_bound = true;
                    for (y1 = yb; y1 >= 0; --y1) { if (y1 == 0) break; if (!m3_row_free(s, x1, x + xx, y1 - 1, z)) break; }
                    push(x1, y1, z, x + xx - 1, yb, z);
                }
            }
        }
        if (z + zz < H) {                                                           // on top
            const int t = z + zz;
            bool full = true;
            for (int q = x; q < x + xx && full; ++q) for (int r = y; r < y + yy; ++r) if (s.v(q, r, t) != 0) { full = false; break; }
            if (full) {
                if (!listed(x, y, t, x + xx - 1, y + yy - 1, t)) push(x, y, t, x + xx - 1, y + yy - 1, t);
            } else {
                // histogram of free run lengths along +x over the block's top face (:2915-2922); hist(i,j) recomputed on
                // demand from the voxels: number of consecutive free cells x+i, x+i+1, ... in column j
                auto hist = [&](int i, int j) { int n = 0; for (int q = i; q < xx && s.v(x + q, y + j, t) == 0; ++q) ++n; return n; };
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
// :3046-3050): a cell is empty iff it is empty now and not inside / under the candidate.
// The reference builds, per level, a histogram map hist[i][j] = length of the run of empty cells starting at (i,j) along
// x, and scans it for the largest rectangle of the form "hist value x maximal y-extent with hist >= that value".  Here a
// level is L bit rows (bit i of row[j] = cell (i,j) empty): hist(i,j) is a count-trailing-ones of row[j] >> i, so a level
// costs W*L voxel reads instead of the ~40 per cell the literal form needs (r02: this function was half of the kernel's
// 123 000 warp-instructions per environment and step).  Same values, same first-maximum semantics (only the max matters).
__device__ __forceinline__ long long m3_usable(const M3State &s, int cx, int cy, int cz, int bx, int by, int bz, int hlim) {
    const int W = s.W, L = s.L;
    const unsigned cmask = ((bx >= 32 ? 0xffffffffu : ((1u << bx) - 1u)) << cx);   // candidate columns along x
    long long score = 0;
    for (int hh = 0; hh < hlim; ++hh) {
        unsigned row[32];                                  // W*L <= 32 cells -> L <= 32 rows of <= 32 bits
        const bool under = hh < cz + bz;
        for (int j = 0; j < L; ++j) {
            unsigned m = 0u;
            for (int i = 0; i < W; ++i) m |= (s.v(i, j, hh) == 0 ? 1u : 0u) << i;
            if (under && j >= cy && j < cy + by) m &= ~cmask;
            row[j] = m;
        }
        auto hist = [&](int i, int j) { const unsigned t = ~(row[j] >> i); return t ? __ffs((int)t) - 1 : 32; };   // run of ones from bit i (bits >= W are 0)
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
        score += level_max;
    }
    return score;
}

// One block for one environment.  Returns the chosen placement (any == false: not placed).
__device__ __forceinline__ M3Best macs3d_place(const DevCfg &c, const M3State &s, int k, const int *positions, const int *blocks,
                                               int bx, int by, int bz, int valid_new, int empty, int nstable, int &anomaly) {
    const int W = s.W, L = s.L, H = s.H;
    const bool hard = (c.flags & TAPENV_RF_HARD) != 0;
    const bool mcs_start = (c.flags & TAPENV_RF_MCS_START) != 0, mcs_in = (c.flags & TAPENV_RF_MCS_IN) != 0;
    M3Best best; best.any = false; best.x = best.y = best.z = best.stable = best.add = 0; best.score = 0.0; best.mus = -1;
    M3Ems ems[kM3MaxEms];
    const int ne = m3_build_ems(s, k, positions, blocks, bx, by, bz, ems, anomaly);
    if (anomaly & 1) return best;
    int hmax0 = 0;
    for (int i = 0; i < s.cells; ++i) hmax0 = max(hmax0, s.h[i]);
    const int X = W - bx + 1, Y = L - by + 1;
    M3Visited vis;

    double best_score = 0.0;     // np.max(ratio_ems): never-settled entries are 0.0
    int count_best = 0, nsettled = 0, ncand = ne * 4, max_height = 0;
    M3Best first; first.any = false;

    for (int pass = 0; pass < 2; ++pass) {
        vis.clear();
        for (int ei = 0; ei < ne; ++ei) {
            const int X1 = ems[ei].x1, Y1 = ems[ei].y1, Z = ems[ei].z, X2 = ems[ei].x2, Y2 = ems[ei].y2;
            const int xr = X2 - bx + 2, yr = Y2 - by + 2;
            for (int corner = 0; corner < 4; ++corner) {
                bool ok;
                if (corner == 0) ok = X1 < X && Y1 < Y;
                else if (corner == 1) ok = xr > 0 && Y1 < Y;
                else if (corner == 2) ok = xr > 0 && yr > 0;
                else ok = X1 < X && yr > 0;
                if (!ok) continue;
                if (pass == 0) max_height = max(max_height, hmax0);                 // heightmap.copy() (:3091)
                // itertools.product order: corner 0 (x asc, y asc), 1 (y asc, x desc), 2 (x desc, y desc), 3 (y desc, x asc)
                const int na = corner == 0 ? X - X1 : (corner == 1 ? Y - Y1 : (corner == 2 ? xr : yr));
                const int nb = corner == 0 ? Y - Y1 : (corner == 1 ? xr : (corner == 2 ? yr : X - X1));
                bool settled = false, st = false;
                int px = 0, py = 0;
                const int vslot = vis.slot(Z, anomaly);
                for (int ia = 0; ia < na && !settled; ++ia)
                    for (int ib = 0; ib < nb && !settled; ++ib) {
                        int _x, _y;
                        if (corner == 0) { _x = X1 + ia; _y = Y1 + ib; }
                        else if (corner == 1) { _y = Y1 + ia; _x = xr - 1 - ib; }
                        else if (corner == 2) { _x = xr - 1 - ia; _y = yr - 1 - ib; }
                        else { _y = yr - 1 - ia; _x = X1 + ib; }
                        if (_x < 0 || _y < 0 || _x + bx > W || _y + by > L) { anomaly |= 1; continue; }
                        const unsigned bit = 1u << ((_x * L + _y) & 31);
                        if (vis.mask[vslot] & bit) continue;                                   // :2956
                        if (Z > 0) {                                                            // floating: skipped, NOT marked (:2957)
                            bool allz = true;
                            for (int q = _x; q < _x + bx && allz; ++q) for (int r = _y; r < _y + by; ++r) if (s.v(q, r, Z - 1) != 0) { allz = false; break; }
                            if (allz) continue;
                        }
                        vis.mask[vslot] |= bit;
                        bool freeall = true;
                        for (int q = _x; q < _x + bx && freeall; ++q) for (int r = _y; r < _y + by && freeall; ++r)
                            for (int t = Z; t < Z + bz && t < H; ++t) if (s.v(q, r, t) != 0) { freeall = false; break; }
                        if (!freeall) continue;
                        bool stable = true;
                        if (Z > 0) {
                            unsigned sup = 0u;
                            for (int i = 0; i < bx; ++i) for (int j = 0; j < by; ++j) if (s.v(_x + i, _y + j, Z - 1) > 0) sup |= 1u << ((i * by + j) & 31);
                            stable = stable3d_from_support(bx, by, sup);
                        }
                        if (!stable && hard) continue;
                        settled = true; st = stable; px = _x; py = _y;
                    }
                if (!settled) continue;
                // calc_C_P_S (:2971-2987)
                int height = 0;
                for (int q = 0; q < W; ++q) for (int r = 0; r < L; ++r) {
                    const int hv = (q >= px && q < px + bx && r >= py && r < py + by) ? Z + bz : s.h[q * L + r];
                    height = max(height, hv);
                }
                const int hm_max = height;
                if (Z + bx > height) height = Z + bz;                                          // sic :2976
                int cnt = 0;
                for (int q = px; q < px + bx; ++q) for (int r = py; r < py + by; ++r) for (int t = 0; t < Z && t < H; ++t) cnt += s.v(q, r, t) == 0 ? 1 : 0;
                const double ratio = mcs_start ? 0.0 : cps_score(c.flags, valid_new, height * W * L, empty + cnt, nstable + (st ? 1 : 0), k);
                if (pass == 0) {
                    ++nsettled;
                    max_height = max(max_height, hm_max);
                    if (!first.any || ratio > best_score) {
                        if (!first.any || ratio > best_score) count_best = 0;
                        best_score = ratio;
                        first.any = true; first.x = px; first.y = py; first.z = Z; first.stable = st ? 1 : 0; first.add = cnt; first.score = ratio; first.mus = 0;
                    }
                    if (ratio == best_score) ++count_best;
                } else if (ratio == best_score) {
                    const long long mus = m3_usable(s, px, py, Z, bx, by, bz, max_height);
                    if (!best.any || mus > best.mus) { best.any = true; best.x = px; best.y = py; best.z = Z; best.stable = st ? 1 : 0; best.add = cnt; best.score = ratio; best.mus = mus; }
                }
            }
        }
        if (pass == 0) {
            if (nsettled == 0) return best;                                                    // :3129-3132
            if (mcs_start) count_best = ncand;                                                 // every entry of ratio_ems is 0.0
            const bool tie = count_best > 1 && mcs_in;                                         // :3143
            if (!tie) return first;
            if (max_height > H) { anomaly |= 1; return best; }                                 // ctn[:, :, h] raises
        }
    }
    return best;
}

// commit (:3161-3171): update_container, update_level_free_space, heightmap
__device__ __forceinline__ void macs3d_commit(const M3State &s, int k, const M3Best &b, int bx, int by, int bz, int &anomaly) {
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

}  // namespace tapenv
