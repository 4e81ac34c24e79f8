// place_lb.cuh -- the LB placement strategy (packing_strategy='LB'; tools.py:1602-1754 2D, :1756-1914 3D).
//
// Unlike LB_GREEDY / MACS this strategy cannot be reduced to the heightmap: its EMS filter compares voxel
// VALUES (block ids) between neighbouring rows (`container[x:, z] == container[x:, z-1]`, :1664), it keeps
// per-level x lists whose order is the order of appends (:1745-1749), and it has no "floating" test, so blocks
// may hover.  The state therefore carries the voxel grid (int16 ids: 0 empty, -1 empty-under-a-block, k+1) and
// the lists, and the algorithm is inherently a sequential walk over levels and list entries.  It runs ONE
// THREAD per environment (the batch is the parallel axis); no BASELINE configuration uses it.
//
// Driven through Container.add_new_block (:3683-3686), which never stores the returned bounding_box
// (:3706) -- every call starts from zeros, so a candidate's compactness is valid / ((_z+bz) * W(*L)).
//
// Two forms.  lb_place / lb_commit: the literal walk by ONE thread (lane 0 of the environment's warp) -- kept for containers
// taller than kLbMaxH.  lb_place_warp (r02): the warp first turns the voxel grid into LEVEL MASKS in shared memory (bit =
// cell x*L + y; emp[z]: voxel == 0, pos[z]: voxel > 0, dz[z] / dy[z]: voxel differs from the one below / from the row
// before), and every test of the walk becomes a bit operation: an EMS's scan for the first free (and stable) position is one
// ballot with lane = position (the scan order x-outer / y-inner IS ascending cell order), the value-equality filters are
// `(dz[z] & rect) == 0`, the empty cells under a footprint a popcount sum with lane = level.  Same results, same order of
// first maxima; the commit stays with lane 0.  The header also compiles for the host (one emulated lane,
// tests/test_lb_host_cpu.py).
#pragma once
#include <stdint.h>
#include "../../include/tapenv.h"
#include "stable3d.cuh"
#if defined(__CUDACC__)
#include "tapenv_common.cuh"
#endif

namespace tapenv {

struct LbState {
    short *vox;            // [cells][H]
    unsigned char *lists;  // [nlists][lcap]: byte 0 = length, then the x entries in append order
    int *h;                // [cells]
    int W, L, H, cells, lcap;
    TAPENV_HD short &v(int cell, int z) const { return vox[cell * H + z]; }
    TAPENV_HD unsigned char *list(int z, int y) const { return lists + (size_t)(z * L + y) * lcap; }
};

TAPENV_HD bool lb_list_has(const unsigned char *l, int x) {
    for (int i = 1; i <= l[0]; ++i) if (l[i] == x) return true;
    return false;
}

TAPENV_HD double lb_score(int flags, int valid_new, int bbox, int empty_new, int stable_cnt, int k) {
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

// is_stable_2d (tools.py:839-868) on a voxel row: leading / trailing empty support cells
TAPENV_HD bool lb_stable_2d(const LbState &s, int x, int z, int bx) {
    int l = 0, r = 0;
    while (l < bx && s.v(x + l, z - 1) == 0) ++l;
    while (r < bx && s.v(x + bx - 1 - r, z - 1) == 0) ++r;
    return 2 * l < bx && 2 * r < bx;             // x+l < x+bx/2 < x+bx-r
}

struct LbBest { double score; int x, y, z, stable, add; bool any; };

// One block for one environment.  DIM = 2 or 3.  Returns the placement (any == false: not placed).
template <int DIM>
TAPENV_HD LbBest lb_place(int flags, const LbState &s, int k, const int *positions, const int *blocks,
                          int bx, int by, int bz, int valid_new, int empty, int nstable, int &anomaly) {
    const int W = s.W, L = s.L, H = s.H;
    const bool hard = (flags & TAPENV_RF_HARD) != 0;
    const int X = W - bx + 1, Y = L - by + 1;
    LbBest best; best.any = false; best.score = -1.0; best.x = best.y = best.z = best.stable = best.add = 0;

    // scan one EMS (x0,y0,z0): first free (and, with `hard`, stable) position in x-outer / y-inner order
    auto try_ems = [&](int x0, int y0, int z0) {
        for (int _x = x0; _x < X; ++_x) {
            for (int _y = y0; _y < Y; ++_y) {
                bool free_all = true;
                unsigned sup = 0;
                int cnt_empty = 0;
                for (int i = 0; i < bx && free_all; ++i)
                    for (int j = 0; j < by && free_all; ++j) {
                        const int cell = (_x + i) * L + (_y + j);
                        for (int zz = z0; zz < z0 + bz && zz < H; ++zz) if (s.v(cell, zz) != 0) { free_all = false; break; }
                    }
                if (!free_all) continue;
                bool st;
                if (DIM == 2) st = z0 == 0 || lb_stable_2d(s, _x, z0, bx);
                else {
                    if (z0 == 0) st = true;
                    else {
                        for (int i = 0; i < bx; ++i) for (int j = 0; j < by; ++j)
                            if (s.v((_x + i) * L + (_y + j), z0 - 1) > 0) sup |= 1u << ((i * by + j) & 31);
                        st = stable3d_from_support(bx, by, sup);
                    }
                }
                if (!st && hard) continue;
                for (int i = 0; i < bx; ++i) for (int j = 0; j < by; ++j) {
                    const int cell = (_x + i) * L + (_y + j);
                    for (int zz = 0; zz < z0 && zz < H; ++zz) cnt_empty += s.v(cell, zz) == 0 ? 1 : 0;
                }
                const double score = lb_score(flags, valid_new, (z0 + bz) * W * L, empty + cnt_empty, nstable + (st ? 1 : 0), k);
                if (!best.any || score > best.score) {             // first maximum in EMS order
                    best.any = true; best.score = score; best.x = _x; best.y = _y; best.z = z0; best.stable = st ? 1 : 0; best.add = cnt_empty;
                }
                return;
            }
        }
    };

    // EMS list A: the per-level x lists (:1659-1666 / :1810-1823).  List B entries must not duplicate list A
    // entries, so list A is kept (bounded) for the membership test.
    constexpr int kMaxA = 160;
    short ea[kMaxA][3];
    int na = 0;
    for (int z = 0; z < H; ++z) {
        if (z + bz > H) break;
        if (z > 0) {
            bool zero = true;
            for (int cell = 0; cell < s.cells && zero; ++cell) zero = s.v(cell, z - 1) == 0;
            if (zero) break;
        }
        for (int y = 0; y < L; ++y) {
            if (DIM == 3) {
                if (y + by > L) break;
                if (y > 0) { const unsigned char *p = s.list(z, y - 1); if (p[0] == 1 && p[1] == 0) continue; }
            }
            const unsigned char *fs = s.list(z, y);
            for (int i = 1; i <= fs[0]; ++i) {
                const int x = fs[i];
                if (x + bx > W) break;
                if (DIM == 3 && y > 0 && lb_list_has(s.list(z, y - 1), x)) {
                    bool same = true;
                    for (int q = x; q < W && same; ++q) same = s.v(q * L + y, z) == s.v(q * L + y - 1, z);
                    if (same) continue;
                }
                if (z > 0 && lb_list_has(s.list(z - 1, y), x)) {
                    bool same = true;
                    for (int q = x; q < W && same; ++q)
                        for (int r = y; r < L && same; ++r) same = s.v(q * L + r, z) == s.v(q * L + r, z - 1);
                    if (same) continue;
                }
                if (na < kMaxA) { ea[na][0] = (short)x; ea[na][1] = (short)y; ea[na][2] = (short)z; }
                else anomaly |= 8;
                ++na;
                try_ems(x, y, z);
            }
        }
    }
    // EMS list B: corners on / behind the previous blocks (:1668-1677 / :1824-1833), de-duplicated against
    // everything listed so far
    constexpr int kMaxB = 128;                     // 2 per previous block (kMaxBlocks = 64)
    short eb[kMaxB][3];
    int nb = 0;
    auto listed = [&](int x, int y, int z) {
        for (int i = 0; i < na && i < kMaxA; ++i) if (ea[i][0] == x && ea[i][1] == y && ea[i][2] == z) return true;
        for (int i = 0; i < nb; ++i) if (eb[i][0] == x && eb[i][1] == y && eb[i][2] == z) return true;
        return false;
    };
    auto push_b = [&](int x, int y, int z) {
        if (nb < kMaxB) { eb[nb][0] = (short)x; eb[nb][1] = (short)y; eb[nb][2] = (short)z; ++nb; } else anomaly |= 8;
        try_ems(x, y, z);
    };
    for (int i = 0; i < k; ++i) {
        const int x = positions[i * DIM], y = DIM == 3 ? positions[i * DIM + 1] : 0, z = positions[i * DIM + DIM - 1];
        const int yy = DIM == 3 ? blocks[i * DIM + 1] : 0, zz = blocks[i * DIM + DIM - 1];
        if (DIM == 3 && y + yy < L) {
            if (s.v(x * L + y + yy, z) == 0 && !listed(x, y + yy, z)) push_b(x, y + yy, z);
        }
        if (z + zz < H) {
            if (s.v(x * L + y, z + zz) == 0 && !listed(x, y, z + zz)) push_b(x, y, z + zz);
        }
    }
    return best;
}

// commit of the winner (:1746-1764 / :1896-1911)
template <int DIM>
TAPENV_HD void lb_commit(const LbState &s, int k, const LbBest &b, int bx, int by, int bz, int &anomaly) {
    const int W = s.W, L = s.L, H = s.H;
    if (b.z + bz > H) { anomaly |= 1; return; }        // level_free_space[_z+zz] raises IndexError in the reference
    for (int i = 0; i < bx; ++i) for (int j = 0; j < by; ++j) {
        const int cell = (b.x + i) * L + (b.y + j);
        for (int zz = b.z; zz < b.z + bz; ++zz) s.v(cell, zz) = (short)(k + 1);
        for (int zz = 0; zz < b.z; ++zz) if (s.v(cell, zz) == 0) s.v(cell, zz) = -1;
        s.h[cell] = b.z + bz;
    }
    for (int zz = 0; zz < bz; ++zz) for (int yy = 0; yy < by; ++yy) {
        unsigned char *fs = s.list(b.z + zz, b.y + yy);
        for (int i = 1; i <= fs[0]; ++i) if (fs[i] == b.x) {           // list.remove: first occurrence
            for (int q = i; q < fs[0]; ++q) fs[q] = fs[q + 1];
            --fs[0];
            break;
        }
        if (b.x + bx < W && s.v((b.x + bx) * L + b.y + yy, b.z + zz) == 0) {
            if (fs[0] + 1 < s.lcap) { fs[++fs[0]] = (unsigned char)(b.x + bx); } else anomaly |= 8;
        }
    }
}

// ======================================================================================================================
// The warp form (see the header comment).  Results identical to lb_place / lb_commit above.
// ======================================================================================================================
constexpr int kLbMaxH = 256;         // tallest container the level masks are kept for (taller: the one-thread walk above)
constexpr int kLbMaxA = 160, kLbMaxB = 128;

struct LbScratch {                   // per warp, shared memory (5.3 kB)
    unsigned emp[kLbMaxH];           // bit (x*L + y): voxel (x, y, z) == 0
    unsigned pos[kLbMaxH];           // ... > 0 (a block)
    unsigned dz[kLbMaxH];            // ... differs from voxel (x, y, z-1)
    unsigned dy[kLbMaxH];            // ... differs from voxel (x, y-1, z)   (y > 0)
    unsigned keys[kLbMaxA + kLbMaxB];// the EMS listed so far: x | y << 8 | z << 16 (list A from 0, list B from kLbMaxA)
    unsigned verdict;                // lane 0's commit verdict, handed to the warp
};

struct LbWarp { LbScratch *sm; int lane, nl; };   // nl = lanes cooperating: 32 on the device, 1 in the host build

TAPENV_HD int lbw_sum(int v) {
#if defined(__CUDA_ARCH__)
    return __reduce_add_sync(0xffffffffu, v);
#else
    return v;
#endif
}
TAPENV_HD int lbw_max(int v) {
#if defined(__CUDA_ARCH__)
    return __reduce_max_sync(0xffffffffu, v);
#else
    return v;
#endif
}
TAPENV_HD bool lbw_any(bool v) {
#if defined(__CUDA_ARCH__)
    return __any_sync(0xffffffffu, v);
#else
    return v;
#endif
}
TAPENV_HD void lbw_sync() {
#if defined(__CUDA_ARCH__)
    __syncwarp();
#endif
}
// mask over the 32 positions p of pred(p): one ballot with lane = position (the host build loops)
template <class F>
TAPENV_HD unsigned lbw_ballot(int lane, F pred) {
#if defined(__CUDA_ARCH__)
    return __ballot_sync(0xffffffffu, pred(lane));
#else
    (void)lane;
    unsigned m = 0u;
    for (int p = 0; p < 32; ++p) if (pred(p)) m |= 1u << p;
    return m;
#endif
}
TAPENV_HD unsigned lb_bits(int n) { return n >= 32 ? 0xffffffffu : (n <= 0 ? 0u : ((1u << n) - 1u)); }

// level masks for z < zlim: lane = level, coalesced voxel reads
TAPENV_HD void lbw_build_masks(const LbState &s, const LbWarp &w, int zlim) {
    LbScratch &sm = *w.sm;
    for (int z = w.lane; z < zlim; z += w.nl) {
        unsigned e = 0u, p = 0u, dzm = 0u, dym = 0u;
        int y = 0, prev = 0;
        for (int cell = 0; cell < s.cells; ++cell) {
            const int val = s.v(cell, z);
            const unsigned bit = 1u << cell;
            if (val == 0) e |= bit;
            if (val > 0) p |= bit;
            if (z > 0 && val != s.v(cell, z - 1)) dzm |= bit;
            if (y > 0 && val != prev) dym |= bit;
            prev = val;
            if (++y == s.L) y = 0;
        }
        sm.emp[z] = e; sm.pos[z] = p; sm.dz[z] = dzm; sm.dy[z] = dym;
    }
    lbw_sync();
}

// One block for one environment, the whole warp.  hc: the height of the heightmap cell this lane owns (0 beyond the cells;
// the host build reads s.h instead).
template <int DIM>
TAPENV_HD LbBest lb_place_warp(int flags, const LbState &s, const LbWarp &w, int k, const int *positions, const int *blocks,
                               int bx, int by, int bz, int valid_new, int empty, int nstable, int hc, int &anomaly) {
    const int W = s.W, L = s.L, H = s.H, cells = s.cells;
    LbScratch &sm = *w.sm;
    const bool hard = (flags & TAPENV_RF_HARD) != 0;
    LbBest best; best.any = false; best.score = -1.0; best.x = best.y = best.z = best.stable = best.add = 0;
    // the masks are needed up to the highest level any scan can touch: the pile (or the phantom top of an unplaced block,
    // which sits at the origin with its own height) plus the new block
#if defined(__CUDA_ARCH__)
    int top = lbw_max(hc);
#else
    int top = 0;
    for (int i = 0; i < cells; ++i) top = top > s.h[i] ? top : s.h[i];
    (void)hc;
#endif
    for (int i = 0; i < k; ++i) { const int t = positions[i * DIM + DIM - 1] + blocks[i * DIM + DIM - 1]; top = top > t ? top : t; }
    const int zlim = top + bz < H ? top + bz : H;
    lbw_build_masks(s, w, zlim);
    const unsigned allcells = lb_bits(cells);
    unsigned fpat = 0u, colpat = 0u;
    for (int i = 0; i < bx; ++i) fpat |= lb_bits(by) << (i * L);         // footprint at the origin
    for (int q = 0; q < W; ++q) colpat |= 1u << (q * L);                 // cells (q, 0)
    const unsigned invL = (65536u + (unsigned)L - 1u) / (unsigned)L;     // p / L for p < 32 with one multiply

    // scan one EMS (x0,y0,z0): first free (and, with `hard`, stable) position in x-outer / y-inner order = ascending cell
    auto try_ems = [&](int x0, int y0, int z0) {
        unsigned f = allcells;
        for (int zz = z0; zz < z0 + bz && zz < H; ++zz) f &= sm.emp[zz];
        const unsigned under_e = z0 > 0 ? sm.emp[z0 - 1] : 0u, under_p = z0 > 0 ? sm.pos[z0 - 1] : 0u;
        auto stable_at = [&](int p) {                                     // z0 > 0
            const int px = (int)(((unsigned)p * invL) >> 16), py = p - px * L;
            if (DIM == 2) {                                               // is_stable_2d: leading / trailing empty support cells
                const unsigned sup = (~under_e >> px) & lb_bits(bx);
                const int l = sup ? tap_ctz(sup) : bx, r = sup ? bx - 1 - tap_fls(sup) : bx;
                return 2 * l < bx && 2 * r < bx;
            }
            const unsigned under = under_p & (fpat << p);
            const int cnt = tap_popc(under);
            if (2 * cnt > bx * by) return true;
            if (cnt <= 1) return false;
            unsigned sup = 0u;
            for (unsigned m = under; m; m &= m - 1u) {
                const int cbit = tap_ctz(m);
                const int cx = (int)(((unsigned)cbit * invL) >> 16), cy = cbit - cx * L;
                sup |= 1u << (((cx - px) * by + (cy - py)) & 31);
            }
            return stable3d_from_support(bx, by, sup);
        };
        const unsigned okm = lbw_ballot(w.lane, [&](int p) {
            const int px = (int)(((unsigned)p * invL) >> 16), py = p - px * L;
            if (!(p < cells && px + bx <= W && py + by <= L && px >= x0 && py >= y0)) return false;
            const unsigned foot = fpat << p;
            if ((f & foot) != foot) return false;
            return !hard || z0 == 0 || stable_at(p);
        });
        if (!okm) return;
        const int p = tap_ctz(okm);
        const int _x = (int)(((unsigned)p * invL) >> 16), _y = p - _x * L;
        const bool st = z0 == 0 || hard || stable_at(p);                  // hard: only stable positions are in okm
        const unsigned foot = fpat << p;
        int part = 0;
        for (int zz = w.lane; zz < z0 && zz < H; zz += w.nl) part += tap_popc(sm.emp[zz] & foot);
        const int cnt_empty = lbw_sum(part);
        const double score = lb_score(flags, valid_new, (z0 + bz) * W * L, empty + cnt_empty, nstable + (st ? 1 : 0), k);
        if (!best.any || score > best.score) {                            // first maximum in EMS order
            best.any = true; best.score = score; best.x = _x; best.y = _y; best.z = z0; best.stable = st ? 1 : 0; best.add = cnt_empty;
        }
    };

    // EMS list A: the per-level x lists (:1659-1666 / :1810-1823)
    int na = 0, nb = 0;
    for (int z = 0; z < H; ++z) {
        if (z + bz > H) break;
        if (z > 0 && sm.emp[z - 1] == allcells) break;                    // the level below is empty: above the pile
        for (int y = 0; y < L; ++y) {
            if (DIM == 3) {
                if (y + by > L) break;
                if (y > 0) { const unsigned char *p = s.list(z, y - 1); if (p[0] == 1 && p[1] == 0) continue; }
            }
            const unsigned char *fs = s.list(z, y);
            for (int i = 1; i <= fs[0]; ++i) {
                const int x = fs[i];
                if (x + bx > W) break;
                if (DIM == 3 && y > 0 && lb_list_has(s.list(z, y - 1), x)) {     // container[x:, y, z] == container[x:, y-1, z]
                    const unsigned col = ((colpat << (x * L)) & allcells) << y;
                    if ((sm.dy[z] & col) == 0u) continue;
                }
                if (z > 0 && lb_list_has(s.list(z - 1, y), x)) {                 // container[x:, y:, z] == container[x:, y:, z-1]
                    const unsigned rect = ((colpat * (lb_bits(L) & ~lb_bits(y))) << (x * L)) & allcells;
                    if ((sm.dz[z] & rect) == 0u) continue;
                }
                if (na < kLbMaxA) { if (w.lane == 0) sm.keys[na] = (unsigned)x | ((unsigned)y << 8) | ((unsigned)z << 16); }
                else anomaly |= 8;
                ++na;
                try_ems(x, y, z);
            }
        }
    }
    // EMS list B: corners on / behind the previous blocks (:1668-1677 / :1824-1833), de-duplicated against everything listed
    auto listed = [&](int x, int y, int z) {
        lbw_sync();
        const unsigned key = (unsigned)x | ((unsigned)y << 8) | ((unsigned)z << 16);
        bool hit = false;
        const int nA = na < kLbMaxA ? na : kLbMaxA;
        for (int i = w.lane; i < nA; i += w.nl) hit |= sm.keys[i] == key;
        for (int i = w.lane; i < nb; i += w.nl) hit |= sm.keys[kLbMaxA + i] == key;
        return lbw_any(hit);
    };
    auto push_b = [&](int x, int y, int z) {
        if (nb < kLbMaxB) { if (w.lane == 0) sm.keys[kLbMaxA + nb] = (unsigned)x | ((unsigned)y << 8) | ((unsigned)z << 16); ++nb; }
        else anomaly |= 8;
        try_ems(x, y, z);
    };
    for (int i = 0; i < k; ++i) {
        const int x = positions[i * DIM], y = DIM == 3 ? positions[i * DIM + 1] : 0, z = positions[i * DIM + DIM - 1];
        const int yy = DIM == 3 ? blocks[i * DIM + 1] : 0, zz = blocks[i * DIM + DIM - 1];
        if (DIM == 3 && y + yy < L) {
            if (((sm.emp[z] >> (x * L + y + yy)) & 1u) && !listed(x, y + yy, z)) push_b(x, y + yy, z);
        }
        if (z + zz < H) {
            if (((sm.emp[z + zz] >> (x * L + y)) & 1u) && !listed(x, y, z + zz)) push_b(x, y, z + zz);
        }
    }
    return best;
}

// Container.add_new_block for one environment with the LB strategy, the whole warp.  scal: (valid, empty, #stable, k).
// Returns the anomaly bits (identical in every lane).
template <int DIM>
TAPENV_HD int lb_env_add_block_warp(int flags, int cap, const LbState &s, const LbWarp &w, int *scal, int *positions, int *blks,
                                    unsigned char *stable_out, int bx, int by, int bz) {
    const int s0 = scal[0], s1 = scal[1], s2 = scal[2], k = scal[3];
    int anomaly = 0;
    if (k >= cap) return 2;
#if defined(__CUDA_ARCH__)
    const int hc = w.lane < s.cells ? s.h[w.lane] : 0;
#else
    const int hc = 0;
#endif
    lbw_sync();                                          // every lane has read the state before lane 0 edits it
    int o0 = s0, o1 = s1, o2 = s2;
    unsigned char stable = 0;
    if (bx >= 1 && by >= 1 && bz >= 1 && bx <= s.W && by <= s.L) {
        const int vol = bx * by * bz;
        const LbBest best = lb_place_warp<DIM>(flags, s, w, k, positions, blks, bx, by, bz, s0 + vol, s1, s2, hc, anomaly);
        if (best.any) {
            lbw_sync();
            if (w.lane == 0) {
                int a2 = 0;
                lb_commit<DIM>(s, k, best, bx, by, bz, a2);
                if (!(a2 & 1)) { positions[k * DIM] = best.x; if (DIM == 3) positions[k * DIM + 1] = best.y; positions[k * DIM + DIM - 1] = best.z; }
                w.sm->verdict = (unsigned)a2;
            }
            lbw_sync();
            const int a2 = (int)w.sm->verdict;
            anomaly |= a2;
            if (!(a2 & 1)) { stable = (unsigned char)best.stable; o0 = s0 + vol; o1 = s1 + best.add; o2 = s2 + best.stable; }
        }
    }
    if (w.lane == 0) {
        blks[k * DIM] = bx; if (DIM == 3) blks[k * DIM + 1] = by; blks[k * DIM + DIM - 1] = bz;
        stable_out[k] = stable;
        scal[0] = o0; scal[1] = o1; scal[2] = o2; scal[3] = k + 1;
    }
    return anomaly;
}

}  // namespace tapenv
