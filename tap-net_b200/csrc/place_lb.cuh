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
#pragma once
#include "tapenv_common.cuh"
#include "stable3d.cuh"

namespace tapenv {

struct LbState {
    short *vox;            // [cells][H]
    unsigned char *lists;  // [nlists][lcap]: byte 0 = length, then the x entries in append order
    int *h;                // [cells]
    int W, L, H, cells, lcap;
    __device__ __forceinline__ short &v(int cell, int z) const { return vox[cell * H + z]; }
    __device__ __forceinline__ unsigned char *list(int z, int y) const { return lists + (size_t)(z * L + y) * lcap; }
};

__device__ __forceinline__ bool lb_list_has(const unsigned char *l, int x) {
    for (int i = 1; i <= l[0]; ++i) if (l[i] == x) return true;
    return false;
}

// is_stable_2d (tools.py:839-868) on a voxel row: leading / trailing empty support cells
__device__ __forceinline__ bool lb_stable_2d(const LbState &s, int x, int z, int bx) {
    int l = 0, r = 0;
    while (l < bx && s.v(x + l, z - 1) == 0) ++l;
    while (r < bx && s.v(x + bx - 1 - r, z - 1) == 0) ++r;
    return 2 * l < bx && 2 * r < bx;             // x+l < x+bx/2 < x+bx-r
}

struct LbBest { double score; int x, y, z, stable, add; bool any; };

// One block for one environment.  DIM = 2 or 3.  Returns the placement (any == false: not placed).
template <int DIM>
__device__ __forceinline__ LbBest lb_place(const DevCfg &c, const LbState &s, int k, const int *positions, const int *blocks,
                                           int bx, int by, int bz, int valid_new, int empty, int nstable, int &anomaly) {
    const int W = s.W, L = s.L, H = s.H;
    const bool hard = (c.flags & TAPENV_RF_HARD) != 0;
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
                const double score = cps_score(c.flags, valid_new, (z0 + bz) * W * L, empty + cnt_empty, nstable + (st ? 1 : 0), k);
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
    constexpr int kMaxB = 2 * kMaxBlocks;
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
__device__ __forceinline__ void lb_commit(const LbState &s, int k, const LbBest &b, int bx, int by, int bz, int &anomaly) {
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

}  // namespace tapenv
