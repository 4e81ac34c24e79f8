// window.cuh -- the rolling window of one environment (generate.InitialContainer, generate.py:1589-1825), one warp
// per environment.
//
// The reference keeps five networkx DiGraphs per instance and, before EVERY decode step of rolling inference
// (rolling.py:593), rebuilds the network's window: remove the block chosen last (remove_block :1810), admit
// in-degree-0 nodes of the remaining movement graph in ascending id order until `child_graph_size` nodes are in
// the window (sub_deps_graph :1726-1748), extract the five induced sub-graphs (decompose :1682-1724) and lay them
// out as static [1+dim,S] / dynamic [3n,S] (convert_to_input :1770-1808).
//
// Here a graph is T (<= 64) predecessor words: bit u of pred[g][v] <=> edge u -> v.  Node sets (`gone` = removed
// from self.gm, `after` = after_nodes_list, the window itself) are 64-bit masks, two nodes per lane:
//   in-degree-0 test     (pred_move[v] & alive) == 0              one AND per node, ballot -> 64-bit wave
//   admission order      rank of v inside the wave = popc(wave & below(v)); the first `need` are taken
//   sub-matrix column jj  bit i' = (pred[g][P[jj]] >> P[i']) & 1   P = node enumeration of the sub-graph
//   self-loop rule        (pred[g][P[jj]] & after) != 0           generate.py:1690-1705
// P is NOT the sorted window when 2*window < total: networkx's subgraph view then enumerates the Python *set* of the
// window nodes (coreviews.FilterAtlas.__iter__), whose order is CPython's open-addressing layout (setobject.c).  The
// reference indexes `dynamic` by P while `static` uses the sorted list; this is reproduced bit-exactly -- the
// probe sequence of every insertion runs on warp-uniform values (8-slot table = one 64-bit word, 32-slot table = an
// occupancy word + one slot per lane).
#pragma once
#include "tapenv_common.cuh"

namespace tapenv {

constexpr int kWinStateWords = 16;   // 64 bytes per environment: gone u64 | after u64 | window u8[32] | len | flags | pad
constexpr int kWinMaxTotal = 64;
constexpr int kWinMaxWindow = 32;

struct WinCfg {
    int B, T, n, dim, R, S;
    int setorder;            // 1: P = CPython set order (reference behaviour when 2n < T), 0: ascending
    int SV, RP, PB;          // 128-bit emission geometry (as dynpass.cuh): vectors per row, rows per pass, passes per band
    unsigned inv_n, inv_SV;  // ceil(65536/d)
    unsigned lastcodes;      // 2 bits per rotation: 0 -> (left,right), 1 -> (forward,backward), 2 -> zeros (generate.py:1793-1806)
    unsigned blocks_env;     // R*T*dim
    // replication multipliers: sum of 2^(r*n) over all rotations / those with code 0 / code 1 -- an n-bit sub-matrix row
    // times one of these is that row repeated into the column blocks of the rotations it applies to
    unsigned long long mul_all, mul_c0, mul_c1;
};

struct WinShared {           // per warp
    unsigned long long rows[96];   // dynamic row (band*n + i') as a bit row over the S candidate columns
    unsigned char list[32];  // window in list order (survivors sorted, then admissions), later sorted
    unsigned char perm[32];  // P
    unsigned char tab[2][128];   // sequential set emulation for windows above 18 nodes
};

__device__ __forceinline__ unsigned long long below64(int v) { return v >= 64 ? ~0ull : ((1ull << v) - 1ull); }

// one CPython set insertion (set_add_entry / set_insert_clean probe order, distinct keys) into the 32-slot table.  The
// occupancy word `occ` and the key are warp-uniform, so the whole probe sequence runs on uniform values; only the owner
// lane of the chosen slot records the key.
__device__ __forceinline__ void pyset_insert32(unsigned &occ, int &slot, int lane, int key) {
    unsigned i = (unsigned)key & 31u, perturb = (unsigned)key;
    for (int guard = 0; guard < 64; ++guard) {
        const unsigned probes = (i + 9u <= 31u) ? 9u : 0u;                   // LINEAR_PROBES
        const unsigned win = ((~occ) >> i) & ((2u << probes) - 1u);
        if (win) {
            const unsigned s = i + (unsigned)__ffs((int)win) - 1u;
            occ |= 1u << s;
            if (lane == (int)s) slot = key;
            return;
        }
        perturb >>= 5;                                                       // PERTURB_SHIFT
        i = (i * 5u + 1u + perturb) & 31u;
    }
}

// iteration order of set(list[0..len)) -> perm[0..len).  len <= 18 keeps CPython's table at <= 32 slots.
__device__ __forceinline__ void pyset_order_warp(WinShared &sh, int lane, int len) {
    if (len <= 18) {
        // the first 5 insertions go into the initial 8-slot table (mask 7: no linear probes, only the perturbed
        // recurrence), kept as 8 bytes (key + 1, 0 = unused) of one uniform 64-bit word
        unsigned long long t8 = 0ull;
        const int first = len < 5 ? len : 5;
        int q = 0;
        for (; q < first; ++q) {
            const unsigned key = sh.list[q];
            unsigned i = key & 7u, perturb = key;
            for (int guard = 0; guard < 64 && ((t8 >> (8u * i)) & 0xffull); ++guard) { perturb >>= 5; i = (i * 5u + 1u + perturb) & 7u; }
            t8 |= (unsigned long long)(key + 1u) << (8u * i);
        }
        unsigned occ = 0u;
        int slot = -1;
        if (len < 5) {                                                       // fill * 5 < mask * 3: the table never grows
            const unsigned byte = lane < 8 ? (unsigned)((t8 >> (8 * lane)) & 0xffull) : 0u;
            slot = (int)byte - 1;
        } else {                                                             // set_table_resize(used * 4): 8 -> 32 slots at fill 5,
            for (int s = 0; s < 8; ++s) {                                    // old entries re-inserted in slot order
                const int k = (int)((t8 >> (8 * s)) & 0xffull) - 1;
                if (k >= 0) pyset_insert32(occ, slot, lane, k);
            }
            for (; q < len; ++q) pyset_insert32(occ, slot, lane, sh.list[q]);
        }
        const unsigned full = __ballot_sync(TAPENV_FULL_MASK, slot >= 0);
        if (slot >= 0) sh.perm[__popc(full & ((1u << lane) - 1u))] = (unsigned char)slot;
    } else {                                                                 // 19..32 nodes: the table reaches 128 slots
        if (lane == 0) {
            unsigned char *table = sh.tab[0], *other = sh.tab[1];
            unsigned mask = 7u;
            int fill = 0;
            for (unsigned s = 0; s < 128u; ++s) table[s] = 0xff;
            for (int q = 0; q < len; ++q) {
                const unsigned key = sh.list[q];
                unsigned i = key & mask, perturb = key;
                for (bool placed = false; !placed;) {
                    const unsigned probes = (i + 9u <= mask) ? 9u : 0u;
                    for (unsigned j = 0; j <= probes; ++j) if (table[i + j] == 0xff) { table[i + j] = (unsigned char)key; placed = true; break; }
                    perturb >>= 5; i = (i * 5u + 1u + perturb) & mask;
                }
                ++fill;
                if ((unsigned)fill * 5u >= mask * 3u) {
                    const unsigned nmask = mask == 7u ? 31u : 127u;
                    for (unsigned s = 0; s <= nmask; ++s) other[s] = 0xff;
                    for (unsigned s = 0; s <= mask; ++s) {
                        if (table[s] == 0xff) continue;
                        const unsigned k = table[s];
                        unsigned i2 = k & nmask, pt = k;
                        for (bool placed = false; !placed;) {
                            const unsigned probes = (i2 + 9u <= nmask) ? 9u : 0u;
                            for (unsigned j = 0; j <= probes; ++j) if (other[i2 + j] == 0xff) { other[i2 + j] = (unsigned char)k; placed = true; break; }
                            pt >>= 5; i2 = (i2 * 5u + 1u + pt) & nmask;
                        }
                    }
                    unsigned char *t = table; table = other; other = t;
                    mask = nmask;
                }
            }
            int cnt = 0;
            for (unsigned s = 0; s <= mask; ++s) if (table[s] != 0xff) sh.perm[cnt++] = table[s];
        }
    }
    __syncwarp();
}

// remove_block(sub_graph_nodes[rm]) (rm < 0: nothing) + sub_deps_graph + convert_to_input for environment b.
// On entry sh.list holds the stored window (list order == sorted).  Emits the tensors and stores the new state.
template <bool FAST>
__device__ __forceinline__ void window_advance(const WinCfg &w, WinShared &sh, const uint4 *lut, int b, int lane, unsigned long long gone,
                                               unsigned long long after, int len, int flags, int rm, unsigned *ws,
                                               const unsigned long long *__restrict__ pe, unsigned long long pm0,
                                               unsigned long long pm1, const int *__restrict__ blk,
                                               float *__restrict__ static_out, float *__restrict__ dynamic_out,
                                               float *__restrict__ cur_mask, float *__restrict__ mask_out,
                                               int *__restrict__ nodes_out, int *__restrict__ remaining_out) {
    const int T = w.T, n = w.n, S = w.S;
    // ---- remove_block: list.remove(value) (generate.py:1818) ----
    if (rm >= 0 && rm < len) {
        const int mine = sh.list[lane];
        const int nxt = __shfl_down_sync(TAPENV_FULL_MASK, mine, 1);
        __syncwarp();
        if (lane >= rm && lane < 31) sh.list[lane] = (unsigned char)nxt;
        --len;
        __syncwarp();
    }
    // ---- sub_deps_graph: admit in-degree-0 nodes of gm_copy, ascending, wave by wave (generate.py:1726-1748) ----
    const unsigned long long fullT = below64(T);
    unsigned long long alive = fullT & ~gone;
    bool decompose = false;
    const int v0 = lane, v1 = lane + 32;
    for (int guard = 0; guard <= kWinMaxTotal; ++guard) {
        const int cnt = __popcll(alive);
        if (cnt == 0) break;
        const bool z0 = ((alive >> v0) & 1ull) && (cnt == 1 || (pm0 & alive) == 0ull);
        const bool z1 = ((alive >> v1) & 1ull) && (cnt == 1 || (pm1 & alive) == 0ull);   // v1 >= 32; alive has no bits >= T
        const unsigned long long wave = (unsigned long long)__ballot_sync(TAPENV_FULL_MASK, z0) |
                                        ((unsigned long long)__ballot_sync(TAPENV_FULL_MASK, z1) << 32);
        if (wave == 0ull) { flags |= 1; break; }            // no in-degree-0 node: the reference spins forever
        if (len == n) { decompose = true; break; }          // :1737-1740
        const int need = n - len;
        const int r0 = __popcll(wave & below64(v0)), r1 = __popcll(wave & below64(v1));
        const bool t0 = z0 && r0 < need, t1 = z1 && r1 < need;
        if (t0) sh.list[len + r0] = (unsigned char)v0;      // :1742
        if (t1) sh.list[len + r1] = (unsigned char)v1;
        const unsigned long long take = (unsigned long long)__ballot_sync(TAPENV_FULL_MASK, t0) |
                                        ((unsigned long long)__ballot_sync(TAPENV_FULL_MASK, t1) << 32);
        len += __popcll(take);
        alive &= ~take;                                     // :1743
        after &= ~take;                                     // :1744
        if (len == n) { decompose = true; break; }          // :1745-1748
    }
    __syncwarp();
    const int mynode = lane < len ? (int)sh.list[lane] : -1;
    const unsigned long long mybit = mynode >= 0 ? (1ull << mynode) : 0ull;
    const unsigned long long winmask = (unsigned long long)__reduce_or_sync(TAPENV_FULL_MASK, (unsigned)mybit) |
                                       ((unsigned long long)__reduce_or_sync(TAPENV_FULL_MASK, (unsigned)(mybit >> 32)) << 32);
    // ---- decompose: node enumeration P, the five [n,n] sub-matrices (generate.py:1682-1724, :1750-1764) ----
    // Lane jj holds the predecessor words of column node P[jj]; row i' of sub-matrix g is the ballot over the columns of
    // "P[i'] is a predecessor of P[jj]" (+ the diagonal self-loop of the rotation graphs), kept by lane i'.
    unsigned myrow[5] = {0u, 0u, 0u, 0u, 0u};
    if (decompose) {
        if (w.setorder) pyset_order_warp(sh, lane, len);
        else { if (mynode >= 0) sh.perm[__popcll(winmask & below64(mynode))] = (unsigned char)mynode; __syncwarp(); }
        const int ng = w.dim == 3 ? 5 : 3;                  // 2D: forward/backward have no edges
        unsigned lo[5] = {0u, 0u, 0u, 0u, 0u}, hi[5] = {0u, 0u, 0u, 0u, 0u};
        unsigned loopbits = 0u;                             // bit g: a predecessor of this column still waits outside (:1690-1705)
        if (lane < len) {
            const int v = sh.perm[lane];
            const unsigned alo = (unsigned)after, ahi = (unsigned)(after >> 32);
#pragma unroll
            for (int g = 0; g < 5; ++g) {
                if (g < ng) {
                    const unsigned long long pg = pe[g * T + v];
                    lo[g] = (unsigned)pg; hi[g] = (unsigned)(pg >> 32);
                    if (g >= 1 && ((lo[g] & alo) | (hi[g] & ahi))) loopbits |= 1u << g;
                }
            }
        }
        for (int i = 0; i < len; ++i) {
            const int u = sh.perm[i];                       // warp-uniform
            const unsigned diag = lane == i ? loopbits : 0u;
            if (u < 32) {
#pragma unroll
                for (int g = 0; g < 5; ++g) {
                    if (g < ng) {
                        const unsigned rowbits = __ballot_sync(TAPENV_FULL_MASK, ((lo[g] >> u) | (diag >> g)) & 1u);
                        if (lane == i) myrow[g] = rowbits;
                    }
                }
            } else {
#pragma unroll
                for (int g = 0; g < 5; ++g) {
                    if (g < ng) {
                        const unsigned rowbits = __ballot_sync(TAPENV_FULL_MASK, ((hi[g] >> (u - 32)) | (diag >> g)) & 1u);
                        if (lane == i) myrow[g] = rowbits;
                    }
                }
            }
        }
        gone |= winmask;                                    // :1713-1723
    }
    // dynamic rows: band 0 = move in every rotation block; bands 1/2 = (left,right) where the rotation's last axis is x,
    // (forward,backward) where it is y, zeros where it is the vertical axis (generate.py:1790-1806)
    const unsigned long long row0 = (unsigned long long)myrow[0] * w.mul_all;
    const unsigned long long row1 = (unsigned long long)myrow[1] * w.mul_c0 + (unsigned long long)myrow[3] * w.mul_c1;
    const unsigned long long row2 = (unsigned long long)myrow[2] * w.mul_c0 + (unsigned long long)myrow[4] * w.mul_c1;
    if (lane < n) { sh.rows[lane] = row0; sh.rows[n + lane] = row1; sh.rows[2 * n + lane] = row2; }
    const unsigned long long any0 = (unsigned long long)warp_or((unsigned)row0) | ((unsigned long long)warp_or((unsigned)(row0 >> 32)) << 32);
    const unsigned long long any1 = (unsigned long long)warp_or((unsigned)row1) | ((unsigned long long)warp_or((unsigned)(row1 >> 32)) << 32);
    const unsigned long long any2 = (unsigned long long)warp_or((unsigned)row2) | ((unsigned long long)warp_or((unsigned)(row2 >> 32)) << 32);
    const unsigned long long blocked = any0 | (any1 & any2);      // rolling.py:325-335
    // ---- self.sub_graph_nodes.sort() (:1766) ----
    __syncwarp();
    if (mynode >= 0) sh.list[__popcll(winmask & below64(mynode))] = (unsigned char)mynode;
    if (len < n) flags |= 2;                                // the reference raises in np.concatenate (:1788)
    __syncwarp();
    const int sorted = lane < len ? (int)sh.list[lane] : 0xff;
    // ---- state ----
    if (lane == 0) {
        *reinterpret_cast<unsigned long long *>(ws) = gone;
        *reinterpret_cast<unsigned long long *>(ws + 2) = after;
        ws[12] = (unsigned)len; ws[13] = (unsigned)flags;
    }
    reinterpret_cast<unsigned char *>(ws + 4)[lane] = (unsigned char)sorted;
    if (nodes_out && lane < n) nodes_out[(size_t)b * n + lane] = lane < len ? sorted : -1;
    if (remaining_out && lane == 0) remaining_out[b] = __popcll(after);
    // ---- static [1+dim,S]: row 0 = window-local index, rows 1.. = blocks[node + r*T] (:1779-1788) ----
    float *so = static_out + (size_t)b * (1 + w.dim) * S;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int j = lane + 32 * half;
        if (j < S) {
            const int r = (int)(((unsigned)j * w.inv_n) >> 16), i = j - r * n;
            so[j] = (float)i;
            const int node = i < len ? (int)sh.list[i] : -1;
            for (int d = 0; d < w.dim; ++d) so[(1 + d) * S + j] = node >= 0 ? (float)blk[(node + r * T) * w.dim + d] : 0.f;
        }
    }
    // ---- dynamic [3n,S] and the initial masks (rolling.py:325-335) ----
    float *dyo = dynamic_out + (size_t)b * 3 * n * S;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int j = lane + 32 * half;
        if (j < S) {
            if (cur_mask) cur_mask[(size_t)b * S + j] = ((blocked >> j) & 1ull) ? 0.f : 1.f;
            if (mask_out) mask_out[(size_t)b * S + j] = 1.f;
        }
    }
    if (FAST) {
        // lane = (row-in-pass, column group): one 64-bit row word from shared memory, one nibble, one 128-bit store
        const int rsub = (int)(((unsigned)lane * w.inv_SV) >> 16), cv = lane - rsub * w.SV;
        const bool on = rsub < w.RP;
        const bool hi_half = 4 * cv >= 32;
        const int sh4 = (4 * cv) & 31;
        uint4 *dst = reinterpret_cast<uint4 *>(dyo) + lane;
        const int pstride = w.RP * w.SV, rows3 = 3 * n;
        const uint2 *rws = reinterpret_cast<const uint2 *>(sh.rows);
        for (int fr = rsub; fr < rows3; fr += w.RP) {
            if (on) {
                const uint2 rw = rws[fr];
                const unsigned nib = ((hi_half ? rw.y : rw.x) >> sh4) & 0xfu;
                stg_stream4(dst, lut[nib]);
            }
            dst += pstride;
        }
    } else {
        for (int q = lane; q < 3 * n * S; q += 32) {
            const int fr = q / S, col = q - fr * S;
            dyo[q] = ((sh.rows[fr] >> col) & 1ull) ? 1.f : 0.f;
        }
    }
}

}  // namespace tapenv
