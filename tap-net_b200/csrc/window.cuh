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
    // rotblocks != 0: blocks[r*T + i][d] == blocks[i][perm_r[d]] for the rotations of itertools.permutations(range(dim))
    // (true for every reference dataset: generate_blocks writes the rotations that way) -- the kernel then reads the n
    // un-rotated rows instead of S scattered ones.  permcodes: 2 bits per (r, d) = perm_r[d], 6 bits per rotation.
    int rotblocks;
    unsigned long long permcodes;
};

struct WinShared {           // per warp
    unsigned long long rows[96];   // dynamic row (band*n + i') as a bit row over the S candidate columns
    unsigned char list[32];  // window in list order (survivors sorted, then admissions)
    unsigned char sorted[32];// ... sorted ascending (sub_graph_nodes after :1766)
    unsigned char perm[32];  // P
};

__device__ __forceinline__ unsigned long long below64(int v) { return v >= 64 ? ~0ull : ((1ull << v) - 1ull); }

// one CPython set insertion (set_add_entry / set_insert_clean probe order, distinct keys) into the 32-slot table.  The
// occupancy word `occ` and the key are warp-uniform, so the whole probe sequence runs on uniform values; only the owner
// lane of the chosen slot records the key.
__device__ __forceinline__ void pyset_insert32(unsigned &occ, int &slot, int lane, int key) {
    unsigned i = (unsigned)key & 31u, perturb = (unsigned)key;
#pragma unroll 1
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

// iteration order of set(list[0..len)) -> perm[0..len), keys < 64.
//   len >= 19: the table has grown to 128 slots (fill 19 >= 31*3/5), every key sits in slot `key` -> ascending order.
//   5 <= len <= 18 (32 slots): when no two keys agree modulo 32 every key sits in slot key & 31 whatever the insertion
//   order -> order by key & 31 (a rank computation); otherwise the insertions are replayed.
//   len < 5: the initial 8-slot table, replayed.
__device__ __forceinline__ void pyset_order_warp(WinShared &sh, int lane, int len, int mynode, unsigned long long winmask) {
    if (len >= 19) {
        if (mynode >= 0) sh.perm[__popcll(winmask & below64(mynode))] = (unsigned char)mynode;
        __syncwarp();
        return;
    }
    const unsigned homes = (unsigned)winmask | (unsigned)(winmask >> 32);            // slots key & 31 in use
    if (len >= 5 && __popc(homes) == len) {
        if (mynode >= 0) sh.perm[__popc(homes & ((1u << (mynode & 31)) - 1u))] = (unsigned char)mynode;
        __syncwarp();
        return;
    }
    {
        // the first 5 insertions go into the initial 8-slot table (mask 7: no linear probes, only the perturbed
        // recurrence), kept as 8 bytes (key + 1, 0 = unused) of one uniform 64-bit word
        unsigned long long t8 = 0ull;
        const int first = len < 5 ? len : 5;
        int q = 0;
#pragma unroll 1
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
#pragma unroll 1
            for (int s = 0; s < 8; ++s) {                                    // old entries re-inserted in slot order
                const int k = (int)((t8 >> (8 * s)) & 0xffull) - 1;
                if (k >= 0) pyset_insert32(occ, slot, lane, k);
            }
#pragma unroll 1
            for (; q < len; ++q) pyset_insert32(occ, slot, lane, sh.list[q]);
        }
        const unsigned full = __ballot_sync(TAPENV_FULL_MASK, slot >= 0);
        if (slot >= 0) sh.perm[__popc(full & ((1u << lane) - 1u))] = (unsigned char)slot;
    }
    __syncwarp();
}

struct WinEarly {                // what phase A hands to phase B (registers)
    unsigned long long gone, after, winmask;
    int len, flags, mynode;
    bool decompose;
    unsigned long long pa, pb;   // predecessor words of (graph slot, SORTED column) for the 3-graphs-per-ballot layout
    int bd0, bd1, bd2;           // un-rotated edge lengths of sorted node `lane` (rotation-structured blocks)
};

// Phase A: remove_block(sub_graph_nodes[rm]) (rm < 0: nothing) + the admission loop of sub_deps_graph.  Leaves the
// window in sh.list (list order: survivors sorted, then admissions) and sh.sorted, and ISSUES the global loads whose
// addresses depend only on the window's node set, so that their DRAM latency is covered by whatever the caller runs
// between the two phases (the placement) and by the set-order replay.
__device__ __forceinline__ WinEarly window_refill(const WinCfg &w, WinShared &sh, int lane, unsigned long long gone,
                                                  unsigned long long after, int len, int flags, int rm,
                                                  const unsigned long long *__restrict__ pe, unsigned long long pm0,
                                                  unsigned long long pm1, const int *__restrict__ blk) {
    const int T = w.T, n = w.n;
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
#pragma unroll 1
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
    WinEarly e;
    e.mynode = lane < len ? (int)sh.list[lane] : -1;
    const unsigned long long mybit = e.mynode >= 0 ? (1ull << e.mynode) : 0ull;
    e.winmask = (unsigned long long)__reduce_or_sync(TAPENV_FULL_MASK, (unsigned)mybit) |
                ((unsigned long long)__reduce_or_sync(TAPENV_FULL_MASK, (unsigned)(mybit >> 32)) << 32);
    if (e.mynode >= 0) sh.sorted[__popcll(e.winmask & below64(e.mynode))] = (unsigned char)e.mynode;   // .sort() (:1766)
    __syncwarp();
    e.gone = gone; e.after = after; e.len = len; e.flags = flags; e.decompose = decompose;
    e.pa = e.pb = 0ull;
    if (decompose && 3 * len <= 32) {
        const int q = lane >= 2 * len ? 2 : (lane >= len ? 1 : 0), jj = lane - q * len;
        if (lane < 3 * len) {
            const int v = sh.sorted[jj];
            e.pa = pe[q * T + v];
            if (w.dim == 3 && q < 2) e.pb = pe[(3 + q) * T + v];
        }
    }
    e.bd0 = e.bd1 = e.bd2 = 0;
    if (w.rotblocks && lane < len) {
        const int *p = blk + (int)sh.sorted[lane] * w.dim;
        e.bd0 = p[0]; e.bd1 = p[1];
        if (w.dim == 3) e.bd2 = p[2];
    }
    return e;
}

// Phase B: decompose + convert_to_input for environment b.  Emits the tensors and stores the new state.
// NWc / RWc > 0: window size and rotation count known at compile time (the emission loop unrolls into straight-line code
// with immediate offsets; r01 spent 350 warp-instructions per instance in it).  0 = runtime shape from WinCfg.
template <bool FAST, int NWc = 0, int RWc = 0>
__device__ __forceinline__ void window_emit(const WinCfg &w, WinShared &sh, const uint4 *lut, int b, int lane, WinEarly e,
                                            unsigned *ws, const unsigned long long *__restrict__ pe,
                                            const int *__restrict__ blk, float *__restrict__ static_out,
                                            float *__restrict__ dynamic_out, float *__restrict__ cur_mask,
                                            float *__restrict__ mask_out, int *__restrict__ nodes_out,
                                            int *__restrict__ remaining_out) {
    const int T = w.T, n = w.n, S = w.S, len = e.len;
    const unsigned long long after = e.after, winmask = e.winmask;
    unsigned long long gone = e.gone;
    int flags = e.flags;
    const int mynode = e.mynode;
    // ---- decompose: node enumeration P, the five [n,n] sub-matrices (generate.py:1682-1724, :1750-1764) ----
    // Row i' of sub-matrix g is the ballot over the columns jj of "P[i'] is a predecessor of P[jj]" (+ the diagonal
    // self-loop of the rotation graphs), kept by lane i'.
    unsigned myrow[5] = {0u, 0u, 0u, 0u, 0u};
    if (e.decompose) {
        if (w.setorder) pyset_order_warp(sh, lane, len, mynode, winmask);
        else { if (mynode >= 0) sh.perm[__popcll(winmask & below64(mynode))] = (unsigned char)mynode; __syncwarp(); }
        const int ng = w.dim == 3 ? 5 : 3;                  // 2D: forward/backward have no edges
        if (3 * len <= 32) {
            // lane = (graph slot q, column jj): one predecessor word per lane and pass, one ballot per sub-matrix row
            // yields that row of three graphs at once.  Pass 0: move/left/right, pass 1 (3D): forward/backward.  The
            // words were loaded in phase A by sorted column index; fetch the one of column node P[jj].
            const int q = lane >= 2 * len ? 2 : (lane >= len ? 1 : 0), jj = lane - q * len;
            const bool col = lane < 3 * len;
            const int v = col ? (int)sh.perm[jj] : 0;
            const int src = col ? q * len + __popcll(winmask & below64(v)) : 0;
            const unsigned alo = (unsigned)after, ahi = (unsigned)(after >> 32);
            const unsigned fld = (1u << len) - 1u;
            for (int pass = 0; pass < (w.dim == 3 ? 2 : 1); ++pass) {
                const int g = pass * 3 + q;
                const bool act = col && g < 5;
                const unsigned long long word = pass == 0 ? e.pa : e.pb;
                unsigned plo = __shfl_sync(TAPENV_FULL_MASK, (unsigned)word, src);
                unsigned phi = __shfl_sync(TAPENV_FULL_MASK, (unsigned)(word >> 32), src);
                if (!act) { plo = 0u; phi = 0u; }
                const bool loop = g >= 1 && ((plo & alo) | (phi & ahi)) != 0u;      // :1690-1705
                // lane (q, jj) now also OWNS row jj of sub-matrix g: entry (row jj, column c) = "P[jj] is a predecessor of
                // P[c]" = bit P[jj] of the word lane (q, c) holds.  (r01 built every row with one ballot per row and pass:
                // 426 of the kernel's 2 560 warp-instructions per instance; this gather form needs 2 shuffles per column.)
                unsigned keep = 0u;
                const unsigned ubit = (unsigned)v & 31u;
                const bool uhi = v >= 32;
                const int sbase = q * len;
                if (NWc > 0) {                                  // compile-time window: straight-line code, columns >= len masked below
#pragma unroll
                    for (int cc = 0; cc < (NWc > 0 ? NWc : 1); ++cc) {
                        const unsigned wlo = __shfl_sync(TAPENV_FULL_MASK, plo, (sbase + cc) & 31);
                        const unsigned whi = __shfl_sync(TAPENV_FULL_MASK, phi, (sbase + cc) & 31);
                        keep |= (((uhi ? whi : wlo) >> ubit) & 1u) << cc;
                    }
                    keep &= fld;
                } else {
#pragma unroll 1
                    for (int cc = 0; cc < len; ++cc) {
                        const unsigned wlo = __shfl_sync(TAPENV_FULL_MASK, plo, sbase + cc);
                        const unsigned whi = __shfl_sync(TAPENV_FULL_MASK, phi, sbase + cc);
                        keep |= (((uhi ? whi : wlo) >> ubit) & 1u) << cc;
                    }
                }
                if (loop) keep |= 1u << jj;                     // diagonal self-loop of the rotation graphs
                if (!act) keep = 0u;
                // rows of graph slot q live in lanes q*len .. q*len+len-1: hand them to the row lanes 0 .. len-1
                const unsigned k1 = __shfl_sync(TAPENV_FULL_MASK, keep, (lane + len) & 31);
                const unsigned k2 = __shfl_sync(TAPENV_FULL_MASK, keep, (lane + 2 * len) & 31);
                if (lane < len) {
                    if (pass == 0) { myrow[0] = keep & fld; myrow[1] = k1 & fld; myrow[2] = k2 & fld; }
                    else { myrow[3] = keep & fld; myrow[4] = k1 & fld; }
                }
            }
        } else {
            unsigned lo[5] = {0u, 0u, 0u, 0u, 0u}, hi[5] = {0u, 0u, 0u, 0u, 0u};
            unsigned loopbits = 0u;                         // bit g: a predecessor of this column still waits outside (:1690-1705)
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
#pragma unroll 1
            for (int i = 0; i < len; ++i) {
                const int u = sh.perm[i];                   // warp-uniform
                const unsigned diag = lane == i ? loopbits : 0u;
#pragma unroll
                for (int g = 0; g < 5; ++g) {
                    if (g < ng) {
                        const unsigned bit = u < 32 ? (lo[g] >> u) : (hi[g] >> (u - 32));
                        const unsigned rowbits = __ballot_sync(TAPENV_FULL_MASK, (bit | (diag >> g)) & 1u);
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
    if (len < n) flags |= 2;                                // the reference raises in np.concatenate (:1788)
    __syncwarp();
    const int sorted = lane < len ? (int)sh.sorted[lane] : 0xff;
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
    constexpr bool FX = NWc > 0;
    const int dimv = FX ? (RWc == 6 ? 3 : 2) : w.dim;   // compile-time with a fixed shape (R = dim!)
    const int Sv = FX ? NWc * RWc : S, nv = FX ? NWc : n;
    float *so = static_out + (size_t)b * (1 + dimv) * Sv;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int j = lane + 32 * half;
        if (FX && 32 * half >= NWc * RWc) break;
        const bool on = j < Sv;
        const int r = on ? (FX ? j / (FX ? NWc : 1) : (int)(((unsigned)j * w.inv_n) >> 16)) : 0, i = on ? j - r * nv : 0;
        if (w.rotblocks) {                                  // blocks[r*T + node][d] == blocks[node][perm_r[d]]
            const int a0 = __shfl_sync(TAPENV_FULL_MASK, e.bd0, i), a1 = __shfl_sync(TAPENV_FULL_MASK, e.bd1, i);
            const int a2 = dimv == 3 ? __shfl_sync(TAPENV_FULL_MASK, e.bd2, i) : 0;
            if (on) {
                so[j] = (float)i;
                const unsigned pc = (unsigned)(w.permcodes >> (6 * r));
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    if (d < dimv) {
                        const unsigned c = (pc >> (2 * d)) & 3u;
                        so[(1 + d) * Sv + j] = i < len ? (float)(c == 0u ? a0 : (c == 1u ? a1 : a2)) : 0.f;
                    }
                }
            }
        } else if (on) {
            so[j] = (float)i;
            const int node = i < len ? (int)sh.sorted[i] : -1;
            for (int d = 0; d < dimv; ++d) so[(1 + d) * Sv + j] = node >= 0 ? (float)blk[(node + r * T) * dimv + d] : 0.f;
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
    if (FAST && NWc > 0) {
        // compile-time shape: lane = (row-in-pass, column group), every pass at an immediate offset
        constexpr int SVc = NWc > 0 ? NWc * RWc / 4 : 1, RPc = 32 / SVc, rows3c = 3 * NWc, ITER = (rows3c + RPc - 1) / RPc;
        const int rsub = lane / SVc, cv = lane - rsub * SVc;
        const bool on = rsub < RPc;
        const bool hi_half = 4 * cv >= 32;
        const int sh4 = (4 * cv) & 31;
        uint4 *dst = reinterpret_cast<uint4 *>(dyo) + lane;
        const uint2 *rws = reinterpret_cast<const uint2 *>(sh.rows) + rsub;
#pragma unroll
        for (int it = 0; it < ITER; ++it) {
            if (on && it * RPc + rsub < rows3c) {
                const uint2 rw = rws[it * RPc];
                const unsigned nib = ((hi_half ? rw.y : rw.x) >> sh4) & 0xfu;
                stg_stream4(dst + it * RPc * SVc, lut[nib]);
            }
        }
    } else if (FAST) {
        // lane = (row-in-pass, column group): one 64-bit row word from shared memory, one nibble, one 128-bit store
        const int rsub = (int)(((unsigned)lane * w.inv_SV) >> 16), cv = lane - rsub * w.SV;
        const bool on = rsub < w.RP;
        const bool hi_half = 4 * cv >= 32;
        const int sh4 = (4 * cv) & 31;
        uint4 *dst = reinterpret_cast<uint4 *>(dyo) + lane;
        const int pstride = w.RP * w.SV, rows3 = 3 * n;
        const uint2 *rws = reinterpret_cast<const uint2 *>(sh.rows);
#pragma unroll 1
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
