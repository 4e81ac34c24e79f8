// tapenv_common.cuh -- shared device helpers for the TAP packing-environment kernels (sm_100a).
//
// Execution model of every kernel in this library: ONE WARP owns ONE environment instance; a CTA is
// kWarpsPerCta such warps.  All per-environment state lives in registers of that warp (one lane per
// heightmap column / cell); cross-column reductions are warp-level (shfl / ballot / redux.sync),
// never block-level, so no __syncthreads appears anywhere on the hot path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/tapenv.h"

#define TAPENV_FULL_MASK 0xffffffffu

namespace tapenv {

// Compiled limits (reported by tapenv_get_limits).
constexpr int kMaxWidth2D = 32;      // one lane per column
constexpr int kMaxCells3D = 32;      // one lane per heightmap cell
constexpr int kMaxCandidates = 64;   // S: one 64-bit accessibility word per band
constexpr int kMaxBlocks = 64;       // n (window) -- two history slots per lane for MACS
#ifndef TAPENV_WARPS_PER_CTA
#define TAPENV_WARPS_PER_CTA 4
#endif
constexpr int kWarpsPerCta = TAPENV_WARPS_PER_CTA;      // environments per CTA
constexpr int kMaxEms = 32 + 2 * kMaxBlocks;   // MACS: list A (<= W) + list B (<= 2 per previous block)

struct DevCfg {  // by-value kernel argument, derived from tapenv_config on the host
    int B, n, dim, R, W, L, H, S;
    int cap;          // capacity of the per-environment positions/blocks/stable arrays
    int strategy, hm_type, flags, ratio_mode;
    int static_rows, dyn_rows, update_time;
    int enc_len;      // encoded heightmap elements per env
    // host-precomputed geometry of the precedence pass (dynpass.cuh) and small-divisor reciprocals
    int SV, RP, PB, nbands;          // vectors per row, rows per pass, passes per band, bands present
    unsigned inv_SV, inv_n, inv_L;   // ceil(65536/d): q = (x*inv) >> 16 is exact for x < 64, d <= 64
    unsigned dyn_env, static_env;    // elements per environment of `dynamic` / `static`
    int lcap, nlists;                // LB strategy: bytes per x list (capacity+2) and lists per environment (H or H*L)
};

// Compile-time problem shape.  NT > 0: blocks_num = NT, rotate_types = RT, 'bot'-like input (3 bands, all
// zeroed by update_dynamic) -- every loop bound and divisor below folds to a constant.  NT == 0: runtime
// shape from DevCfg.
template <int NT, int RT, int DIM>
struct Shape {
    static constexpr bool fixed = NT > 0;
    static constexpr int NTc = NT;
    static constexpr int SVc = fixed ? (NT * RT) / 4 : 1;
    __device__ __forceinline__ static int n(const DevCfg &c) { return fixed ? NT : c.n; }
    __device__ __forceinline__ static int R(const DevCfg &c) { return fixed ? RT : c.R; }
    __device__ __forceinline__ static int S(const DevCfg &c) { return fixed ? NT * RT : c.S; }
    __device__ __forceinline__ static int SV(const DevCfg &c) { return fixed ? SVc : c.SV; }
    __device__ __forceinline__ static int RP(const DevCfg &c) { return fixed ? 32 / SVc : c.RP; }
    __device__ __forceinline__ static int PB(const DevCfg &c) { return fixed ? (NT + 32 / SVc - 1) / (32 / SVc) : c.PB; }
    __device__ __forceinline__ static int nbands(const DevCfg &c) { return fixed ? 3 : c.nbands; }
    __device__ __forceinline__ static int update_time(const DevCfg &c) { return fixed ? 3 : c.update_time; }
    __device__ __forceinline__ static int static_rows(const DevCfg &c) { return fixed ? 1 + DIM : c.static_rows; }
    __device__ __forceinline__ static unsigned dyn_env(const DevCfg &c) { return fixed ? 3u * NT * NT * RT : c.dyn_env; }
    __device__ __forceinline__ static unsigned static_env(const DevCfg &c) { return fixed ? (1u + DIM) * NT * RT : c.static_env; }
    __device__ __forceinline__ static int div_SV(const DevCfg &c, int x) { return fixed ? x / SVc : (int)(((unsigned)x * c.inv_SV) >> 16); }
    __device__ __forceinline__ static int mod_n(const DevCfg &c, int x) {
        return fixed ? x % NT : x - (int)(((unsigned)x * c.inv_n) >> 16) * c.n;
    }
};

// base + b * stride elements, as one 32x32->64 multiply-add
template <typename T>
__device__ __forceinline__ T *env_ptr(T *base, int b, unsigned stride) {
    return base + (unsigned long long)(unsigned)b * stride;
}

struct StatePtrs {
    int4 *scal;          // [B] (valid, empty, nstable, k)
    int *heightmap;      // [B][cells]
    int *positions;      // [B][cap][dim]
    int *blocks;         // [B][cap][dim]
    unsigned char *stable;  // [B][cap]
    int *flags;          // [B]
    short *voxels;       // LB only: [B][cells][H]
    unsigned char *lists;   // LB only: [B][nlists][lcap]
    float *pending;      // LB only: [B][4]
};

struct Scal { int valid, empty, nstable, k; };

struct PlaceOut {
    int placed;       // 0: the block could not be placed (state unchanged, tools.py:2084-2087, :2155-2158)
    int x, y, z;      // position (valid when placed; y = 0 in 2D)
    int stable;       // is_stable flag of the winner
    int top;          // z + block height (valid when placed)
};

// ---- streaming 128-bit global accesses (read-once / write-once data) ----
__device__ __forceinline__ uint4 ldg_stream4(const void *p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_stream4(void *p, const uint4 &v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ int warp_max(int v) { return __reduce_max_sync(TAPENV_FULL_MASK, v); }
__device__ __forceinline__ int warp_min(int v) { return __reduce_min_sync(TAPENV_FULL_MASK, v); }
__device__ __forceinline__ unsigned warp_or(unsigned v) { return __reduce_or_sync(TAPENV_FULL_MASK, v); }
__device__ __forceinline__ int warp_add(int v) { return __reduce_add_sync(TAPENV_FULL_MASK, v); }

// First-maximum argmax over positive finite fp64 scores held one per lane.
// `valid` lanes only; ties resolved by the smallest `key` (the reference takes the FIRST maximum in
// candidate order, np.argmax tools.py:2162).  Positive doubles order like their bit patterns.
// Returns the winning key (0xffffffff when no lane is valid); `hi`/`lo` return the winning score bits.
__device__ __forceinline__ unsigned warp_argmax_first(bool valid, double score, unsigned key,
                                                      unsigned *best_hi = nullptr, unsigned *best_lo = nullptr) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(score);
    const unsigned hi = (unsigned)(bits >> 32), lo = (unsigned)bits;
    const unsigned mh = __reduce_max_sync(TAPENV_FULL_MASK, valid ? hi : 0u);
    const bool c1 = valid && hi == mh;
    const unsigned ml = __reduce_max_sync(TAPENV_FULL_MASK, c1 ? lo : 0u);
    const bool c2 = c1 && lo == ml;
    if (best_hi) *best_hi = mh;
    if (best_lo) *best_lo = ml;
    return __reduce_min_sync(TAPENV_FULL_MASK, c2 ? key : 0xffffffffu);
}

// C+P+S score of one candidate, IEEE fp64 exactly as the reference evaluates it
// (tools.py:2124-2140, :2161): true divisions of integers, summed left to right.
__device__ __forceinline__ double cps_score(int flags, int valid_new, int bbox, int empty_new, int stable_cnt, int k) {
    const double vd = (double)valid_new;
    const double c = __ddiv_rn(vd, (double)bbox);
    const double p = (flags & TAPENV_RF_P) ? __ddiv_rn(vd, (double)(empty_new + valid_new)) : 0.0;
    const double s = (flags & TAPENV_RF_S) ? __ddiv_rn((double)stable_cnt, (double)(k + 1)) : 0.0;
    return __dadd_rn(__dadd_rn(c, p), s);
}

// Heightmap encodings returned by add_new_block (tools.py:3716-3743), 2D, lane = column.
__device__ __forceinline__ void encode_heightmap_2d(const DevCfg &c, int lane, int h, float *out) {
    if (c.hm_type == TAPENV_HM_DIFF) {             // h[i+1] - h[i], length W-1
        const int hn = __shfl_down_sync(TAPENV_FULL_MASK, h, 1);
        if (lane < c.W - 1) out[lane] = (float)(hn - h);
    } else if (c.hm_type == TAPENV_HM_ZERO) {      // h - min(h)
        const int m = warp_min(lane < c.W ? h : 0x7fffffff);
        if (lane < c.W) out[lane] = (float)(h - m);
    } else {
        if (lane < c.W) out[lane] = (float)h;
    }
}

// 3D, lane = cell x*L + y.  'diff' is [2,W,L]: backward differences along x and along y with a zero
// first row / column (tools.py:3721-3737).
__device__ __forceinline__ void encode_heightmap_3d(const DevCfg &c, int lane, int x, int y, int h, float *out) {
    const int cells = c.W * c.L;
    if (c.hm_type == TAPENV_HM_DIFF) {
        const int hu = __shfl_up_sync(TAPENV_FULL_MASK, h, c.L);   // (x-1, y)
        const int hl = __shfl_up_sync(TAPENV_FULL_MASK, h, 1);     // (x, y-1)
        if (lane < cells) {
            out[lane] = x > 0 ? (float)(h - hu) : 0.f;
            out[cells + lane] = y > 0 ? (float)(h - hl) : 0.f;
        }
    } else if (c.hm_type == TAPENV_HM_ZERO) {
        const int m = warp_min(lane < cells ? h : 0x7fffffff);
        if (lane < cells) out[lane] = (float)(h - m);
    } else {
        if (lane < cells) out[lane] = (float)h;
    }
}

}  // namespace tapenv
