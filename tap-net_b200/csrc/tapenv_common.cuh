// tapenv_common.cuh -- shared device helpers for the TAP packing-environment kernels (sm_100a).
//
// Execution model used by every kernel in this library: ONE WARP owns ONE
// environment instance.  With the default launch shape a CTA is exactly one
// warp (one CTA per environment); EPC > 1 packs EPC such warps into a CTA.
// All per-environment state lives in registers of that warp; cross-column
// reductions are warp-level (shfl / ballot / redux.sync), never block-level.
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
constexpr int kMaxBlocks = 64;

struct DevCfg {  // by-value kernel argument, derived from tapenv_config
    int B, n, dim, R, W, L, H, S;
    int strategy, hm_type, flags, ratio_mode;
    int static_rows, dyn_rows, update_time;
    int enc_len;      // encoded heightmap elements per env
};

struct StatePtrs {
    int4 *scal;          // [B] (valid, empty, nstable, k)
    int *heightmap;      // [B][cells]
    int *positions;      // [B][n][dim]
    int *blocks;         // [B][n][dim]
    unsigned char *stable;  // [B][n]
    int *flags;          // [B]
};

struct Scal { int valid, empty, nstable, k; };

// ---- streaming 128-bit global accesses (read-once / write-once data) ----
__device__ __forceinline__ float4 ldg_stream4(const float4 *p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_stream4(float4 *p, const float4 &v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ int warp_max(int v) { return __reduce_max_sync(TAPENV_FULL_MASK, v); }
__device__ __forceinline__ int warp_min(int v) { return __reduce_min_sync(TAPENV_FULL_MASK, v); }
__device__ __forceinline__ unsigned warp_or(unsigned v) { return __reduce_or_sync(TAPENV_FULL_MASK, v); }

// First-maximum argmax over positive finite fp64 scores held one per lane.
// `valid` lanes only; ties resolved by the smallest `key` (the reference takes the
// FIRST maximum in EMS order, np.argmax tools.py:2162).  Positive doubles order
// like their bit patterns.  Returns the winning key (0xffffffff when no lane is valid).
__device__ __forceinline__ unsigned warp_argmax_first(bool valid, double score, unsigned key) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(score);
    const unsigned hi = (unsigned)(bits >> 32), lo = (unsigned)bits;
    const unsigned mh = __reduce_max_sync(TAPENV_FULL_MASK, valid ? hi : 0u);
    const bool c1 = valid && hi == mh;
    const unsigned ml = __reduce_max_sync(TAPENV_FULL_MASK, c1 ? lo : 0u);
    const bool c2 = c1 && lo == ml;
    return __reduce_min_sync(TAPENV_FULL_MASK, c2 ? key : 0xffffffffu);
}

// C+P+S score of one candidate, IEEE fp64 exactly as the reference evaluates it
// (tools.py:2124-2140, :2161): true divisions of integers, summed left to right.
__device__ __forceinline__ double cps_score(int flags, int valid_new, long long bbox, int empty_new,
                                            int stable_cnt, int k) {
    const double vd = (double)valid_new;
    const double c = vd / (double)bbox;
    const double p = (flags & TAPENV_RF_P) ? vd / (double)(empty_new + valid_new) : 0.0;
    const double s = (flags & TAPENV_RF_S) ? (double)stable_cnt / (double)(k + 1) : 0.0;
    return (c + p) + s;
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

}  // namespace tapenv
