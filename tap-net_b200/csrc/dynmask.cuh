// dynmask.cuh -- precedence tensor pass (update_dynamic + the reductions of update_mask),
// one warp per environment.
//
// `dynamic` f32 [dyn_rows, S] of one environment is streamed with 128-bit
// accesses (VW = 4; 64-/32-bit when S is not a multiple of 4), CH vectors per lane
// in flight.  While it passes through registers the warp
//   * zeroes the rows  real + n*i, i < update_time      (pack.py:370-374),
//   * writes the out-of-place copy                      (pack.py:370 clone),
//   * ORs a "non-zero" bit per (band, column) into three 64-bit words -- the
//     move / rot-small / rot-large column sums of pack.py:324-326 reduced to
//     what update_mask needs (entries are non-negative 0/1, see tapenv.h).
// The words are combined across lanes with redux.sync.or.
#pragma once
#include "tapenv_common.cuh"

namespace tapenv {

template <int VW> struct Vec;
template <> struct Vec<4> {
    float4 v;
    __device__ __forceinline__ void load(const float *p) { v = ldg_stream4(reinterpret_cast<const float4 *>(p)); }
    __device__ __forceinline__ void store(float *p) const { stg_stream4(reinterpret_cast<float4 *>(p), v); }
    __device__ __forceinline__ void zero() { v = make_float4(0.f, 0.f, 0.f, 0.f); }
    __device__ __forceinline__ unsigned nz() const {
        return (v.x != 0.f ? 1u : 0u) | (v.y != 0.f ? 2u : 0u) | (v.z != 0.f ? 4u : 0u) | (v.w != 0.f ? 8u : 0u);
    }
};
template <> struct Vec<2> {
    float2 v;
    __device__ __forceinline__ void load(const float *p) { v = __ldg(reinterpret_cast<const float2 *>(p)); }
    __device__ __forceinline__ void store(float *p) const { *reinterpret_cast<float2 *>(p) = v; }
    __device__ __forceinline__ void zero() { v = make_float2(0.f, 0.f); }
    __device__ __forceinline__ unsigned nz() const { return (v.x != 0.f ? 1u : 0u) | (v.y != 0.f ? 2u : 0u); }
};
template <> struct Vec<1> {
    float v;
    __device__ __forceinline__ void load(const float *p) { v = __ldg(p); }
    __device__ __forceinline__ void store(float *p) const { *p = v; }
    __device__ __forceinline__ void zero() { v = 0.f; }
    __device__ __forceinline__ unsigned nz() const { return v != 0.f ? 1u : 0u; }
};

struct BandBits {   // per-lane partial, then warp-combined: bit j of a word = column j has a non-zero entry in that band
    unsigned long long move, small, large;
    __device__ __forceinline__ void clear() { move = small = large = 0ull; }
    __device__ __forceinline__ void combine(int S) {
        unsigned lo, hi;
        lo = warp_or((unsigned)move);  hi = S > 32 ? warp_or((unsigned)(move >> 32)) : 0u;  move = ((unsigned long long)hi << 32) | lo;
        lo = warp_or((unsigned)small); hi = S > 32 ? warp_or((unsigned)(small >> 32)) : 0u; small = ((unsigned long long)hi << 32) | lo;
        lo = warp_or((unsigned)large); hi = S > 32 ? warp_or((unsigned)(large >> 32)) : 0u; large = ((unsigned long long)hi << 32) | lo;
    }
    // pack.py:327-329: dynamic_mask = small_sum*large_sum + move_sum ; blocked where != 0
    __device__ __forceinline__ unsigned long long blocked() const { return move | (small & large); }
};

// One tile = CH vectors per lane starting at vector index `base`.
template <int VW, int CH>
struct DynTile {
    Vec<VW> v[CH];

    __device__ __forceinline__ void load(const float *din, int base, int lane, int total) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            const int q = base + i * 32 + lane;
            if (q < total) v[i].load(din + (size_t)q * VW);
        }
    }

    // real < 0: no row is zeroed.  dout == nullptr: no copy is written.
    __device__ __forceinline__ void process(const DevCfg &c, unsigned sv_magic, int SV, float *dout, int base, int lane,
                                            int total, int real, BandBits &bits) {
        const int n = c.n;
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            const int q = base + i * 32 + lane;
            if (q < total) {
                const int row = (int)__umulhi((unsigned)q, sv_magic);   // q / SV (exact for q < 2^16)
                const int cv = q - row * SV;
                const int band = (row >= n ? 1 : 0) + (row >= 2 * n ? 1 : 0);
                const int rin = row - band * n;
                if (rin == real && band < c.update_time) v[i].zero();
                const unsigned long long b = (unsigned long long)v[i].nz() << (cv * VW);
                bits.move |= band == 0 ? b : 0ull;
                bits.small |= band == 1 ? b : 0ull;
                bits.large |= band == 2 ? b : 0ull;
                if (dout) v[i].store(dout + (size_t)q * VW);
            }
        }
    }
};

// chosen_mask / new_mask of pack.update_mask (pack.py:318-331) for one environment.
// mask_in == nullptr -> ones (initial mask, model.py:297-307); realm < 0 -> nothing cleared.
__device__ __forceinline__ void mask_pass(const DevCfg &c, int lane, const float *mask_in, int realm,
                                          unsigned long long blocked, float *new_out, float *chosen_out) {
    for (int j = lane; j < c.S; j += 32) {
        float m = mask_in ? mask_in[j] : 1.0f;
        if (realm >= 0 && (j % c.n) == realm) m = 0.0f;          // real + n*i, i < R (pack.py:320-321)
        if (chosen_out) chosen_out[j] = m;
        if (new_out) new_out[j] = ((blocked >> j) & 1ull) ? 0.0f : m;
    }
}

}  // namespace tapenv
