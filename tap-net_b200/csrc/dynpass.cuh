// dynpass.cuh -- the precedence-tensor pass of one environment (update_dynamic + the reductions of
// update_mask), one warp per environment.
//
// `dynamic` f32 [dyn_rows, S] is three bands (move | rot-small | rot-large, pack.py:195) of n rows.
// Lane mapping (fast path, S % 4 == 0, 16-byte aligned): a row is SV = S/4 128-bit vectors; a warp covers
// RP = 32/SV rows per pass with lane = rsub*SV + cv, so every lane keeps ONE column group `cv` for the
// whole pass sequence and the three band accumulators are plain bitwise ORs of the raw words:
//
//   for pass p (rows p*RP + rsub of every band), for band b:
//       v = ld.global.nc.v4 (coalesced: lanes 0..RP*SV-1 read RP*S*4 contiguous bytes)
//       if (row == real) v = 0                     pack.py:370-374  (all `update_time` bands)
//       acc[b] |= v                                column "any non-zero" == column sum != 0 (entries >= 0)
//       st.global.v4 (out-of-place copy, pack.py:370 clone)
//
// After the last pass each lane turns acc[b] into a 4-bit nibble, shifts it to its column group and the
// warp ORs the words with redux.sync: bit j of `move/small/large` = column j has a non-zero entry in that
// band -- exactly what pack.py:324-329 needs (move_sum + small_sum*large_sum != 0).
//
// With a compile-time Shape (NT > 0) every bound, stride and divisor is a constant and the pass unrolls
// into straight-line code (C2: 6 vectors per lane, all loads issued before the first use).
#pragma once
#include "tapenv_common.cuh"

namespace tapenv {

struct BandBits {   // warp-uniform after combine(): bit j = column j has a non-zero entry in that band
    unsigned long long move, small, large;
    // pack.py:327-329: dynamic_mask = small_sum*large_sum + move_sum ; blocked where != 0
    __device__ __forceinline__ unsigned long long blocked() const { return move | (small & large); }
};

__device__ __forceinline__ unsigned nz_bits(unsigned w) { return (w & 0x7fffffffu) != 0u ? 1u : 0u; }   // -0.0 == 0

// CHP = passes per chunk: all loads of a chunk (CHP * 3 vectors per lane) are issued before the first use.
// load() and process() are separate so that a kernel can put chunk 0 in flight BEFORE it waits for the
// pointer / block id it needs to process it.
// TPE = threads that share one environment's tensor (32: one warp per environment, `lane` is the lane; 32*CW: the CW copy
// warps of a CTA-per-environment launch, `lane` is the thread index inside that group -- a row pass then covers
// TPE/SV rows, so the same tensor needs proportionally fewer dependent chunks per thread).
template <class SH, int CHP, int TPE = 32>
struct DynPassFast {
    uint4 acc[3];
    uint4 v[CHP][3];
    int rsub, cv;
    bool lane_on;

    __device__ __forceinline__ static int RPg(const DevCfg &c) { return TPE == 32 ? SH::RP(c) : (SH::fixed ? TPE / SH::SVc : TPE / c.SV); }
    __device__ __forceinline__ static int PBg(const DevCfg &c) {
        return TPE == 32 ? SH::PB(c) : (SH::fixed ? (SH::NTc + TPE / SH::SVc - 1) / (TPE / SH::SVc) : (c.n + RPg(c) - 1) / RPg(c));
    }

    __device__ __forceinline__ void init(const DevCfg &c, int lane) {
        acc[0] = acc[1] = acc[2] = make_uint4(0u, 0u, 0u, 0u);
        rsub = SH::div_SV(c, lane);
        cv = lane - rsub * SH::SV(c);              // vector index of (band 0, pass 0) for this lane == lane
        lane_on = rsub < RPg(c);
    }

    __device__ __forceinline__ bool on(const DevCfg &c, int p) const {
        return lane_on && p * RPg(c) + rsub < SH::n(c) && p < PBg(c);
    }

    __device__ __forceinline__ void load(const DevCfg &c, const float *din, int lane, int p0) {
        const uint4 *src = reinterpret_cast<const uint4 *>(din) + lane;
        const int pstride = RPg(c) * SH::SV(c), bstride = SH::n(c) * SH::SV(c), nb = SH::nbands(c);
#pragma unroll
        for (int i = 0; i < CHP; ++i) {
            const bool o = on(c, p0 + i);
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                v[i][b] = make_uint4(0u, 0u, 0u, 0u);
                if (o && b < nb) v[i][b] = ldg_stream4(src + (p0 + i) * pstride + b * bstride);
            }
        }
    }

    // real < 0: no row is zeroed.  dout == nullptr: no copy is written.
    __device__ __forceinline__ void process(const DevCfg &c, float *dout, int lane, int p0, int real) {
        uint4 *dst = reinterpret_cast<uint4 *>(dout) + lane;
        const int RP = RPg(c), nb = SH::nbands(c), ut = SH::update_time(c);
        const int pstride = RP * SH::SV(c), bstride = SH::n(c) * SH::SV(c);
        const int zrow = real - rsub;                        // pass p zeroes this lane's row iff p*RP == zrow
#pragma unroll
        for (int i = 0; i < CHP; ++i) {
            const bool o = on(c, p0 + i);
            const bool zero = (p0 + i) * RP == zrow && real >= 0;
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                if (zero && b < ut) v[i][b] = make_uint4(0u, 0u, 0u, 0u);
                acc[b].x |= v[i][b].x; acc[b].y |= v[i][b].y; acc[b].z |= v[i][b].z; acc[b].w |= v[i][b].w;
                if (dout && o && b < nb) stg_stream4(dst + (p0 + i) * pstride + b * bstride, v[i][b]);
            }
        }
    }

    // chunk 0 must already be loaded
    __device__ __forceinline__ void finish(const DevCfg &c, const float *din, float *dout, int lane, int real) {
        process(c, dout, lane, 0, real);
#pragma unroll
        for (int p0 = CHP; p0 < PBg(c); p0 += CHP) {
            load(c, din, lane, p0);
            process(c, dout, lane, p0, real);
        }
    }

    __device__ __forceinline__ BandBits combine(const DevCfg &c) const {
        BandBits out;
        unsigned long long w[3];
        const int sh = cv * 4;
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const unsigned nib = nz_bits(acc[b].x) | (nz_bits(acc[b].y) << 1) | (nz_bits(acc[b].z) << 2) | (nz_bits(acc[b].w) << 3);
            if (SH::S(c) <= 32) {
                w[b] = warp_or(nib << sh);
            } else {
                const unsigned long long word = (unsigned long long)nib << sh;
                const unsigned lo = warp_or((unsigned)word), hi = warp_or((unsigned)(word >> 32));
                w[b] = ((unsigned long long)hi << 32) | lo;
            }
        }
        out.move = w[0]; out.small = w[1]; out.large = w[2];
        return out;
    }
};

// Slow generic path (S not a multiple of 4, or misaligned tensors): scalar accesses, row-major sweep.
__device__ __forceinline__ BandBits dynpass_scalar(const DevCfg &c, int lane, const float *din, float *dout, int real) {
    unsigned long long w[3] = {0ull, 0ull, 0ull};
    const int total = c.dyn_rows * c.S;
    for (int q = lane; q < total; q += 32) {
        const int row = q / c.S, col = q - row * c.S;
        const int band = row / c.n, rin = row - band * c.n;
        unsigned v = __float_as_uint(__ldg(din + q));
        if (real >= 0 && rin == real && band < c.update_time) v = 0u;
        if (band < 3) w[band] |= (unsigned long long)nz_bits(v) << col;
        if (dout) dout[q] = __uint_as_float(v);
    }
    BandBits out;
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        const unsigned lo = warp_or((unsigned)w[b]), hi = warp_or((unsigned)(w[b] >> 32));
        w[b] = ((unsigned long long)hi << 32) | lo;
    }
    out.move = w[0]; out.small = w[1]; out.large = w[2];
    return out;
}

// One entry point for both paths.
template <class SH, bool FAST>
__device__ __forceinline__ BandBits dynpass(const DevCfg &c, int lane, const float *din, float *dout, int real) {
    if (FAST) {
        DynPassFast<SH, 2> pass;
        pass.init(c, lane);
        pass.load(c, din, lane, 0);
        pass.finish(c, din, dout, lane, real);
        return pass.combine(c);
    }
    return dynpass_scalar(c, lane, din, dout, real);
}

// chosen_mask / new_mask of pack.update_mask (pack.py:318-331) for one environment.
// have_in == false -> ones (initial mask, model.py:297-307); realm < 0 -> nothing cleared.
// m0 / m1: this lane's mask_in[lane] / mask_in[lane+32], loaded early by the caller.
template <class SH>
__device__ __forceinline__ void mask_pass(const DevCfg &c, int lane, bool have_in, float m0, float m1, int realm,
                                          unsigned long long blocked, float *new_out, float *chosen_out) {
    const int S = SH::S(c);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int j = lane + 32 * half;
        if (S > 32 * half && j < S) {
            float m = have_in ? (half ? m1 : m0) : 1.0f;
            if (realm >= 0 && SH::mod_n(c, j) == realm) m = 0.0f;     // j == real + n*i, i < R (pack.py:320-321)
            if (chosen_out) chosen_out[j] = m;
            const unsigned bw = half ? (unsigned)(blocked >> 32) : (unsigned)blocked;
            if (new_out) new_out[j] = ((bw >> lane) & 1u) ? 0.0f : m;
        }
    }
}

}  // namespace tapenv
