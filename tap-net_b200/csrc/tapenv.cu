// tapenv.cu -- kernels and C ABI of the B200-native TAP packing-environment step.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -shared -Xcompiler -fPIC
// No CPU fallback exists: every entry point launches a CUDA kernel or returns an error.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <math.h>
#include <stdlib.h>
#include <stddef.h>

#include "../../include/tapenv.h"
#include "tapenv_common.cuh"
#include "dynpass.cuh"
#include "place_lbg2d.cuh"
#include "place_lbg3d.cuh"
#include "place_macs2d.cuh"
#include "place_lb.cuh"
#include "place_macs3d.cuh"
#include "window.cuh"

// Speculative load of the edge-length rows of `static` in the warp-per-environment step: removes the DRAM round trip of the
// gather behind `ptr` at the price of (dim*S - dim) extra floats per environment.  Build-time A/B switch.  Measured
// (profiles/r02z_spec_gather_ab.txt): within +-3 % of the dependent gather at every batch size and workload, no consistent
// sign -- the round trip hides behind the precedence tensor that is in flight anyway -- so it is OFF (no extra bytes).
#ifndef TAPENV_SPEC_GATHER
#define TAPENV_SPEC_GATHER 0
#endif

namespace tapenv {

enum { STRAT_LBG2D = 0, STRAT_LBG3D = 1, STRAT_MACS2D = 2, STRAT_LB = 3, STRAT_MACS3D = 4 };

// strategies that keep a voxel grid + interval lists in the state (the warp works on level masks of it, place_lb.cuh / place_macs3d.cuh)
static bool voxel_state(const tapenv_config *c) { return c->strategy == TAPENV_LB || (c->strategy == TAPENV_MACS && c->dim == 3); }
static int lcap_of(const tapenv_config *c) { const int a = c->capacity + 2, b = c->width + 4; return a > b ? a : b; }

// ------------------------------------------------------------------------------------
// host-side helpers
// ------------------------------------------------------------------------------------
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int cells_of(const tapenv_config *c) { return c->dim == 2 ? c->width : c->width * c->length; }

static int enc_len_of(const tapenv_config *c) {
    const int cells = cells_of(c);
    if (c->dim == 2) return c->heightmap_type == TAPENV_HM_DIFF ? c->width - 1 : c->width;
    return c->heightmap_type == TAPENV_HM_DIFF ? 2 * cells : cells;
}

static void layout_of(const tapenv_config *c, tapenv_state_layout *L) {
    const size_t B = (size_t)c->batch, cap = (size_t)c->capacity, dim = (size_t)c->dim;
    size_t off = 0;
    L->scalars = off;   off = align_up(off + B * 4 * sizeof(int32_t), 256);
    L->heightmap = off; off = align_up(off + B * (size_t)cells_of(c) * sizeof(int32_t), 256);
    L->positions = off; off = align_up(off + B * cap * dim * sizeof(int32_t), 256);
    L->blocks = off;    off = align_up(off + B * cap * dim * sizeof(int32_t), 256);
    L->stable = off;    off = align_up(off + B * cap, 256);
    L->flags = off;     off = align_up(off + B * sizeof(int32_t), 256);
    const bool lb = voxel_state(c);
    const size_t nlists = (size_t)c->height * (dim == 3 ? (size_t)c->length : 1);
    L->voxels = off;    off = align_up(off + (lb ? B * (size_t)cells_of(c) * (size_t)c->height * sizeof(int16_t) : 0), 256);
    L->lists = off;     off = align_up(off + (lb ? B * nlists * (size_t)lcap_of(c) : 0), 256);
    L->pending = off;   off = align_up(off + (lb ? B * 4 * sizeof(float) : 0), 256);
    L->total = off;
}

static DevCfg devcfg_of(const tapenv_config *c) {
    DevCfg d;
    d.B = c->batch; d.n = c->blocks_num; d.dim = c->dim; d.R = c->rotate_types;
    d.W = c->width; d.L = c->length; d.H = c->height; d.S = c->blocks_num * c->rotate_types;
    d.cap = c->capacity;
    d.strategy = c->strategy; d.hm_type = c->heightmap_type; d.flags = c->reward_flags; d.ratio_mode = c->ratio_mode;
    d.static_rows = c->static_rows; d.dyn_rows = c->dyn_rows; d.update_time = c->update_time;
    d.enc_len = enc_len_of(c);
    // geometry of the 128-bit precedence pass (dynpass.cuh); SV = 1 keeps the divisions defined when S % 4 != 0
    d.SV = d.S % 4 == 0 && d.S >= 4 ? d.S / 4 : 1;
    d.RP = 32 / d.SV > 0 ? 32 / d.SV : 1;
    d.PB = (d.n + d.RP - 1) / d.RP;
    d.nbands = d.dyn_rows / d.n;
    d.inv_SV = (65536u + d.SV - 1) / d.SV;
    d.inv_n = (65536u + d.n - 1) / d.n;
    d.inv_L = (65536u + d.L - 1) / (d.L > 0 ? d.L : 1);
    d.dyn_env = (unsigned)(d.dyn_rows * d.S);
    d.static_env = (unsigned)(d.static_rows * d.S);
    d.lcap = lcap_of(c);
    d.nlists = c->height * (c->dim == 3 ? c->length : 1);
    return d;
}

static StatePtrs stateptrs_of(const tapenv_config *c, void *state) {
    tapenv_state_layout L; layout_of(c, &L);
    char *base = (char *)state;
    StatePtrs s;
    s.scal = (int4 *)(base + L.scalars);
    s.heightmap = (int *)(base + L.heightmap);
    s.positions = (int *)(base + L.positions);
    s.blocks = (int *)(base + L.blocks);
    s.stable = (unsigned char *)(base + L.stable);
    s.flags = (int *)(base + L.flags);
    s.voxels = (short *)(base + L.voxels);
    s.lists = (unsigned char *)(base + L.lists);
    s.pending = (float *)(base + L.pending);
    return s;
}

static int check_cfg(const tapenv_config *c) {
    if (!c) return TAPENV_EINVAL;
    if (c->batch < 0 || c->blocks_num < 1 || c->width < 1 || c->height < 1) return TAPENV_EINVAL;
    if (c->dim != 2 && c->dim != 3) return TAPENV_EINVAL;
    if (c->dim == 2 && c->length != 1) return TAPENV_ESHAPE;
    if (c->dim == 3 && c->length < 1) return TAPENV_EINVAL;
    if (c->rotate_types < 1) return TAPENV_EINVAL;
    if (c->strategy != TAPENV_LB_GREEDY && c->strategy != TAPENV_MACS && c->strategy != TAPENV_LB) return TAPENV_EENUM;
    if (voxel_state(c) && (c->capacity > 250 || c->height > 32767)) return TAPENV_ELIMIT;
    if (c->strategy == TAPENV_MACS && c->dim == 3 && (c->height > 255 || c->width > 120)) return TAPENV_ELIMIT;   // EMS coordinates are bytes
    if (c->heightmap_type < 0 || c->heightmap_type > 2) return TAPENV_EENUM;
    if (c->ratio_mode < 0 || c->ratio_mode > TAPENV_RATIO_CP_HALF) return TAPENV_EENUM;
    if (c->static_rows < 1 + c->dim) return TAPENV_ESHAPE;
    if (c->dyn_rows != c->blocks_num && c->dyn_rows != 3 * c->blocks_num && c->dyn_rows != c->blocks_num + 1) return TAPENV_ESHAPE;
    if (c->update_time != 1 && c->update_time != 3) return TAPENV_ESHAPE;
    if (c->update_time * c->blocks_num > c->dyn_rows) return TAPENV_ESHAPE;
    if (c->capacity < 1) return TAPENV_ESHAPE;
    if (c->dim == 2 && c->width > kMaxWidth2D) return TAPENV_ELIMIT;
    if (c->dim == 3 && c->width * c->length > kMaxCells3D) return TAPENV_ELIMIT;
    if (c->blocks_num * c->rotate_types > kMaxCandidates) return TAPENV_ELIMIT;
    if (c->blocks_num > kMaxBlocks) return TAPENV_ELIMIT;
    if (c->strategy == TAPENV_MACS && c->capacity > kMaxBlocks) return TAPENV_ELIMIT;
    if (c->height > (1 << 18)) return TAPENV_ELIMIT;
    return TAPENV_OK;
}

// The legacy 'rot-old' layout (pack.py:218-223): dynamic = n movement rows + ONE rotate-state row.  pack.update_dynamic /
// pack.update_mask / the initial mask are defined for it (and served: the generic row sweep), but the reference's own decode
// loop is not -- model.py:391-392 gathers the block from ALL 1+dim rows of `static`, so Container.add_new_block receives
// (index, w, h) and `block_x, block_z = block` (tools.py:2060) raises.  The fused entry points say EUNSUPPORTED.
static bool rot_old_layout(const tapenv_config *c) { return c->dyn_rows == c->blocks_num + 1; }

static int strategy_kernel(const tapenv_config *c) {
    if (c->strategy == TAPENV_LB_GREEDY) return c->dim == 2 ? STRAT_LBG2D : STRAT_LBG3D;
    if (c->strategy == TAPENV_MACS && c->dim == 2) return STRAT_MACS2D;
    if (c->strategy == TAPENV_MACS && c->dim == 3) return STRAT_MACS3D;
    if (c->strategy == TAPENV_LB) return STRAT_LB;
    return -1;
}

static int launch_status() { return cudaGetLastError() == cudaSuccess ? TAPENV_OK : TAPENV_ECUDA; }

// Programmatic dependent launch: every kernel of this library is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, waits for its predecessor's memory with
// griddepcontrol.wait before its first global access and only then lets ITS dependent start launching, so the
// launch latency and CTA ramp of decode step t+1 overlap the tail of step t (at most one grid is ever parked).
// TAPENV_PDL=0 in the environment turns the attribute off (plain stream order).
static bool pdl_enabled() {
    static const bool on = [] { const char *e = getenv("TAPENV_PDL"); return !(e && e[0] == '0'); }();
    return on;
}

template <typename... KArgs, typename... Args>
static void launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ void grid_dependency_sync() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// ------------------------------------------------------------------------------------
// device: per-environment pieces
// ------------------------------------------------------------------------------------
__device__ __forceinline__ int env_index(int &lane, int &warp) {
    lane = threadIdx.x & 31;
    warp = threadIdx.x >> 5;
    return blockIdx.x * kWarpsPerCta + warp;
}

__device__ __forceinline__ Scal load_scal(const StatePtrs &st, int b) {
    const int4 v = st.scal[b];
    Scal s; s.valid = v.x; s.empty = v.y; s.nstable = v.z; s.k = v.w;
    return s;
}

// Everything a placement needs from the state buffer, loaded up-front so the requests overlap the
// precedence-tensor traffic.
template <int STRAT>
struct EnvRegs {
    int h, x, y;
    Scal sc;
    MacsHist hist;

    __device__ __forceinline__ void load(const DevCfg &c, const StatePtrs &st, int b, int lane) {
        if (STRAT == STRAT_LB || STRAT == STRAT_MACS3D) return;   // voxel-state strategies keep their state in global memory (voxel_add_block)
        const int cells = (STRAT == STRAT_LBG3D) ? c.W * c.L : c.W;
        h = lane < cells ? st.heightmap[(size_t)b * cells + lane] : 0;
        sc = load_scal(st, b);
        x = 0; y = 0;
        if (STRAT == STRAT_LBG3D) { x = (int)(((unsigned)lane * c.inv_L) >> 16); y = lane - x * c.L; }
        if (STRAT == STRAT_MACS2D) {
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const int i = lane + 32 * s;
                hist.x[s] = hist.z[s] = hist.xx[s] = hist.zz[s] = 0;
                if (i < c.cap) {                     // rows >= k are zero since the reset and masked by the scan (i < k): no wait on `scal`
                    const int2 p = *reinterpret_cast<const int2 *>(st.positions + ((size_t)b * c.cap + i) * 2);
                    const int2 q = *reinterpret_cast<const int2 *>(st.blocks + ((size_t)b * c.cap + i) * 2);
                    hist.x[s] = p.x; hist.z[s] = p.y; hist.xx[s] = q.x; hist.zz[s] = q.y;
                }
            }
        }
    }
};

// Container.calc_CPS + calc_ratio (tools.py:3887-3966) in IEEE fp64, operation for operation
__device__ __forceinline__ double calc_ratio_dev(const DevCfg &c, int valid, int empty, int nstable, int k, int height) {
    const int cells = c.dim == 2 ? c.W : c.W * c.L;
    double C = 0.0, P = 0.0, S = 0.0;              // current_blocks_num == 0 -> 0, 0, 0 (tools.py:3888-3889)
    if (k != 0) {
        C = __ddiv_rn((double)valid, (double)(cells * height));
        P = __ddiv_rn((double)valid, (double)(empty + valid));
        S = __ddiv_rn((double)nstable, (double)k);
    }
    switch (c.ratio_mode) {
        case TAPENV_RATIO_C: return __ddiv_rn(C, 3.0);
        case TAPENV_RATIO_CS: return __ddiv_rn(__dmul_rn(C, S), 3.0);
        case TAPENV_RATIO_C_P: return __ddiv_rn(__dadd_rn(C, P), 3.0);
        case TAPENV_RATIO_CP_S: return __ddiv_rn(__dmul_rn(__dadd_rn(C, P), S), 3.0);
        case TAPENV_RATIO_2C_SUM: return __ddiv_rn(__dadd_rn(__dadd_rn(__dmul_rn(2.0, C), P), S), 3.0);
        case TAPENV_RATIO_CPS: return __ddiv_rn(__dmul_rn(__dmul_rn(C, P), S), 3.0);
        case TAPENV_RATIO_CP_HALF: return __ddiv_rn(__dadd_rn(C, P), 2.0);
        default: return __ddiv_rn(__dadd_rn(__dadd_rn(C, P), S), 3.0);
    }
}

template <int STRAT>
__device__ __forceinline__ void voxel_add_block(const DevCfg &c, const StatePtrs &st, int b, int lane, int bx, int by, int bz,
                                                float *dec_dyn, int extra_flags, float *reward);

// block dimension of a strategy instantiation: fixed for the heightmap-only strategies, from the config for the voxel ones
template <int STRAT>
__device__ __forceinline__ int dim_of(const DevCfg &c) {
    return (STRAT == STRAT_LB || STRAT == STRAT_MACS3D) ? c.dim : (STRAT == STRAT_LBG3D ? 3 : 2);
}

// get_heightmap() of the state as it stands (tools.py:3824-3856), warp-parallel
__device__ __forceinline__ int encode_state_heightmap(const DevCfg &c, const StatePtrs &st, int b, int lane, int dim, float *out) {
    const int cells = dim == 2 ? c.W : c.W * c.L;
    const int h = lane < cells ? *((volatile int *)(st.heightmap + (size_t)b * cells + lane)) : 0;
    if (out) {
        if (dim == 3) {
            const int x = (int)(((unsigned)lane * c.inv_L) >> 16), y = lane - x * c.L;
            encode_heightmap_3d(c, lane, x, y, h, out);
        } else encode_heightmap_2d(c, lane, h, out);
    }
    return h;
}

// Container.__init__ / clear_container of one environment (tools.py:3611-3661, :3858-3885), warp-parallel
__device__ __forceinline__ void reset_env_state(const DevCfg &c, const StatePtrs &st, int b, int lane) {
    const int cells = c.dim == 2 ? c.W : c.W * c.L;
    for (int i = lane; i < cells; i += 32) st.heightmap[(size_t)b * cells + i] = 0;
    for (int i = lane; i < c.cap * c.dim; i += 32) { st.positions[(size_t)b * c.cap * c.dim + i] = 0; st.blocks[(size_t)b * c.cap * c.dim + i] = 0; }
    for (int i = lane; i < c.cap; i += 32) st.stable[(size_t)b * c.cap + i] = 0;
    if (lane == 0) { st.scal[b] = make_int4(0, 0, 0, 0); st.flags[b] = 0; }
    if (c.strategy == TAPENV_LB) {                   // voxel grid zero, every x list = [0] (tools.py:3649-3653)
        const size_t nv = (size_t)cells * c.H;
        for (size_t i = lane; i < nv; i += 32) st.voxels[(size_t)b * nv + i] = 0;
        const size_t nl = (size_t)c.nlists * c.lcap;
        for (size_t i = lane; i < nl; i += 32) st.lists[(size_t)b * nl + i] = (i % c.lcap) == 0 ? 1 : 0;
    } else if (c.strategy == TAPENV_MACS && c.dim == 3) {   // every (level, row) interval list = [0, W-1] (tools.py:3644-3648)
        const size_t nv = (size_t)cells * c.H;
        for (size_t i = lane; i < nv; i += 32) st.voxels[(size_t)b * nv + i] = 0;
        const size_t nl = (size_t)c.nlists * c.lcap;
        for (size_t i = lane; i < nl; i += 32) {
            const int q = (int)(i % c.lcap);
            st.lists[(size_t)b * nl + i] = q == 0 ? 2 : (q == 2 ? (unsigned char)(c.W - 1) : 0);
        }
    }
}

// Container.add_new_block for one environment (tools.py:3663-3744): placement, commit,
// current_blocks_num += 1 even when the placement failed (tools.py:3713), heightmap encoding.
template <int STRAT, bool SMALLN = false>             // SMALLN: blocks_num <= 32 known at compile time (MACS: one history slot)
__device__ __forceinline__ void container_add_block(const DevCfg &c, const StatePtrs &st, int b, int lane,
                                                    EnvRegs<STRAT> &e, int bx, int by, int bz, float *dec_dyn,
                                                    unsigned *ems_keys, int extra_flags, float *reward = nullptr) {
    if (STRAT == STRAT_LB || STRAT == STRAT_MACS3D) {   // voxel-state strategies: the warp works on level masks of the grid
        voxel_add_block<STRAT>(c, st, b, lane, bx, by, bz, dec_dyn, extra_flags, reward);
        return;
    }
    const int dim = (STRAT == STRAT_LBG3D) ? 3 : 2;
    const int cells = (STRAT == STRAT_LBG3D) ? c.W * c.L : c.W;
    int anomaly = extra_flags;
    if (e.sc.k >= c.cap) {       // the reference raises IndexError (rotate_state[k], tools.py:3677)
        anomaly |= 2;
    } else {
        PlaceOut r;
        r.placed = 0; r.x = r.y = r.z = r.stable = r.top = 0;
        if (STRAT == STRAT_LBG2D) r = lbg2d_place(c, lane, bx, bz, e.h, e.sc);
        else if (STRAT == STRAT_LBG3D) r = lbg3d_place(c, lane, e.x, e.y, bx, by, bz, e.h, e.sc);
        else r = macs2d_place<SMALLN>(c, lane, bx, bz, e.h, e.sc, e.hist, ems_keys, anomaly);
        if (lane < cells) st.heightmap[(size_t)b * cells + lane] = e.h;
        if (lane == 0) {
            const size_t o = ((size_t)b * c.cap + e.sc.k) * dim;
            if (dim == 2) {
                *reinterpret_cast<int2 *>(st.blocks + o) = make_int2(bx, bz);
                if (r.placed) *reinterpret_cast<int2 *>(st.positions + o) = make_int2(r.x, r.z);
            } else {
                st.blocks[o] = bx; st.blocks[o + 1] = by; st.blocks[o + 2] = bz;
                if (r.placed) { st.positions[o] = r.x; st.positions[o + 1] = r.y; st.positions[o + 2] = r.z; }
            }
            st.stable[(size_t)b * c.cap + e.sc.k] = (unsigned char)r.stable;
            // a stack above the container: the reference's voxel grid silently clips here and raises
            // IndexError the next time it touches that column -- flagged instead
            if (r.placed && r.top > c.H) anomaly |= 1;
            st.scal[b] = make_int4(e.sc.valid, e.sc.empty, e.sc.nstable, e.sc.k + 1);
        }
    }
    if (anomaly && lane == 0) st.flags[b] |= anomaly;
    if (dec_dyn) {
        if (STRAT == STRAT_LBG3D) encode_heightmap_3d(c, lane, e.x, e.y, e.h, dec_dyn + (size_t)b * c.enc_len);
        else encode_heightmap_2d(c, lane, e.h, dec_dyn + (size_t)b * c.enc_len);
    }
    if (reward) {                // Container.calc_ratio on the state this step leaves behind (model.py:509-510 after the last step)
        const int height = warp_max(lane < cells ? e.h : 0);
        const int knew = e.sc.k >= c.cap ? e.sc.k : e.sc.k + 1;
        if (lane == 0) reward[b] = (float)calc_ratio_dev(c, e.sc.valid, e.sc.empty, e.sc.nstable, knew, height);
    }
}

// ------------------------------------------------------------------------------------
// K0 reset: Container.__init__ / clear_container + initial accessibility mask
// ------------------------------------------------------------------------------------
template <bool FAST>
__global__ void __launch_bounds__(32 * kWarpsPerCta)
reset_kernel(DevCfg c, StatePtrs st, int clear_state, const float *__restrict__ dynamic, float *__restrict__ cur_mask,
             float *__restrict__ mask) {
    typedef Shape<0, 0, 0> SH;
    int lane, warp; const int b = env_index(lane, warp);
    grid_dependency_sync();
    if (b >= c.B) return;
    if (clear_state) reset_env_state(c, st, b, lane);
    if (dynamic == nullptr) return;
    const BandBits bits = dynpass<SH, FAST>(c, lane, env_ptr(dynamic, b, c.dyn_env), nullptr, -1);
    mask_pass<SH>(c, lane, false, 0.f, 0.f, -1, bits.blocked(), cur_mask + (size_t)b * c.S, mask ? mask + (size_t)b * c.S : nullptr);
}


// ------------------------------------------------------------------------------------
// K0 for packed host formats: the dataset's tensors hold small integers (static) and 0/1 (dynamic) in fp32
// (pack.py:101-223), 2 720 B per environment at C2 of which 135 B are information.  A loader that keeps them as
// u8 / bit rows uploads 20x fewer PCIe bytes; this kernel expands them into the fp32 tensors the network and the
// step kernels consume, clears the containers and derives the initial masks (model.py:294-307) in one launch.
// Bit q = row*S + col of `bits` (u32 words, little-endian bit order) <=> dynamic[row, col] == 1.
// ------------------------------------------------------------------------------------
template <bool FAST>
__global__ void __launch_bounds__(32 * kWarpsPerCta)
unpack_reset_kernel(DevCfg c, StatePtrs st, int clear_state, const unsigned char *__restrict__ static_u8,
                    const unsigned *__restrict__ bits, int words_env, float *__restrict__ static_out,
                    float *__restrict__ dynamic_out, float *__restrict__ cur_mask, float *__restrict__ mask) {
    typedef Shape<0, 0, 0> SH;
    __shared__ uint4 lut[16];                         // nibble -> four fp32 0/1 values
    int lane, warp; const int b = env_index(lane, warp);
    if (threadIdx.x < 16) {
        const unsigned t = threadIdx.x, one = 0x3f800000u;
        lut[t] = make_uint4((t & 1u) * one, ((t >> 1) & 1u) * one, ((t >> 2) & 1u) * one, ((t >> 3) & 1u) * one);
    }
    __syncthreads();
    grid_dependency_sync();
    if (b >= c.B) return;
    const unsigned *wb = bits + (size_t)b * words_env;
    if (clear_state) {
        const int cells = c.dim == 2 ? c.W : c.W * c.L;
        for (int i = lane; i < cells; i += 32) st.heightmap[(size_t)b * cells + i] = 0;
        for (int i = lane; i < c.cap * c.dim; i += 32) { st.positions[(size_t)b * c.cap * c.dim + i] = 0; st.blocks[(size_t)b * c.cap * c.dim + i] = 0; }
        for (int i = lane; i < c.cap; i += 32) st.stable[(size_t)b * c.cap + i] = 0;
        if (lane == 0) { st.scal[b] = make_int4(0, 0, 0, 0); st.flags[b] = 0; }
    }
    const unsigned char *su = static_u8 + (size_t)b * c.static_env;
    float *so = static_out + (size_t)b * c.static_env;
    for (int i = lane; i < (int)c.static_env; i += 32) so[i] = (float)su[i];
    float *dyo = env_ptr(dynamic_out, b, c.dyn_env);
    unsigned long long w[3] = {0ull, 0ull, 0ull};
    if (FAST) {
        const int rsub = (int)(((unsigned)lane * c.inv_SV) >> 16), cv = lane - rsub * c.SV;
        const bool on = rsub < c.RP;
        unsigned nibs[3] = {0u, 0u, 0u};
        uint4 *dst = reinterpret_cast<uint4 *>(dyo) + lane;
        const int pstride = c.RP * c.SV, bstride = c.n * c.SV;
        for (int p = 0; p < c.PB; ++p) {
            const int row = p * c.RP + rsub;
            if (!(on && row < c.n)) continue;
#pragma unroll
            for (int bd = 0; bd < 3; ++bd) {
                if (bd >= c.nbands) continue;
                const int q0 = (bd * c.n + row) * c.S + 4 * cv;      // S % 4 == 0: the four bits share one word
                const int wi = q0 >> 5;
                const unsigned word = __ldg(wb + wi);                 // <= 300 B per environment: L1-resident after the first touch
                const unsigned nib = (word >> (q0 & 31)) & 0xfu;
                nibs[bd] |= nib;
                stg_stream4(dst + p * pstride + bd * bstride, lut[nib]);
            }
        }
#pragma unroll
        for (int bd = 0; bd < 3; ++bd) {
            const unsigned long long word = on ? ((unsigned long long)nibs[bd] << (4 * cv)) : 0ull;
            w[bd] = (unsigned long long)warp_or((unsigned)word) | ((unsigned long long)warp_or((unsigned)(word >> 32)) << 32);
        }
    } else {
        const int total = c.dyn_rows * c.S;
        for (int q = lane; q < total; q += 32) {
            const int row = q / c.S, col = q - row * c.S;
            const int band = row / c.n;
            const unsigned bit = (__ldg(wb + (q >> 5)) >> (q & 31)) & 1u;
            if (band < 3) w[band] |= (unsigned long long)bit << col;
            dyo[q] = bit ? 1.f : 0.f;
        }
#pragma unroll
        for (int bd = 0; bd < 3; ++bd)
            w[bd] = (unsigned long long)warp_or((unsigned)w[bd]) | ((unsigned long long)warp_or((unsigned)(w[bd] >> 32)) << 32);
    }
    BandBits bb; bb.move = w[0]; bb.small = w[1]; bb.large = w[2];
    mask_pass<SH>(c, lane, false, 0.f, 0.f, -1, bb.blocked(), cur_mask + (size_t)b * c.S, mask ? mask + (size_t)b * c.S : nullptr);
}

// ------------------------------------------------------------------------------------
// unfused pieces (signature parity with pack.update_dynamic / pack.update_mask / add_new_block)
// ------------------------------------------------------------------------------------
template <bool FAST>
__global__ void __launch_bounds__(32 * kWarpsPerCta)
update_dynamic_kernel(DevCfg c, const float *__restrict__ dynamic, const float *__restrict__ static_,
                      const int64_t *__restrict__ ptr, float *__restrict__ out) {
    typedef Shape<0, 0, 0> SH;
    int lane, warp; const int b = env_index(lane, warp);
    grid_dependency_sync();
    if (b >= c.B) return;
    long long p = ptr[b];
    if (p < 0 || p >= c.S) p = 0;                                          // the reference's index would raise
    const int real = (int)env_ptr(static_, b, c.static_env)[p];            // pack.py:347 (.long() truncates)
    dynpass<SH, FAST>(c, lane, env_ptr(dynamic, b, c.dyn_env), env_ptr(out, b, c.dyn_env), real);
}

template <bool FAST>
__global__ void __launch_bounds__(32 * kWarpsPerCta)
update_mask_kernel(DevCfg c, const float *__restrict__ mask, const float *__restrict__ dynamic,
                   const int64_t *__restrict__ ptr, float *__restrict__ new_mask, float *__restrict__ chosen_mask) {
    typedef Shape<0, 0, 0> SH;
    int lane, warp; const int b = env_index(lane, warp);
    grid_dependency_sync();
    if (b >= c.B) return;
    const float m0 = lane < c.S ? mask[(size_t)b * c.S + lane] : 0.f;
    const float m1 = lane + 32 < c.S ? mask[(size_t)b * c.S + lane + 32] : 0.f;
    long long p = ptr[b];
    if (p < 0) p = 0;
    const BandBits bits = dynpass<SH, FAST>(c, lane, env_ptr(dynamic, b, c.dyn_env), nullptr, -1);
    const int realm = (int)(p % c.n);                                      // pack.py:314-316
    mask_pass<SH>(c, lane, true, m0, m1, realm, bits.blocked(), new_mask + (size_t)b * c.S, chosen_mask + (size_t)b * c.S);
}

template <int STRAT>
__global__ void __launch_bounds__(32 * kWarpsPerCta)
add_blocks_kernel(DevCfg c, StatePtrs st, const float *__restrict__ blocks, float *__restrict__ dec_dyn) {
    unsigned *const ems_keys_none = nullptr;       // (the per-warp EMS key list of r01; the MACS scan no longer needs shared memory)
    int lane, warp; const int b = env_index(lane, warp);
    grid_dependency_sync();
    if (b >= c.B) return;
    EnvRegs<STRAT> e; e.load(c, st, b, lane);
    const float *blk = blocks + (size_t)b * c.dim;
    const int bx = (int)blk[0];                                            // .astype(int) tools.py:3689
    const int by = c.dim == 3 ? (int)blk[1] : 1;
    const int bz = (int)blk[c.dim - 1];
    container_add_block<STRAT>(c, st, b, lane, e, bx, by, bz, dec_dyn, ems_keys_none, 0);
}

// ------------------------------------------------------------------------------------
// Voxel-state strategies (LB tools.py:1602-1914, MACS 3D :2751-3165): the placement walks a voxel grid and incrementally
// edited lists kept in the state buffer.  It runs inside every kernel of the library (voxel_add_block: add_blocks / step /
// episode / rolling / mul) -- one launch per decode step.  r01: thread-per-environment kernels behind a separate tensor pass
// (117 / 340 / 13 700 us per step at B=4096 for LB 2D / LB 3D / MACS 3D); r02 first form: lane 0 of the environment's warp
// walks while the warp waits (42 / 99 / 2 940 us, profiles/r02m_voxel_ab.txt); r02 second form: the WHOLE warp on level masks
// in shared memory (place_macs3d.cuh, place_lb.cuh warp form: 31 / 84 / 370 us per fused step, profiles/r02af_bench_*.json).
// ------------------------------------------------------------------------------------
// Container.add_new_block for one environment, LB strategy, the one-thread walk (containers above kLbMaxH levels).  Returns
// the anomaly bits.
template <int DIM>
__device__ __forceinline__ int lb_env_add_block(const DevCfg &c, const StatePtrs &st, int b, int bx, int by, int bz) {
    const int cells = DIM == 2 ? c.W : c.W * c.L;
    LbState s;
    s.W = c.W; s.L = DIM == 3 ? c.L : 1; s.H = c.H; s.cells = cells; s.lcap = c.lcap;
    s.vox = st.voxels + (size_t)b * cells * c.H;
    s.lists = st.lists + (size_t)b * c.nlists * c.lcap;
    s.h = st.heightmap + (size_t)b * cells;
    const int4 sc = st.scal[b];
    int anomaly = 0;
    const int k = sc.w;
    if (k >= c.cap) return 2;
    int *positions = st.positions + (size_t)b * c.cap * DIM, *blks = st.blocks + (size_t)b * c.cap * DIM;
    blks[k * DIM] = bx; if (DIM == 3) blks[k * DIM + 1] = by; blks[k * DIM + DIM - 1] = bz;
    int4 out = make_int4(sc.x, sc.y, sc.z, k + 1);
    unsigned char stable = 0;
    if (bx >= 1 && by >= 1 && bz >= 1 && bx <= c.W && by <= s.L) {
        const int vol = bx * by * bz;
        const LbBest best = lb_place<DIM>(c.flags, s, k, positions, blks, bx, by, bz, sc.x + vol, sc.y, sc.z, anomaly);
        if (best.any) {
            lb_commit<DIM>(s, k, best, bx, by, bz, anomaly);
            if (!(anomaly & 1)) {
                positions[k * DIM] = best.x; if (DIM == 3) positions[k * DIM + 1] = best.y; positions[k * DIM + DIM - 1] = best.z;
                stable = (unsigned char)best.stable;
                out = make_int4(sc.x + vol, sc.y + best.add, sc.z + best.stable, k + 1);
            }
        }
    }
    st.stable[(size_t)b * c.cap + k] = stable;
    st.scal[b] = out;
    return anomaly;
}

// Container.add_new_block for one environment, LB strategy, executed by the WHOLE warp (place_lb.cuh, warp form; containers
// up to kLbMaxH levels).  Returns the anomaly bits (identical in every lane).
template <int DIM>
__device__ __forceinline__ int lb_env_add_block_warp_dev(const DevCfg &c, const StatePtrs &st, int b, int lane, LbScratch *scratch,
                                                         int bx, int by, int bz) {
    const int cells = DIM == 2 ? c.W : c.W * c.L;
    LbState s;
    s.W = c.W; s.L = DIM == 3 ? c.L : 1; s.H = c.H; s.cells = cells; s.lcap = c.lcap;
    s.vox = st.voxels + (size_t)b * cells * c.H;
    s.lists = st.lists + (size_t)b * c.nlists * c.lcap;
    s.h = st.heightmap + (size_t)b * cells;
    LbWarp w; w.sm = scratch; w.lane = lane; w.nl = 32;
    return lb_env_add_block_warp<DIM>(c.flags, c.cap, s, w, reinterpret_cast<int *>(st.scal + b), st.positions + (size_t)b * c.cap * DIM,
                                      st.blocks + (size_t)b * c.cap * DIM, st.stable + (size_t)b * c.cap, bx, by, bz);
}

// Container.add_new_block for one environment, MACS 3D, executed by the WHOLE warp (place_macs3d.cuh: level masks in shared
// memory, warp-uniform walk, lane-parallel reductions; lane 0 commits).  Returns the anomaly bits (identical in every lane).
__device__ __forceinline__ int macs3d_env_add_block_warp(const DevCfg &c, const StatePtrs &st, int b, int lane, M3Scratch *scratch,
                                                         int bx, int by, int bz) {
    const int cells = c.W * c.L;
    M3State s;
    s.W = c.W; s.L = c.L; s.H = c.H; s.cells = cells; s.lcap = c.lcap;
    s.vox = st.voxels + (size_t)b * cells * c.H;
    s.lists = reinterpret_cast<signed char *>(st.lists) + (size_t)b * c.nlists * c.lcap;
    s.h = st.heightmap + (size_t)b * cells;
    s.sm = scratch; s.lane = lane; s.nl = 32;
    return macs3d_env_add_block(c.flags, c.cap, s, reinterpret_cast<int *>(st.scal + b), st.positions + (size_t)b * c.cap * 3,
                                st.blocks + (size_t)b * c.cap * 3, st.stable + (size_t)b * c.cap, bx, by, bz);
}

// Container.add_new_block inside a warp-per-environment kernel: the warp works together on level masks of the voxel grid in
// shared memory (LB containers above kLbMaxH levels: lane 0 walks the grid).  Then the warp encodes (and, optionally, emits calc_ratio of the state left behind).
// STRAT: STRAT_LB (dim from the config) or STRAT_MACS3D.  Every kernel calling this runs at most kWarpsPerCta warps per CTA.
template <int STRAT>
__device__ __forceinline__ void voxel_add_block(const DevCfg &c, const StatePtrs &st, int b, int lane, int bx, int by, int bz,
                                                float *dec_dyn, int extra_flags, float *reward) {
    if constexpr (STRAT == STRAT_MACS3D) {
        __shared__ M3Scratch m3_scratch[kWarpsPerCta];
        const int anomaly = extra_flags | macs3d_env_add_block_warp(c, st, b, lane, &m3_scratch[(threadIdx.x >> 5) % kWarpsPerCta], bx, by, bz);
        if (anomaly && lane == 0) st.flags[b] |= anomaly;
    } else {
        __shared__ LbScratch lb_scratch[kWarpsPerCta];
        if (c.H <= kLbMaxH) {                            // the warp works on level masks
            LbScratch *scr = &lb_scratch[(threadIdx.x >> 5) % kWarpsPerCta];
            const int anomaly = extra_flags | (c.dim == 2 ? lb_env_add_block_warp_dev<2>(c, st, b, lane, scr, bx, 1, bz)
                                                          : lb_env_add_block_warp_dev<3>(c, st, b, lane, scr, bx, by, bz));
            if (anomaly && lane == 0) st.flags[b] |= anomaly;
        } else if (lane == 0) {                          // tall containers: lane 0 walks the grid
            int anomaly = extra_flags;
            if (c.dim == 2) anomaly |= lb_env_add_block<2>(c, st, b, bx, 1, bz);
            else anomaly |= lb_env_add_block<3>(c, st, b, bx, by, bz);
            if (anomaly) st.flags[b] |= anomaly;
        }
    }
    __syncwarp();                                    // lane 0's global writes are visible to the warp behind this barrier
    if (dec_dyn || reward) {
        const int cells = c.dim == 2 ? c.W : c.W * c.L;
        const int h = lane < cells ? *((volatile int *)(st.heightmap + (size_t)b * cells + lane)) : 0;
        if (dec_dyn) {
            if (c.dim == 3) {
                const int x = (int)(((unsigned)lane * c.inv_L) >> 16), y = lane - x * c.L;
                encode_heightmap_3d(c, lane, x, y, h, dec_dyn + (size_t)b * c.enc_len);
            } else encode_heightmap_2d(c, lane, h, dec_dyn + (size_t)b * c.enc_len);
        }
        if (reward) {
            const int height = warp_max(h);
            if (lane == 0) {
                const int4 s4 = st.scal[b];                 // written by this very thread
                reward[b] = (float)calc_ratio_dev(c, s4.x, s4.y, s4.z, s4.w, height);
            }
        }
    }
}

// ------------------------------------------------------------------------------------
// fused decode-step kernel  (K1 + K2/K3/K4): update_dynamic + update_mask + gather + add_new_block
// ------------------------------------------------------------------------------------
// PLACE_FIRST: run the placement before the precedence pass is consumed (see (3) below) -- chosen when the whole
// batch is resident in a single wave, where no other warp is left to hide this warp's load latency.
// MINB: CTAs per SM the register allocation must allow.  0 = the strategy's default (2D LB_GREEDY: 8 -> 64 registers;
// 3D / MACS: 4 -> 88 / 106 registers).  The heavy placements are also built with MINB = 7 (72 registers): that variant is
// slower per warp but lets 1036 CTAs be resident at once, so a batch of 4096 runs as ONE wave instead of 1.4
// (profiles/r01q_step_variants.txt: C3 B=4096 16.8 -> 14.4 us) -- chosen per launch in tapenv_step.
template <int STRAT, bool FAST, int NT, int RT, bool PLACE_FIRST, int MINB = 0>
__global__ void __launch_bounds__(32 * kWarpsPerCta, MINB > 0 ? MINB : (STRAT == STRAT_LBG2D ? 32 / kWarpsPerCta : 16 / kWarpsPerCta))
step_kernel(DevCfg c, StatePtrs st, const int64_t *__restrict__ ptr, const float *__restrict__ static_,
            const float *__restrict__ dynamic_in, const float *__restrict__ mask_in, float *__restrict__ dynamic_out,
            float *__restrict__ cur_mask_out, float *__restrict__ mask_out, float *__restrict__ dec_static,
            float *__restrict__ dec_dyn, float *__restrict__ reward) {
    constexpr int DIMC = STRAT == STRAT_LBG3D ? 3 : 2;
    typedef Shape<NT, RT, DIMC> SH;
    unsigned *const ems_keys_none = nullptr;       // (the per-warp EMS key list of r01; the MACS scan no longer needs shared memory)
    int lane, warp; const int b = env_index(lane, warp);
    grid_dependency_sync();
    if (b >= c.B) return;
    const int DIM = dim_of<STRAT>(c);
    const int S = SH::S(c);
    const float *srow = env_ptr(static_, b, SH::static_env(c));
    const float *din = env_ptr(dynamic_in, b, SH::dyn_env(c));
    float *dout = env_ptr(dynamic_out, b, SH::dyn_env(c));
    const float *min_ = env_ptr(mask_in, b, (unsigned)S);

    // (1) independent requests first: pointer, the block-id row of `static` (speculative: S floats, so that
    //     `real` needs no second round trip), the masks and the environment state
    const long long p64 = ptr[b];
    const float id0 = lane < S ? srow[lane] : 0.f;
    const float id1 = (S > 32 && lane + 32 < S) ? srow[lane + 32] : 0.f;
    // the edge-length rows too (DIM * S floats, L2-resident across the steps of an episode): the gather of the chosen block
    // (model.py:404-406) then needs no second DRAM round trip behind `ptr`
    float ev0[3] = {0.f, 0.f, 0.f}, ev1[3] = {0.f, 0.f, 0.f};
    if (TAPENV_SPEC_GATHER) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            ev0[r] = (r < DIM && lane < S) ? srow[(1 + r) * S + lane] : 0.f;
            ev1[r] = (r < DIM && S > 32 && lane + 32 < S) ? srow[(1 + r) * S + lane + 32] : 0.f;
        }
    }
    const float m0 = lane < S ? min_[lane] : 0.f;
    const float m1 = (S > 32 && lane + 32 < S) ? min_[lane + 32] : 0.f;
    EnvRegs<STRAT> e; e.load(c, st, b, lane);
    DynPassFast<SH, 2> pass;
    if (FAST) { pass.init(c, lane); pass.load(c, din, lane, 0); }   // first chunk of the precedence tensor in flight

    // (2) the chosen candidate: block id (pack.py:347) and edge lengths (model.py:404-406)
    const bool badp = p64 < 0 || p64 >= S;           // the reference's gather would raise
    const int p = badp ? 0 : (int)p64;
    const bool hi = S > 32 && p >= 32;
    const int real = (int)__shfl_sync(TAPENV_FULL_MASK, hi ? id1 : id0, p & 31);
    float e0, e1, e2, dimv;
    if (TAPENV_SPEC_GATHER) {
        e0 = __shfl_sync(TAPENV_FULL_MASK, hi ? ev1[0] : ev0[0], p & 31);
        e1 = __shfl_sync(TAPENV_FULL_MASK, hi ? ev1[1] : ev0[1], p & 31);
        e2 = __shfl_sync(TAPENV_FULL_MASK, hi ? ev1[2] : ev0[2], p & 31);
        dimv = lane == 0 ? e0 : (lane == 1 ? e1 : e2);
    } else {                                         // dependent gather behind `ptr` (one more DRAM round trip, no extra bytes)
        dimv = lane < DIM ? srow[(1 + lane) * S + p] : 0.f;
        e0 = __shfl_sync(TAPENV_FULL_MASK, dimv, 0);
        e1 = __shfl_sync(TAPENV_FULL_MASK, dimv, 1);
        e2 = __shfl_sync(TAPENV_FULL_MASK, dimv, 2);
    }
    if (dec_static && lane < DIM) dec_static[(size_t)b * (SH::static_rows(c) - 1) + lane] = dimv;

    // (3) environment transition on the register-resident state.  It only needs the pointer, the block's edge
    //     lengths and the (tiny) state, all of which arrive long before the precedence tensor has streamed in.
    //     With a single resident wave (B <= ~4.7k) it runs FIRST so its ~250 warp-instructions overlap the HBM
    //     read phase; with many waves other warps provide that overlap and the shorter register live range wins.
    const int bx = (int)e0;
    const int by = DIM == 3 ? (int)e1 : 1;
    const int bz = (int)(DIM == 3 ? e2 : e1);
    constexpr bool SMALLN = NT > 0 && NT <= 32;
    if (PLACE_FIRST)
        container_add_block<STRAT, SMALLN>(c, st, b, lane, e, bx, by, bz, dec_dyn, ems_keys_none, badp ? 4 : 0, reward);

    // (4) masked copy + column reductions of the precedence tensor
    BandBits bits;
    if (FAST) { pass.finish(c, din, dout, lane, real); bits = pass.combine(c); }
    else bits = dynpass_scalar(c, lane, din, dout, real);

    // (5) masks (pack.py:318-331); block id for the mask is ptr mod n (pack.py:314-316)
    mask_pass<SH>(c, lane, true, m0, m1, SH::mod_n(c, p), bits.blocked(), env_ptr(cur_mask_out, b, (unsigned)S),
                  env_ptr(mask_out, b, (unsigned)S));

    if (!PLACE_FIRST)
        container_add_block<STRAT, SMALLN>(c, st, b, lane, e, bx, by, bz, dec_dyn, ems_keys_none, badp ? 4 : 0, reward);
}

// ------------------------------------------------------------------------------------
// The fused decode step, CTA-per-environment form: CW "copy" warps share the precedence tensor of ONE environment (a row
// pass covers 32*CW/SV rows, so every thread has its whole share in flight at once) while ONE more warp runs the
// placement on the register-resident heightmap at the same time; the column bits meet in shared memory for the masks.
// Chosen when a warp-per-environment grid would leave the machine mostly empty (BASELINE C4: 1 024 environments per GPU =
// 7 warps per SM, 9.6 kB of `dynamic` each, four dependent load chunks per lane and the serial MACS scan behind them:
// r01 10.5 us per launch, 0.31 of the HBM roofline, 10 % warps active).
// ------------------------------------------------------------------------------------
#ifndef TAPENV_SPLIT_CW
#define TAPENV_SPLIT_CW 4
#endif
template <int STRAT, int NT, int RT, int CW>
__global__ void __launch_bounds__(32 * (CW + 1))
step_split_kernel(DevCfg c, StatePtrs st, const int64_t *__restrict__ ptr, const float *__restrict__ static_,
                  const float *__restrict__ dynamic_in, const float *__restrict__ mask_in, float *__restrict__ dynamic_out,
                  float *__restrict__ cur_mask_out, float *__restrict__ mask_out, float *__restrict__ dec_static,
                  float *__restrict__ dec_dyn, float *__restrict__ reward) {
    constexpr int DIM = STRAT == STRAT_LBG3D ? 3 : 2;
    constexpr int TPE = 32 * CW;
    typedef Shape<NT, RT, DIM> SH;
    __shared__ unsigned long long sbits[CW][3];
    unsigned *const ems_keys = nullptr;            // (the EMS key list of r01; the MACS scan no longer needs shared memory)
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    grid_dependency_sync();
    const int S = SH::S(c);
    const float *srow = env_ptr(static_, b, SH::static_env(c));
    const long long p64 = ptr[b];
    // placement warp: the edge-length rows of `static` are requested before `ptr` has arrived (no second round trip)
    float ev0[3] = {0.f, 0.f, 0.f}, ev1[3] = {0.f, 0.f, 0.f};
    if (warp == CW) {
#pragma unroll
        for (int r = 0; r < DIM; ++r) {
            ev0[r] = lane < S ? srow[(1 + r) * S + lane] : 0.f;
            ev1[r] = (S > 32 && lane + 32 < S) ? srow[(1 + r) * S + lane + 32] : 0.f;
        }
    }
    const bool badp = p64 < 0 || p64 >= S;           // the reference's gather would raise
    const int p = badp ? 0 : (int)p64;
    unsigned long long blocked = 0ull;
    float m0 = 0.f, m1 = 0.f;
    if (warp < CW) {
        // ---- copy warps: masked out-of-place copy + column reductions of the precedence tensor (pack.py:333-376, :323-329)
        const float *din = env_ptr(dynamic_in, b, SH::dyn_env(c));
        float *dout = env_ptr(dynamic_out, b, SH::dyn_env(c));
        DynPassFast<SH, 2, TPE> pass;
        pass.init(c, (int)threadIdx.x);
        pass.load(c, din, (int)threadIdx.x, 0);      // in flight before the block id is known
        const int real = (int)srow[p];               // pack.py:347
        pass.finish(c, din, dout, (int)threadIdx.x, real);
        const BandBits bits = pass.combine(c);
        if (lane == 0) { sbits[warp][0] = bits.move; sbits[warp][1] = bits.small; sbits[warp][2] = bits.large; }
    } else {
        // ---- placement warp: gather of the chosen block (model.py:404-406) + Container.add_new_block
        const float *min_ = env_ptr(mask_in, b, (unsigned)S);
        m0 = lane < S ? min_[lane] : 0.f;
        m1 = (S > 32 && lane + 32 < S) ? min_[lane + 32] : 0.f;
        EnvRegs<STRAT> e; e.load(c, st, b, lane);
        const bool hi = S > 32 && p >= 32;
        const float e0 = __shfl_sync(TAPENV_FULL_MASK, hi ? ev1[0] : ev0[0], p & 31);
        const float e1 = __shfl_sync(TAPENV_FULL_MASK, hi ? ev1[1] : ev0[1], p & 31);
        const float e2 = __shfl_sync(TAPENV_FULL_MASK, hi ? ev1[2] : ev0[2], p & 31);
        if (dec_static && lane < DIM) dec_static[(size_t)b * (SH::static_rows(c) - 1) + lane] = lane == 0 ? e0 : (lane == 1 ? e1 : e2);
        const int bx = (int)e0;
        const int by = STRAT == STRAT_LBG3D ? (int)e1 : 1;
        const int bz = (int)(DIM == 3 ? e2 : e1);
        container_add_block<STRAT, (NT > 0 && NT <= 32)>(c, st, b, lane, e, bx, by, bz, dec_dyn, ems_keys, badp ? 4 : 0, reward);
    }
    __syncthreads();
    if (warp == CW) {                                // masks (pack.py:318-331); block id for the mask is ptr mod n (pack.py:314-316)
        BandBits bits; bits.move = bits.small = bits.large = 0ull;
#pragma unroll
        for (int w = 0; w < CW; ++w) { bits.move |= sbits[w][0]; bits.small |= sbits[w][1]; bits.large |= sbits[w][2]; }
        blocked = bits.blocked();
        mask_pass<SH>(c, lane, true, m0, m1, SH::mod_n(c, p), blocked, env_ptr(cur_mask_out, b, (unsigned)S),
                      env_ptr(mask_out, b, (unsigned)S));
    }
}

// ------------------------------------------------------------------------------------
// K6 reward: Container.calc_CPS / calc_ratio (tools.py:3887-3966), one thread per env
// ------------------------------------------------------------------------------------
__global__ void reward_kernel(DevCfg c, StatePtrs st, float *__restrict__ reward) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    grid_dependency_sync();
    if (b >= c.B) return;
    const int4 s = st.scal[b];
    const int cells = c.dim == 2 ? c.W : c.W * c.L;
    int height = 0;
    for (int i = 0; i < cells; ++i) height = max(height, st.heightmap[(size_t)b * cells + i]);
    reward[b] = (float)calc_ratio_dev(c, s.x, s.y, s.z, s.w, height);   // scores[batch_index] = ... (model.py:510), fp32 tensor
}


// ------------------------------------------------------------------------------------
// Two-container inputs ('mul' / 'mul-with', model.py:286-292, :396-447, :503-507): `static` carries one more row, the
// target container id of every candidate (pack.py:212-216); the chosen block goes into container A (id 0) or B (id 1)
// of its environment and the decoder sees both heightmaps, cat(A, B) (model.py:421-447).  Tensor side identical to
// 'bot' (pack.py:300-302, :354-357).
// ------------------------------------------------------------------------------------
template <int STRAT>
__device__ __forceinline__ void mul_place(const DevCfg &c, const StatePtrs &sa, const StatePtrs &sb, int b, int lane, int tgt,
                                          int bx, int by, int bz, float *dec_dyn, unsigned *ems_keys, int extra_flags) {
    const bool valid = tgt == 0 || tgt == 1;          // any other id: the reference appends to neither list and fails later
    const StatePtrs &st = tgt == 1 ? sb : sa;
    if (valid) {
        EnvRegs<STRAT> e; e.load(c, st, b, lane);
        // container_add_block addresses its output as dec_dyn + b*enc_len; here rows are [B][2][enc_len]
        container_add_block<STRAT>(c, st, b, lane, e, bx, by, bz, dec_dyn ? dec_dyn + (size_t)(b + tgt) * c.enc_len : nullptr,
                                   ems_keys, extra_flags);
    } else if (lane == 0) {
        sa.flags[b] |= 4; sb.flags[b] |= 4;
    }
    if (dec_dyn) {                                    // the other container: get_heightmap() (tools.py:3824-3856)
        for (int k = 0; k < 2; ++k) {
            if (valid && k == tgt) continue;          // already written by add_new_block
            const StatePtrs &o = k ? sb : sa;
            encode_state_heightmap(c, o, b, lane, dim_of<STRAT>(c), dec_dyn + ((size_t)b * 2 + k) * c.enc_len);
        }
    }
}

template <int STRAT, bool FAST>
__global__ void __launch_bounds__(32 * kWarpsPerCta)
step_mul_kernel(DevCfg c, StatePtrs sa, StatePtrs sb, const int64_t *__restrict__ ptr, const float *__restrict__ static_,
                const float *__restrict__ dynamic_in, const float *__restrict__ mask_in, float *__restrict__ dynamic_out,
                float *__restrict__ cur_mask_out, float *__restrict__ mask_out, float *__restrict__ dec_static, int dec_rows,
                float *__restrict__ dec_dyn) {
    typedef Shape<0, 0, 0> SH;
    unsigned *const ems_keys_none = nullptr;       // (the per-warp EMS key list of r01; the MACS scan no longer needs shared memory)
    int lane, warp; const int b = env_index(lane, warp);
    grid_dependency_sync();
    if (b >= c.B) return;
    const int DIM = dim_of<STRAT>(c);
    const int S = c.S;
    const float *srow = env_ptr(static_, b, c.static_env);
    const float *din = env_ptr(dynamic_in, b, c.dyn_env);
    float *dout = env_ptr(dynamic_out, b, c.dyn_env);
    const float *min_ = env_ptr(mask_in, b, (unsigned)S);
    const long long p64 = ptr[b];
    const float m0 = lane < S ? min_[lane] : 0.f;
    const float m1 = (S > 32 && lane + 32 < S) ? min_[lane + 32] : 0.f;
    const bool badp = p64 < 0 || p64 >= S;
    const int p = badp ? 0 : (int)p64;
    const int real = (int)srow[p];                                           // pack.py:355
    const int tgt = (int)srow[(c.static_rows - 1) * S + p];                  // target_ids = static[:,-1,:] gathered (model.py:396-401)
    float dimv = 0.f;
    if (lane < dec_rows) dimv = srow[(1 + lane) * S + p];                    // 'mul': static[:,1:-1,:], 'mul-with': static[:,1:,:] (model.py:388-394)
    if (dec_static && lane < dec_rows) dec_static[(size_t)b * dec_rows + lane] = dimv;
    const int bx = (int)__shfl_sync(TAPENV_FULL_MASK, dimv, 0);
    const int by = DIM == 3 ? (int)__shfl_sync(TAPENV_FULL_MASK, dimv, 1) : 1;
    const int bz = (int)__shfl_sync(TAPENV_FULL_MASK, dimv, DIM - 1);
    const BandBits bits = dynpass<SH, FAST>(c, lane, din, dout, real);
    mask_pass<SH>(c, lane, true, m0, m1, SH::mod_n(c, p), bits.blocked(), env_ptr(cur_mask_out, b, (unsigned)S),
                  env_ptr(mask_out, b, (unsigned)S));
    mul_place<STRAT>(c, sa, sb, b, lane, tgt, bx, by, bz, dec_dyn, ems_keys_none, badp ? 4 : 0);
}

template <int STRAT>
__global__ void __launch_bounds__(32 * kWarpsPerCta)
add_blocks_mul_kernel(DevCfg c, StatePtrs sa, StatePtrs sb, const float *__restrict__ blocks, const float *__restrict__ target_ids,
                      float *__restrict__ dec_dyn) {
    unsigned *const ems_keys_none = nullptr;       // (the per-warp EMS key list of r01; the MACS scan no longer needs shared memory)
    int lane, warp; const int b = env_index(lane, warp);
    grid_dependency_sync();
    if (b >= c.B) return;
    const int DIM = dim_of<STRAT>(c);
    const float *blk = blocks + (size_t)b * DIM;
    const int bx = (int)blk[0], by = DIM == 3 ? (int)blk[1] : 1, bz = (int)blk[DIM - 1];
    mul_place<STRAT>(c, sa, sb, b, lane, (int)target_ids[b], bx, by, bz, dec_dyn, ems_keys_none, 0);
}

// scores = (calc_ratio(A) + calc_ratio(B)) / 2 accumulated in an fp32 tensor (model.py:503-507)
__global__ void reward_mul_kernel(DevCfg c, StatePtrs sa, StatePtrs sb, float *__restrict__ reward) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    grid_dependency_sync();
    if (b >= c.B) return;
    const int cells = c.dim == 2 ? c.W : c.W * c.L;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const StatePtrs &st = k ? sb : sa;
        const int4 s = st.scal[b];
        int height = 0;
        for (int i = 0; i < cells; ++i) height = max(height, st.heightmap[(size_t)b * cells + i]);
        acc = __fadd_rn(acc, (float)calc_ratio_dev(c, s.x, s.y, s.z, s.w, height));
    }
    reward[b] = __fdiv_rn(acc, 2.0f);
}

// ------------------------------------------------------------------------------------
// K7 whole-episode kernel: reset + `steps` decode steps + reward in ONE launch, for a known pointer
// sequence.  The precedence tensor is read ONCE and kept as bit rows in shared memory (one 64-bit word per
// band and block row); no intermediate `dynamic` / mask tensor is materialised.
// (tools.calc_positions_lb_greedy :2393, calc_positions_mcs :3213 and pack.reward pack.py:378 are this loop
// on the host, one environment at a time.)
// ------------------------------------------------------------------------------------
template <int STRAT>
__global__ void __launch_bounds__(32 * kWarpsPerCta)
episode_kernel(DevCfg c, StatePtrs st, const float *__restrict__ static_, const float *__restrict__ dynamic,
               const int64_t *__restrict__ ptr_seq, int steps, float *__restrict__ reward,
               float *__restrict__ cur_mask_out, float *__restrict__ mask_out, float *__restrict__ dec_dyn) {
    constexpr bool VOXEL = STRAT == STRAT_LB || STRAT == STRAT_MACS3D;
    constexpr int DIMMAX = (STRAT == STRAT_LBG3D || VOXEL) ? 3 : 2;
    __shared__ unsigned rows[kWarpsPerCta][3][kMaxBlocks][2];          // bit j of (band, row): dynamic[band*n+row][j] != 0
    __shared__ float stat[kWarpsPerCta][1 + DIMMAX][kMaxCandidates];
    unsigned *const ems_keys_none = nullptr;       // (the per-warp EMS key list of r01; the MACS scan no longer needs shared memory)
    int lane, warp; const int b = env_index(lane, warp);
    grid_dependency_sync();
    if (b >= c.B) return;
    const int S = c.S, n = c.n;
    const int DIM = dim_of<STRAT>(c);
    const float *din = env_ptr(dynamic, b, c.dyn_env);
    const float *srow = env_ptr(static_, b, c.static_env);

    // ---- stage the inputs ----
    for (int i = lane; i < 3 * kMaxBlocks * 2; i += 32) (&rows[warp][0][0][0])[i] = 0u;
    for (int i = lane; i < (1 + DIM) * S; i += 32) stat[warp][i / S][i % S] = srow[i];
    __syncwarp();
    if (c.SV * 4 == S && ((uintptr_t)din & 15) == 0) {               // 128-bit sweep, lane = (row-in-pass, column group)
        const int rsub = (int)(((unsigned)lane * c.inv_SV) >> 16), cv = lane - rsub * c.SV;
        const uint4 *src = reinterpret_cast<const uint4 *>(din);
        if (rsub < c.RP) {
            for (int row = rsub; row < c.dyn_rows; row += c.RP) {
                const uint4 v = ldg_stream4(src + row * c.SV + cv);
                const unsigned nib = nz_bits(v.x) | (nz_bits(v.y) << 1) | (nz_bits(v.z) << 2) | (nz_bits(v.w) << 3);
                const int band = row / n, rin = row - band * n;
                if (nib && band < 3) atomicOr(&rows[warp][band][rin][(cv * 4) >> 5], nib << ((cv * 4) & 31));
            }
        }
    } else {
        for (int q = lane; q < c.dyn_rows * S; q += 32) {
            const int row = q / S, col = q - row * S;
            const int band = row / n, rin = row - band * n;
            if (nz_bits(__float_as_uint(__ldg(din + q))) && band < 3) atomicOr(&rows[warp][band][rin][col >> 5], 1u << (col & 31));
        }
    }
    // pointer sequence: lane t (and t+32) keeps the pointer of step t
    long long pq[2] = {0, 0};
#pragma unroll
    for (int s = 0; s < 2; ++s) if (lane + 32 * s < steps) pq[s] = ptr_seq[(size_t)(lane + 32 * s) * c.B + b];
    __syncwarp();

    // ---- reset (Container.__init__, tools.py:3611-3661) ----
    const int cells = DIM == 3 ? c.W * c.L : c.W;
    reset_env_state(c, st, b, lane);
    __syncwarp();
    EnvRegs<STRAT> e;
    e.h = 0; e.sc.valid = e.sc.empty = e.sc.nstable = e.sc.k = 0;
    e.x = 0; e.y = 0;
    if (STRAT == STRAT_LBG3D) { e.x = (int)(((unsigned)lane * c.inv_L) >> 16); e.y = lane - e.x * c.L; }
#pragma unroll
    for (int s = 0; s < 2; ++s) e.hist.x[s] = e.hist.z[s] = e.hist.xx[s] = e.hist.zz[s] = 0;

    unsigned long long mask = S >= 64 ? ~0ull : ((1ull << S) - 1ull);      // model.py:297: ones
    for (int t = 0; t < steps; ++t) {
        const long long p64 = __shfl_sync(TAPENV_FULL_MASK, t < 32 ? pq[0] : pq[1], t & 31);
        const bool badp = p64 < 0 || p64 >= S;
        const int p = badp ? 0 : (int)p64;
        const int real = (int)stat[warp][0][p];                             // pack.py:347
        const int realm = p - (int)(((unsigned)p * c.inv_n) >> 16) * n;     // pack.py:314-316
        const int bx = (int)stat[warp][1][p];
        const int by = DIM == 3 ? (int)stat[warp][2][p] : 1;
        const int bz = (int)stat[warp][DIM][p];
        __syncwarp();
        if (lane < 3 * 2 && real >= 0 && real < n && (lane >> 1) < c.update_time) rows[warp][lane >> 1][real][lane & 1] = 0u;   // pack.py:370-374
        for (int r = 0; r < c.R; ++r) mask &= ~(1ull << (realm + n * r));   // pack.py:318-321
        __syncwarp();
        container_add_block<STRAT>(c, st, b, lane, e, bx, by, bz, t == steps - 1 ? dec_dyn : nullptr,
                                   ems_keys_none, badp ? 4 : 0);
        if (e.sc.k < c.cap) e.sc.k += 1;
        if (STRAT == STRAT_MACS2D) {                                        // refresh the history registers
            __syncwarp();
            e.load(c, st, b, lane);
        }
    }
    // ---- outputs ----
    // accessibility from the remaining bit rows (pack.py:324-329; model.py:297-307 when steps == 0)
    __syncwarp();
    unsigned w[3][2];
#pragma unroll
    for (int bd = 0; bd < 3; ++bd) {
        unsigned lo = 0u, hi = 0u;
        for (int i = lane; i < n; i += 32) { lo |= rows[warp][bd][i][0]; hi |= rows[warp][bd][i][1]; }
        w[bd][0] = warp_or(lo); w[bd][1] = S > 32 ? warp_or(hi) : 0u;
    }
    const unsigned long long mv = ((unsigned long long)w[0][1] << 32) | w[0][0];
    const unsigned long long sm = ((unsigned long long)w[1][1] << 32) | w[1][0];
    const unsigned long long lg = ((unsigned long long)w[2][1] << 32) | w[2][0];
    const unsigned long long cur = mask & ~(mv | (sm & lg));
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int j = lane + 32 * half;
        if (j < S) {
            if (mask_out) mask_out[(size_t)b * S + j] = (float)((mask >> j) & 1ull);
            if (cur_mask_out) cur_mask_out[(size_t)b * S + j] = (float)((cur >> j) & 1ull);
        }
    }
    if (reward) {
        if (VOXEL) {                                 // the state lives in global memory: Container.calc_ratio from there
            __syncwarp();
            const int height = warp_max(encode_state_heightmap(c, st, b, lane, DIM, nullptr));
            if (lane == 0) { const int4 s4 = st.scal[b]; reward[b] = (float)calc_ratio_dev(c, s4.x, s4.y, s4.z, s4.w, height); }
        } else {
            const int height = warp_max(lane < cells ? e.h : 0);
            if (lane == 0) reward[b] = (float)calc_ratio_dev(c, e.sc.valid, e.sc.empty, e.sc.nstable, e.sc.k, height);
        }
    }
}

// deterministic (fixed-order) reduction of (sum r, sum r^2, B) in fp64, single CTA
__global__ void reward_sums_kernel(int B, const float *__restrict__ reward, double *__restrict__ out) {
    __shared__ double s1[1024], s2[1024];
    grid_dependency_sync();
    double a = 0.0, q = 0.0;
    for (int i = threadIdx.x; i < B; i += blockDim.x) { const double r = (double)reward[i]; a += r; q += r * r; }
    s1[threadIdx.x] = a; s2[threadIdx.x] = q;
    __syncthreads();
    for (int w = blockDim.x >> 1; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) { s1[threadIdx.x] += s1[threadIdx.x + w]; s2[threadIdx.x] += s2[threadIdx.x + w]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[0] = s1[0]; out[1] = s2[0]; out[2] = (double)B; }
}

// ------------------------------------------------------------------------------------
// K6 + collective: per-rank sums, then a one-shot exchange of the 24-byte triples over NVLink peer memory.
// Exchange buffer of one rank: slot[kCommDepth][TAPENV_COMM_MAX_RANKS] of six 8-byte words + a call counter.  Every word
// carries 4 bytes of payload and the 4-byte call tag (the "LL" layout NCCL uses for small messages): an 8-byte store is
// atomic, so a word is valid on its own and NO fence is needed on either side -- rank r stores its six words into
// slot[seq % depth][r] of EVERY rank (plain stores to peer-mapped addresses travel over NVLink/NVSwitch) and polls the
// words the peers store into its OWN buffer.  One thread per (peer, word).  A slot is reused after kCommDepth calls; a
// rank can only get that far ahead after every peer has consumed the earlier call (each call waits for all peers).
// (r01 N=2: 7.5 us per episode for the fence-based version of this exchange, 56 us for an NCCL all-gather on a side stream.)
// ------------------------------------------------------------------------------------
constexpr int kCommDepth = 4;
constexpr int kCommWords = 6;
struct CommSlot { unsigned long long w[8]; };           // 6 used; 64-byte stride
// status: sticky, 0 = healthy; bit 0 = a call gave up waiting for a peer (its totals are NaN).  failed_seq: first such call.
struct CommBuf { CommSlot slot[kCommDepth][TAPENV_COMM_MAX_RANKS]; unsigned long long calls, status, failed_seq; unsigned long long pad[5]; };

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// How long a call polls for its peers before it gives up (TAPENV_EXCHANGE_TIMEOUT_MS, default 10 s).  The bound exists so that
// a peer that never calls cannot hang the GPU; it must exceed the worst rank skew of the job (checkpointing or logging on
// one rank, a data-loader stall, first-iteration warm-up).  A timeout is NOT silent: the call writes NaN totals AND sets the
// sticky status word of this rank's exchange buffer (tapenv_comm_status_offset; PeerExchange.check() raises).
static long long exchange_timeout_ns() {
    static const long long ns = [] {
        const char *e = getenv("TAPENV_EXCHANGE_TIMEOUT_MS");
        long long ms = e ? atoll(e) : 10000;
        if (ms < 1) ms = 1;
        return ms * 1000000ll;
    }();
    return ns;
}

__global__ void reward_sums_exchange_kernel(int B, const float *__restrict__ reward, double *__restrict__ out,
                                            double *__restrict__ total, tapenv_peer_comm comm, long long timeout_ns) {
    __shared__ double s1[1024], s2[1024];
    __shared__ unsigned long long seq_sh;
    __shared__ unsigned got[TAPENV_COMM_MAX_RANKS][kCommWords];
    __shared__ int okf;
    grid_dependency_sync();
    double a = 0.0, q = 0.0;
    for (int i = threadIdx.x; i < B; i += blockDim.x) { const double r = (double)reward[i]; a += r; q += r * r; }
    s1[threadIdx.x] = a; s2[threadIdx.x] = q;
    __syncthreads();
    for (int w = blockDim.x >> 1; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) { s1[threadIdx.x] += s1[threadIdx.x + w]; s2[threadIdx.x] += s2[threadIdx.x + w]; }
        __syncthreads();
    }
    CommBuf *mine = reinterpret_cast<CommBuf *>(comm.peer[comm.rank]);
    if (threadIdx.x == 0) {
        if (out) { out[0] = s1[0]; out[1] = s2[0]; out[2] = (double)B; }
        seq_sh = ++mine->calls;
        okf = 1;
    }
    __syncthreads();
    const unsigned long long seq = seq_sh;
    const unsigned tag = (unsigned)seq | 0x80000000u;                      // never 0 (the buffers start zeroed)
    const int t = threadIdx.x;
    if (t < comm.world * kCommWords) {
        const int r = t / kCommWords, wq = t - r * kCommWords;
        const double mv = wq < 2 ? s1[0] : (wq < 4 ? s2[0] : (double)B);
        const unsigned half = (wq & 1) ? (unsigned)__double2hiint(mv) : (unsigned)__double2loint(mv);
        // post word wq of my triple into peer r's buffer
        volatile unsigned long long *dst = reinterpret_cast<CommBuf *>(comm.peer[r])->slot[seq % kCommDepth][comm.rank].w;
        dst[wq] = ((unsigned long long)half << 32) | tag;
        // wait for word wq of peer r's triple in MY buffer
        volatile unsigned long long *src = mine->slot[seq % kCommDepth][r].w;
        unsigned long long v = src[wq];
        if ((unsigned)v != tag) {
            const unsigned long long t0 = global_timer_ns();
            for (;;) {
                for (int spins = 0; (unsigned)v != tag && spins < 256; ++spins) v = src[wq];
                if ((unsigned)v == tag || (long long)(global_timer_ns() - t0) > timeout_ns) break;
            }
        }
        if ((unsigned)v != tag) okf = 0;
        got[r][wq] = (unsigned)(v >> 32);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (!okf) {                                                          // sticky: the host side raises on it
            if (mine->status == 0ull) mine->failed_seq = seq;
            mine->status |= 1ull;
        }
        if (total) {
            double t0 = 0.0, t1 = 0.0, t2 = 0.0;
            for (int k = 0; k < comm.world; ++k) {                           // rank order
                t0 += __hiloint2double((int)got[k][1], (int)got[k][0]);
                t1 += __hiloint2double((int)got[k][3], (int)got[k][2]);
                t2 += __hiloint2double((int)got[k][5], (int)got[k][4]);
            }
            const double bad = nan("");
            total[0] = okf ? t0 : bad; total[1] = okf ? t1 : bad; total[2] = okf ? t2 : bad;
        }
    }
}


// ------------------------------------------------------------------------------------
// Rolling window (generate.InitialContainer, window.cuh) -- alone, or fused behind the placement of the block chosen
// from the PREVIOUS window: ONE launch per rolling decode step.  In rolling.DRL.forward(one_step=True) the outputs of
// update_dynamic / update_mask are locals that are never returned (rolling.py:404-412, :449-453), so the fused rolling
// step performs gather + add_new_block + remove_block + convert_to_input only.
// ------------------------------------------------------------------------------------
__global__ void window_reset_kernel(int B, int T, unsigned *wstate) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    grid_dependency_sync();
    if (q >= B * kWinStateWords) return;
    const int word = q % kWinStateWords;
    const unsigned long long after = T >= 64 ? ~0ull : ((1ull << T) - 1ull);           // generate.py:1672-1673
    unsigned v = 0u;
    if (word == 2) v = (unsigned)after;
    else if (word == 3) v = (unsigned)(after >> 32);
    else if (word >= 4 && word < 12) v = 0xffffffffu;
    wstate[q] = v;
}

#ifndef TAPENV_WINDOW_MIN_BLOCKS
#define TAPENV_WINDOW_MIN_BLOCKS (32 / TAPENV_WARPS_PER_CTA)
#endif
template <int STRAT, bool FAST, int NWc = 0, int RWc = 0>   // STRAT < 0: window only; NWc/RWc > 0: compile-time window shape
__global__ void __launch_bounds__(32 * kWarpsPerCta, TAPENV_WINDOW_MIN_BLOCKS)
window_kernel(WinCfg w, DevCfg c, StatePtrs st, unsigned *__restrict__ wstate, const unsigned long long *__restrict__ pred,
              const int *__restrict__ blocks, const int64_t *__restrict__ ptr, float *__restrict__ dec_static,
              float *__restrict__ dec_dyn, float *__restrict__ static_out, float *__restrict__ dynamic_out,
              float *__restrict__ cur_mask, float *__restrict__ mask_out, int *__restrict__ nodes_out,
              int *__restrict__ remaining_out) {
    constexpr int ES = STRAT < 0 ? STRAT_LBG2D : STRAT;
    __shared__ WinShared shs[kWarpsPerCta];
    __shared__ uint4 lut[16];                         // nibble -> four fp32 0/1 values
    unsigned *const ems_keys_none = nullptr;       // (the per-warp EMS key list of r01; the MACS scan no longer needs shared memory)
    int lane, warp; const int b = env_index(lane, warp);
    if (threadIdx.x < 16) {
        const unsigned t = threadIdx.x, one = 0x3f800000u;
        lut[t] = make_uint4((t & 1u) * one, ((t >> 1) & 1u) * one, ((t >> 2) & 1u) * one, ((t >> 3) & 1u) * one);
    }
    __syncthreads();
    grid_dependency_sync();
    if (b >= w.B) return;
    WinShared &sh = shs[warp];
    unsigned *ws = wstate + (size_t)b * kWinStateWords;
    const unsigned word = lane < kWinStateWords ? ws[lane] : 0u;
    const unsigned long long *pe = pred + (size_t)b * 5 * w.T;
    const unsigned long long pm0 = lane < w.T ? pe[lane] : 0ull;
    const unsigned long long pm1 = lane + 32 < w.T ? pe[lane + 32] : 0ull;
    const long long p64 = ptr ? ptr[b] : -1;
    EnvRegs<ES> e;
    if (STRAT >= 0) e.load(c, st, b, lane);
    const unsigned long long gone = (unsigned long long)__shfl_sync(TAPENV_FULL_MASK, word, 0) |
                                    ((unsigned long long)__shfl_sync(TAPENV_FULL_MASK, word, 1) << 32);
    const unsigned long long after = (unsigned long long)__shfl_sync(TAPENV_FULL_MASK, word, 2) |
                                     ((unsigned long long)__shfl_sync(TAPENV_FULL_MASK, word, 3) << 32);
    const int len = (int)__shfl_sync(TAPENV_FULL_MASK, word, 12);
    int flags = (int)__shfl_sync(TAPENV_FULL_MASK, word, 13);
    sh.list[lane] = (unsigned char)(__shfl_sync(TAPENV_FULL_MASK, word, 4 + (lane >> 2)) >> (8 * (lane & 3)));
    __syncwarp();
    const int *blk = blocks + (size_t)b * w.blocks_env;
    int rm = -1, node = -1, prot = 0;
    bool place = false;
    float dimv = 0.f;
    if (ptr) {
        const bool badp = p64 < 0 || p64 >= w.S;      // the reference's gather / list index would raise
        const int p = badp ? 0 : (int)p64;
        prot = (int)(((unsigned)p * w.inv_n) >> 16);
        const int idx = p - prot * w.n;               // rolling.py:637-638
        if (badp || idx >= len) flags |= 4; else rm = idx;
        if (STRAT >= 0) {
            node = rm >= 0 ? (int)sh.list[idx] : -1;
            place = node >= 0;
            if (lane < w.dim && place) dimv = (float)blk[(node + prot * w.T) * w.dim + lane];     // == static[:,1:,ptr] of the previous window
        }
    }
    // phase A of the window (admission + early loads), then the placement, then decompose / emission: the placement's
    // ~500 warp-instructions run while the window's loads are in flight
    const WinEarly early = window_refill(w, sh, lane, gone, after, len, flags, rm, pe, pm0, pm1, blk);
    if (STRAT >= 0 && ptr) {
        if (dec_static && lane < w.dim) dec_static[(size_t)b * w.dim + lane] = dimv;
        const int bx = (int)__shfl_sync(TAPENV_FULL_MASK, dimv, 0);
        const int by = w.dim == 3 ? (int)__shfl_sync(TAPENV_FULL_MASK, dimv, 1) : 1;
        const int bz = (int)__shfl_sync(TAPENV_FULL_MASK, dimv, w.dim - 1);
        if (place)
            container_add_block<ES>(c, st, b, lane, e, bx, by, bz, dec_dyn, ems_keys_none, 0);
    }
    window_emit<FAST, NWc, RWc>(w, sh, lut, b, lane, early, ws, pe, blk, static_out, dynamic_out, cur_mask, mask_out, nodes_out,
                                remaining_out);
}

// fast path needs 128-bit rows and 16-byte aligned tensors
static bool fast_ok(const DevCfg &d, const void *a, const void *b) {
    const uintptr_t al = (uintptr_t)a | (uintptr_t)b;
    return d.S % 4 == 0 && d.S >= 4 && al % 16 == 0 && d.dyn_rows % d.n == 0;   // whole bands only ('rot-old': n+1 rows)
}

}  // namespace tapenv

// ======================================================================================
// C ABI
// ======================================================================================
using namespace tapenv;

extern "C" {

int tapenv_version(void) { return TAPENV_VERSION; }

const char *tapenv_strerror(int code) {
    switch (code) {
        case TAPENV_OK: return "ok";
        case TAPENV_EINVAL: return "invalid argument (NULL pointer or bad size)";
        case TAPENV_EENUM: return "unknown reward_type / packing_strategy / heightmap_type / input_type";
        case TAPENV_ELIMIT: return "shape exceeds the compiled limits (see tapenv_get_limits)";
        case TAPENV_ESHAPE: return "inconsistent shapes (S != n*R, dyn_rows, static_rows ...)";
        case TAPENV_ECUDA: return "CUDA launch failed";
        case TAPENV_EUNSUPPORTED: return "configuration valid in the reference but not built in this library";
        default: return "unknown error";
    }
}

void tapenv_get_limits(tapenv_limits *out) {
    if (!out) return;
    out->max_width_2d = kMaxWidth2D; out->max_cells_3d = kMaxCells3D;
    out->max_candidates = kMaxCandidates; out->max_blocks = kMaxBlocks;
}

static bool str_ends(const char *s, const char *suf) {
    const size_t a = strlen(s), b = strlen(suf);
    return a >= b && strcmp(s + a - b, suf) == 0;
}
static bool str_in(const char *s, const char *const *set) {
    for (; *set; ++set) if (!strcmp(s, *set)) return true;
    return false;
}

int tapenv_config_init(tapenv_config *cfg, int32_t batch, int32_t blocks_num, int32_t dim, int32_t allow_rot,
                       const int32_t *container_size, const char *reward_type, const char *packing_strategy,
                       const char *heightmap_type, const char *input_type) {
    if (!cfg || !container_size || !reward_type || !packing_strategy || !heightmap_type || !input_type) return TAPENV_EINVAL;
    if (dim != 2 && dim != 3) return TAPENV_EINVAL;
    memset(cfg, 0, sizeof(*cfg));
    cfg->batch = batch; cfg->blocks_num = blocks_num; cfg->dim = dim; cfg->capacity = blocks_num;
    cfg->rotate_types = allow_rot ? (dim == 2 ? 2 : 6) : 1;              // pack.py:306-309 (dim!)
    cfg->width = container_size[0];
    cfg->length = dim == 3 ? container_size[1] : 1;
    cfg->height = container_size[dim - 1];

    // packing strategy, with the reward-type override of tools.py:3617-3620
    if (!strcmp(packing_strategy, "LB_GREEDY")) cfg->strategy = TAPENV_LB_GREEDY;
    else if (!strcmp(packing_strategy, "MACS") || !strcmp(packing_strategy, "MUL")) cfg->strategy = TAPENV_MACS;
    else if (!strcmp(packing_strategy, "LB")) cfg->strategy = TAPENV_LB;
    else return TAPENV_EENUM;
    static const char *const forces_macs[] = {"C+P+S-mul-soft", "C+P+S-mul-hard", "C+P+S-mcs-soft", "C+P+S-mcs-hard", nullptr};
    if (str_in(reward_type, forces_macs)) cfg->strategy = TAPENV_MACS;

    // reward type: the reference's substring tests + the calc_ratio table (tools.py:3923-3964)
    static const char *const sum_types[] = {
        "pyrm-soft-sum", "pyrm-hard-sum", "C+P-mul-soft", "C+P-mul-hard", "C+P-mcs-soft", "C+P-mcs-hard",
        "C+P+S-mul-soft", "C+P+S-mul-hard", "C+P+S-mcs-soft", "C+P+S-mcs-hard", "C+P-lb-hard",
        "C+P+S-lb-soft", "C+P+S-lb-hard", nullptr};
    static const char *const cp_s_types[] = {"pyrm-soft", "pyrm-hard", "mcs-soft", "mcs-hard", nullptr};
    static const char *const sum2_types[] = {"pyrm-soft-SUM", "pyrm-hard-SUM", nullptr};
    if (!strcmp(reward_type, "comp")) cfg->ratio_mode = TAPENV_RATIO_C;
    else if (!strcmp(reward_type, "soft") || !strcmp(reward_type, "hard")) cfg->ratio_mode = TAPENV_RATIO_CS;
    else if (!strcmp(reward_type, "pyrm")) cfg->ratio_mode = TAPENV_RATIO_C_P;
    else if (str_in(reward_type, cp_s_types)) cfg->ratio_mode = TAPENV_RATIO_CP_S;
    else if (str_in(reward_type, sum2_types)) cfg->ratio_mode = TAPENV_RATIO_2C_SUM;
    else if (!strcmp(reward_type, "CPS")) cfg->ratio_mode = TAPENV_RATIO_CPS;
    else if (!strcmp(reward_type, "C+P-lb-soft")) cfg->ratio_mode = TAPENV_RATIO_CP_HALF;
    else if (str_in(reward_type, sum_types)) cfg->ratio_mode = TAPENV_RATIO_SUM;
    else return TAPENV_EENUM;
    cfg->reward_flags = (str_ends(reward_type, "hard") ? TAPENV_RF_HARD : 0) |
                        (strchr(reward_type, 'P') ? TAPENV_RF_P : 0) |
                        (strchr(reward_type, 'S') ? TAPENV_RF_S : 0) |
                        (strstr(reward_type, "mcs") ? TAPENV_RF_MCS_IN : 0) |
                        (!strncmp(reward_type, "mcs", 3) ? TAPENV_RF_MCS_START : 0);

    if (!strcmp(heightmap_type, "full")) cfg->heightmap_type = TAPENV_HM_FULL;
    else if (!strcmp(heightmap_type, "zero")) cfg->heightmap_type = TAPENV_HM_ZERO;
    else if (!strcmp(heightmap_type, "diff")) cfg->heightmap_type = TAPENV_HM_DIFF;
    else return TAPENV_EENUM;

    // input type -> tensor shapes (pack.py:186-223, :338-365)
    const int n = blocks_num;
    if (!strcmp(input_type, "simple") || !strcmp(input_type, "rot")) {
        cfg->static_rows = 1 + dim; cfg->dyn_rows = n; cfg->update_time = 1;
    } else if (!strcmp(input_type, "bot") || !strcmp(input_type, "bot-rot") ||
               !strcmp(input_type, "use-static") || !strcmp(input_type, "use-pnet")) {
        cfg->static_rows = 1 + dim; cfg->dyn_rows = 3 * n; cfg->update_time = 3;
    } else if (!strcmp(input_type, "mul") || !strcmp(input_type, "mul-with")) {
        cfg->static_rows = 2 + dim; cfg->dyn_rows = 3 * n; cfg->update_time = 3;   // + target container id row (pack.py:212-216)
    } else if (!strcmp(input_type, "rot-old")) {
        cfg->static_rows = 1 + dim; cfg->dyn_rows = n + 1; cfg->update_time = 1;   // movement rows + one rotate-state row (pack.py:218-223)
    } else return TAPENV_EENUM;
    const int rc = check_cfg(cfg);
    if (rc != TAPENV_OK) return rc;
    return strategy_kernel(cfg) < 0 ? TAPENV_EUNSUPPORTED : TAPENV_OK;
}

int tapenv_config_check(const tapenv_config *cfg) { return check_cfg(cfg); }

size_t tapenv_state_bytes(const tapenv_config *cfg) {
    if (check_cfg(cfg) != TAPENV_OK) return 0;
    tapenv_state_layout L; layout_of(cfg, &L);
    return L.total;
}

int tapenv_state_get_layout(const tapenv_config *cfg, tapenv_state_layout *out) {
    const int rc = check_cfg(cfg);
    if (rc != TAPENV_OK) return rc;
    if (!out) return TAPENV_EINVAL;
    layout_of(cfg, out);
    return TAPENV_OK;
}

int32_t tapenv_encoded_heightmap_len(const tapenv_config *cfg) {
    if (check_cfg(cfg) != TAPENV_OK) return -1;
    return enc_len_of(cfg);
}

#define TAPENV_PROLOGUE(cfg)                         \
    int rc_ = check_cfg(cfg);                        \
    if (rc_ != TAPENV_OK) return rc_;                \
    const DevCfg d = devcfg_of(cfg);                 \
    const dim3 block(32 * kWarpsPerCta), grid((d.B + kWarpsPerCta - 1) / kWarpsPerCta); \
    cudaStream_t s = (cudaStream_t)stream;

int tapenv_reset(const tapenv_config *cfg, void *state, const float *dynamic, float *cur_mask_out, float *mask_out,
                 void *stream) {
    TAPENV_PROLOGUE(cfg)
    if (d.B == 0) return TAPENV_OK;
    if (!state) return TAPENV_EINVAL;
    if (dynamic && !cur_mask_out) return TAPENV_EINVAL;
    const StatePtrs st = stateptrs_of(cfg, state);
    if (fast_ok(d, dynamic, nullptr)) launch(reset_kernel<true>, grid, block, s, d, st, 1, dynamic, cur_mask_out, mask_out);
    else launch(reset_kernel<false>, grid, block, s, d, st, 1, dynamic, cur_mask_out, mask_out);
    return launch_status();
}

int tapenv_initial_mask(const tapenv_config *cfg, const float *dynamic, float *cur_mask_out, float *mask_out, void *stream) {
    TAPENV_PROLOGUE(cfg)
    if (d.B == 0) return TAPENV_OK;
    if (!dynamic || !cur_mask_out) return TAPENV_EINVAL;
    StatePtrs st; memset(&st, 0, sizeof(st));
    if (fast_ok(d, dynamic, nullptr)) launch(reset_kernel<true>, grid, block, s, d, st, 0, dynamic, cur_mask_out, mask_out);
    else launch(reset_kernel<false>, grid, block, s, d, st, 0, dynamic, cur_mask_out, mask_out);
    return launch_status();
}

int tapenv_update_dynamic(const tapenv_config *cfg, const float *dynamic, const float *static_, const int64_t *ptr,
                          float *dynamic_out, void *stream) {
    TAPENV_PROLOGUE(cfg)
    if (d.B == 0) return TAPENV_OK;
    if (!dynamic || !static_ || !ptr || !dynamic_out) return TAPENV_EINVAL;
    if (fast_ok(d, dynamic, dynamic_out)) launch(update_dynamic_kernel<true>, grid, block, s, d, dynamic, static_, ptr, dynamic_out);
    else launch(update_dynamic_kernel<false>, grid, block, s, d, dynamic, static_, ptr, dynamic_out);
    return launch_status();
}

int tapenv_update_mask(const tapenv_config *cfg, const float *mask, const float *dynamic, const int64_t *ptr,
                       float *new_mask_out, float *chosen_mask_out, void *stream) {
    TAPENV_PROLOGUE(cfg)
    if (d.B == 0) return TAPENV_OK;
    if (!mask || !dynamic || !ptr || !new_mask_out || !chosen_mask_out) return TAPENV_EINVAL;
    if (fast_ok(d, dynamic, nullptr)) launch(update_mask_kernel<true>, grid, block, s, d, mask, dynamic, ptr, new_mask_out, chosen_mask_out);
    else launch(update_mask_kernel<false>, grid, block, s, d, mask, dynamic, ptr, new_mask_out, chosen_mask_out);
    return launch_status();
}

int tapenv_add_blocks(const tapenv_config *cfg, void *state, const float *blocks, float *dec_dynamic_out, void *stream) {
    TAPENV_PROLOGUE(cfg)
    const int strat = strategy_kernel(cfg);
    if (strat < 0) return TAPENV_EUNSUPPORTED;
    if (d.B == 0) return TAPENV_OK;
    if (!state || !blocks) return TAPENV_EINVAL;
    const StatePtrs st = stateptrs_of(cfg, state);
    if (strat == STRAT_MACS3D) launch(add_blocks_kernel<STRAT_MACS3D>, grid, block, s, d, st, blocks, dec_dynamic_out);
    else if (strat == STRAT_LB) launch(add_blocks_kernel<STRAT_LB>, grid, block, s, d, st, blocks, dec_dynamic_out);
    else if (strat == STRAT_LBG2D) launch(add_blocks_kernel<STRAT_LBG2D>, grid, block, s, d, st, blocks, dec_dynamic_out);
    else if (strat == STRAT_LBG3D) launch(add_blocks_kernel<STRAT_LBG3D>, grid, block, s, d, st, blocks, dec_dynamic_out);
    else launch(add_blocks_kernel<STRAT_MACS2D>, grid, block, s, d, st, blocks, dec_dynamic_out);
    return launch_status();
}

static int step_impl(const tapenv_config *cfg, void *state, const int64_t *ptr, const float *static_,
                     const float *dynamic_in, const float *mask_in, float *dynamic_out, float *cur_mask_out,
                     float *mask_out, float *dec_static_out, float *dec_dynamic_out, float *reward_out, void *stream) {
    TAPENV_PROLOGUE(cfg)
    const int strat = strategy_kernel(cfg);
    if (strat < 0 || rot_old_layout(cfg)) return TAPENV_EUNSUPPORTED;
    if (d.B == 0) return TAPENV_OK;
    if (!state || !ptr || !static_ || !dynamic_in || !mask_in || !dynamic_out || !cur_mask_out || !mask_out)
        return TAPENV_EINVAL;
    const StatePtrs st = stateptrs_of(cfg, state);
    const bool fast = fast_ok(d, dynamic_in, dynamic_out);
#define TAPENV_STEP_ARGS d, st, ptr, static_, dynamic_in, mask_in, dynamic_out, cur_mask_out, mask_out, dec_static_out, dec_dynamic_out, reward_out
    // shapes with a fully unrolled instantiation ('bot'-like inputs: 3 bands, all updated); anything else runs generic
    const bool bot = d.dyn_rows == 3 * d.n && d.update_time == 3 && d.static_rows == 1 + d.dim;
    static const int sms = [] { int dev = 0, n = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); return n; }();
    // one resident wave: SMs x (64 regs -> 32 warps) ; PLACE_FIRST only pays off then (profiles/r01_sweep*.json)
    const bool pf = d.B <= sms * 32;
#define TAPENV_STEP_SHAPE(STRAT, N, R)                                                                     \
    do {                                                                                                   \
        if (pf) launch(step_kernel<STRAT, true, N, R, true>, grid, block, s, TAPENV_STEP_ARGS);            \
        else launch(step_kernel<STRAT, true, N, R, false>, grid, block, s, TAPENV_STEP_ARGS);              \
    } while (0)
    // heavy placements (3D, MACS): the 72-register build when it turns a 1.x-wave launch into a single wave
    const int kdef = strat == STRAT_LBG3D ? 5 : 4;          // resident CTAs per SM of the default build (88 / 106 registers)
    // (r02: the capped build for EVERY multi-wave launch was measured too -- no gain at 8 192 .. 65 536 environments,
    // profiles/r02z_lowreg_ab.txt -- so it stays limited to the range where it removes the second wave)
    const bool lowreg = (int)grid.x > sms * kdef && (int)grid.x <= sms * 7;
#define TAPENV_STEP_SHAPE_HEAVY(STRAT, N, R)                                                               \
    do {                                                                                                   \
        if (lowreg) launch(step_kernel<STRAT, true, N, R, true, 7>, grid, block, s, TAPENV_STEP_ARGS);     \
        else TAPENV_STEP_SHAPE(STRAT, N, R);                                                               \
    } while (0)
    // CTA-per-environment form (step_split_kernel) when the warp-per-environment grid would underfill the machine.
    // TAPENV_SPLIT=0/1 forces it off/on (tuning); default: fewer than 12 warps per SM in the warp-per-environment form.
    const char *split_env = getenv("TAPENV_SPLIT");          // read per call: the tests flip it to cover both forms
    const int split_mode = split_env && split_env[0] ? atoi(split_env) : -1;
    const bool split = fast && bot && d.SV <= 32 * TAPENV_SPLIT_CW && strat != STRAT_LB && strat != STRAT_MACS3D &&
                       (split_mode == 1 || (split_mode != 0 && d.B <= sms * 12));
    if (split) {
        const dim3 sgrid(d.B), sblock(32 * (TAPENV_SPLIT_CW + 1));
#define TAPENV_SPLIT_LAUNCH(STRAT, N, R) launch(step_split_kernel<STRAT, N, R, TAPENV_SPLIT_CW>, sgrid, sblock, s, TAPENV_STEP_ARGS)
        if (strat == STRAT_LBG2D) { if (d.n == 10 && d.R == 2) TAPENV_SPLIT_LAUNCH(STRAT_LBG2D, 10, 2); else TAPENV_SPLIT_LAUNCH(STRAT_LBG2D, 0, 0); }
        else if (strat == STRAT_LBG3D) { if (d.n == 10 && d.R == 6) TAPENV_SPLIT_LAUNCH(STRAT_LBG3D, 10, 6); else TAPENV_SPLIT_LAUNCH(STRAT_LBG3D, 0, 0); }
        else { if (d.n == 20 && d.R == 2) TAPENV_SPLIT_LAUNCH(STRAT_MACS2D, 20, 2); else TAPENV_SPLIT_LAUNCH(STRAT_MACS2D, 0, 0); }
        return launch_status();
    }
    if (strat == STRAT_LB) {                          // voxel-state strategies: tensor pass + placement on level masks by the warp, ONE launch
        if (fast) launch(step_kernel<STRAT_LB, true, 0, 0, false>, grid, block, s, TAPENV_STEP_ARGS);
        else launch(step_kernel<STRAT_LB, false, 0, 0, false>, grid, block, s, TAPENV_STEP_ARGS);
    } else if (strat == STRAT_MACS3D) {
        if (fast) launch(step_kernel<STRAT_MACS3D, true, 0, 0, false>, grid, block, s, TAPENV_STEP_ARGS);
        else launch(step_kernel<STRAT_MACS3D, false, 0, 0, false>, grid, block, s, TAPENV_STEP_ARGS);
    } else if (strat == STRAT_LBG2D) {
        if (fast && bot && d.n == 10 && d.R == 2) TAPENV_STEP_SHAPE(STRAT_LBG2D, 10, 2);
        else if (fast && bot && d.n == 20 && d.R == 2) TAPENV_STEP_SHAPE(STRAT_LBG2D, 20, 2);
        else if (fast) TAPENV_STEP_SHAPE(STRAT_LBG2D, 0, 0);
        else launch(step_kernel<STRAT_LBG2D, false, 0, 0, false>, grid, block, s, TAPENV_STEP_ARGS);
    } else if (strat == STRAT_LBG3D) {
        if (fast && bot && d.n == 10 && d.R == 6) TAPENV_STEP_SHAPE_HEAVY(STRAT_LBG3D, 10, 6);
        else if (fast) TAPENV_STEP_SHAPE_HEAVY(STRAT_LBG3D, 0, 0);
        else launch(step_kernel<STRAT_LBG3D, false, 0, 0, false>, grid, block, s, TAPENV_STEP_ARGS);
    } else {
        if (fast && bot && d.n == 20 && d.R == 2) TAPENV_STEP_SHAPE_HEAVY(STRAT_MACS2D, 20, 2);
        else if (fast && bot && d.n == 10 && d.R == 2) TAPENV_STEP_SHAPE_HEAVY(STRAT_MACS2D, 10, 2);
        else if (fast) TAPENV_STEP_SHAPE_HEAVY(STRAT_MACS2D, 0, 0);
        else launch(step_kernel<STRAT_MACS2D, false, 0, 0, false>, grid, block, s, TAPENV_STEP_ARGS);
    }
    return launch_status();
}

int tapenv_step(const tapenv_config *cfg, void *state, const int64_t *ptr, const float *static_,
                const float *dynamic_in, const float *mask_in, float *dynamic_out, float *cur_mask_out,
                float *mask_out, float *dec_static_out, float *dec_dynamic_out, void *stream) {
    return step_impl(cfg, state, ptr, static_, dynamic_in, mask_in, dynamic_out, cur_mask_out, mask_out, dec_static_out,
                     dec_dynamic_out, nullptr, stream);
}

int tapenv_step_reward(const tapenv_config *cfg, void *state, const int64_t *ptr, const float *static_,
                       const float *dynamic_in, const float *mask_in, float *dynamic_out, float *cur_mask_out,
                       float *mask_out, float *dec_static_out, float *dec_dynamic_out, float *reward_out, void *stream) {
    if (!reward_out && cfg && cfg->batch > 0) return TAPENV_EINVAL;
    return step_impl(cfg, state, ptr, static_, dynamic_in, mask_in, dynamic_out, cur_mask_out, mask_out, dec_static_out,
                     dec_dynamic_out, reward_out, stream);
}

int tapenv_reward_sums(const tapenv_config *cfg, const float *reward, double *partial_sums_out, double *total_sums_out,
                       const tapenv_peer_comm *comm, void *stream) {
    int rc_ = check_cfg(cfg);
    if (rc_ != TAPENV_OK) return rc_;
    if ((cfg->batch > 0 && !reward) || (!partial_sums_out && !comm)) return TAPENV_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    if (comm) {
        if (comm->world < 1 || comm->world > TAPENV_COMM_MAX_RANKS || comm->rank < 0 || comm->rank >= comm->world || !total_sums_out)
            return TAPENV_EINVAL;
        for (int r = 0; r < comm->world; ++r) if (!comm->peer[r]) return TAPENV_EINVAL;
        launch(reward_sums_exchange_kernel, 1, 1024, s, (int)cfg->batch, reward, partial_sums_out, total_sums_out, *comm, exchange_timeout_ns());
    } else {
        launch(reward_sums_kernel, 1, 1024, s, (int)cfg->batch, reward, partial_sums_out);
    }
    return launch_status();
}

int tapenv_reward(const tapenv_config *cfg, const void *state, float *reward_out, double *partial_sums_out, void *stream) {
    int rc_ = check_cfg(cfg);
    if (rc_ != TAPENV_OK) return rc_;
    if (cfg->batch > 0 && (!state || !reward_out)) return TAPENV_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    const DevCfg d = devcfg_of(cfg);
    const StatePtrs st = stateptrs_of(cfg, const_cast<void *>(state));
    if (d.B > 0) launch(reward_kernel, (d.B + 127) / 128, 128, s, d, st, reward_out);
    if (partial_sums_out) launch(reward_sums_kernel, 1, 1024, s, d.B, reward_out, partial_sums_out);
    return launch_status();
}

size_t tapenv_comm_bytes(void) { return sizeof(CommBuf); }
size_t tapenv_comm_status_offset(void) { return offsetof(CommBuf, status); }

int tapenv_reward_allreduce(const tapenv_config *cfg, const void *state, float *reward_out, double *partial_sums_out,
                            double *total_sums_out, const tapenv_peer_comm *comm, void *stream) {
    int rc_ = check_cfg(cfg);
    if (rc_ != TAPENV_OK) return rc_;
    if (!comm || comm->world < 1 || comm->world > TAPENV_COMM_MAX_RANKS || comm->rank < 0 || comm->rank >= comm->world)
        return TAPENV_EINVAL;
    for (int r = 0; r < comm->world; ++r) if (!comm->peer[r]) return TAPENV_EINVAL;
    if (!reward_out || !total_sums_out || (cfg->batch > 0 && !state)) return TAPENV_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    const DevCfg d = devcfg_of(cfg);
    const StatePtrs st = stateptrs_of(cfg, const_cast<void *>(state));
    if (d.B > 0) launch(reward_kernel, (d.B + 127) / 128, 128, s, d, st, reward_out);
    launch(reward_sums_exchange_kernel, 1, 1024, s, d.B, reward_out, partial_sums_out, total_sums_out, *comm, exchange_timeout_ns());
    return launch_status();
}

int tapenv_episode(const tapenv_config *cfg, void *state, const float *static_, const float *dynamic,
                   const int64_t *ptr_seq, int32_t steps, float *reward_out, float *cur_mask_out, float *mask_out,
                   float *dec_dynamic_out, void *stream) {
    TAPENV_PROLOGUE(cfg)
    const int strat = strategy_kernel(cfg);
    if (strat < 0 || rot_old_layout(cfg)) return TAPENV_EUNSUPPORTED;
    if (steps < 0 || steps > 64 || steps > cfg->capacity) return TAPENV_ELIMIT;
    if (d.B == 0) return TAPENV_OK;
    if (!state || !static_ || !dynamic || (steps > 0 && !ptr_seq)) return TAPENV_EINVAL;
    const StatePtrs st = stateptrs_of(cfg, state);
    if (strat == STRAT_LB) launch(episode_kernel<STRAT_LB>, grid, block, s, d, st, static_, dynamic, ptr_seq, (int)steps, reward_out, cur_mask_out, mask_out, dec_dynamic_out);
    else if (strat == STRAT_MACS3D) launch(episode_kernel<STRAT_MACS3D>, grid, block, s, d, st, static_, dynamic, ptr_seq, (int)steps, reward_out, cur_mask_out, mask_out, dec_dynamic_out);
    else if (strat == STRAT_LBG2D) launch(episode_kernel<STRAT_LBG2D>, grid, block, s, d, st, static_, dynamic, ptr_seq, (int)steps, reward_out, cur_mask_out, mask_out, dec_dynamic_out);
    else if (strat == STRAT_LBG3D) launch(episode_kernel<STRAT_LBG3D>, grid, block, s, d, st, static_, dynamic, ptr_seq, (int)steps, reward_out, cur_mask_out, mask_out, dec_dynamic_out);
    else launch(episode_kernel<STRAT_MACS2D>, grid, block, s, d, st, static_, dynamic, ptr_seq, (int)steps, reward_out, cur_mask_out, mask_out, dec_dynamic_out);
    return launch_status();
}

int32_t tapenv_packed_words(const tapenv_config *cfg) {
    if (check_cfg(cfg) != TAPENV_OK) return -1;
    const int bitsn = cfg->dyn_rows * cfg->blocks_num * cfg->rotate_types;
    return ((bitsn + 31) / 32 + 3) / 4 * 4;                             // rows of 16-byte multiples
}

int tapenv_reset_packed(const tapenv_config *cfg, void *state, const uint8_t *static_u8, const uint32_t *dynamic_bits,
                        float *static_out, float *dynamic_out, float *cur_mask_out, float *mask_out, void *stream) {
    TAPENV_PROLOGUE(cfg)
    if (d.B == 0) return TAPENV_OK;
    if (!static_u8 || !dynamic_bits || !static_out || !dynamic_out || !cur_mask_out) return TAPENV_EINVAL;
    StatePtrs st; memset(&st, 0, sizeof(st));
    if (state) st = stateptrs_of(cfg, state);
    const int words = tapenv_packed_words(cfg);
    const int clear = state && !voxel_state(cfg) ? 1 : 0;
    if (state && voxel_state(cfg)) {                          // LB keeps voxel grids / x lists: cleared by the plain reset
        launch(reset_kernel<false>, grid, block, s, d, st, 1, (const float *)nullptr, (float *)nullptr, (float *)nullptr);
    }
    if (fast_ok(d, dynamic_out, nullptr) && d.SV * d.RP <= 32)
        launch(unpack_reset_kernel<true>, grid, block, s, d, st, clear, (const unsigned char *)static_u8, (const unsigned *)dynamic_bits, words,
               static_out, dynamic_out, cur_mask_out, mask_out);
    else
        launch(unpack_reset_kernel<false>, grid, block, s, d, st, clear, (const unsigned char *)static_u8, (const unsigned *)dynamic_bits, words,
               static_out, dynamic_out, cur_mask_out, mask_out);
    return launch_status();
}

// ---- two-container inputs ('mul' / 'mul-with') --------------------------------------------------------------------
int tapenv_step_mul(const tapenv_config *cfg, void *state_a, void *state_b, const int64_t *ptr, const float *static_,
                    const float *dynamic_in, const float *mask_in, float *dynamic_out, float *cur_mask_out,
                    float *mask_out, float *dec_static_out, int32_t dec_static_rows, float *dec_dynamic_out, void *stream) {
    TAPENV_PROLOGUE(cfg)
    const int strat = strategy_kernel(cfg);
    if (strat < 0) return TAPENV_EUNSUPPORTED;
    if (cfg->static_rows != 2 + cfg->dim) return TAPENV_ESHAPE;
    if (dec_static_rows != cfg->dim && dec_static_rows != cfg->dim + 1) return TAPENV_ESHAPE;
    if (d.B == 0) return TAPENV_OK;
    if (!state_a || !state_b || state_a == state_b || !ptr || !static_ || !dynamic_in || !mask_in || !dynamic_out || !cur_mask_out || !mask_out)
        return TAPENV_EINVAL;
    const StatePtrs sa = stateptrs_of(cfg, state_a), sb = stateptrs_of(cfg, state_b);
    const bool fast = fast_ok(d, dynamic_in, dynamic_out);
#define TAPENV_MUL(STRAT)                                                                                               \
    do {                                                                                                                \
        if (fast) launch(step_mul_kernel<STRAT, true>, grid, block, s, d, sa, sb, ptr, static_, dynamic_in, mask_in,     \
                         dynamic_out, cur_mask_out, mask_out, dec_static_out, (int)dec_static_rows, dec_dynamic_out);   \
        else launch(step_mul_kernel<STRAT, false>, grid, block, s, d, sa, sb, ptr, static_, dynamic_in, mask_in,         \
                    dynamic_out, cur_mask_out, mask_out, dec_static_out, (int)dec_static_rows, dec_dynamic_out);        \
    } while (0)
    if (strat == STRAT_LBG2D) TAPENV_MUL(STRAT_LBG2D);
    else if (strat == STRAT_LBG3D) TAPENV_MUL(STRAT_LBG3D);
    else if (strat == STRAT_LB) TAPENV_MUL(STRAT_LB);
    else if (strat == STRAT_MACS3D) TAPENV_MUL(STRAT_MACS3D);
    else TAPENV_MUL(STRAT_MACS2D);
    return launch_status();
}

int tapenv_add_blocks_mul(const tapenv_config *cfg, void *state_a, void *state_b, const float *blocks,
                          const float *target_ids, float *dec_dynamic_out, void *stream) {
    TAPENV_PROLOGUE(cfg)
    const int strat = strategy_kernel(cfg);
    if (strat < 0) return TAPENV_EUNSUPPORTED;
    if (d.B == 0) return TAPENV_OK;
    if (!state_a || !state_b || state_a == state_b || !blocks || !target_ids) return TAPENV_EINVAL;
    const StatePtrs sa = stateptrs_of(cfg, state_a), sb = stateptrs_of(cfg, state_b);
    if (strat == STRAT_LB) launch(add_blocks_mul_kernel<STRAT_LB>, grid, block, s, d, sa, sb, blocks, target_ids, dec_dynamic_out);
    else if (strat == STRAT_MACS3D) launch(add_blocks_mul_kernel<STRAT_MACS3D>, grid, block, s, d, sa, sb, blocks, target_ids, dec_dynamic_out);
    else if (strat == STRAT_LBG2D) launch(add_blocks_mul_kernel<STRAT_LBG2D>, grid, block, s, d, sa, sb, blocks, target_ids, dec_dynamic_out);
    else if (strat == STRAT_LBG3D) launch(add_blocks_mul_kernel<STRAT_LBG3D>, grid, block, s, d, sa, sb, blocks, target_ids, dec_dynamic_out);
    else launch(add_blocks_mul_kernel<STRAT_MACS2D>, grid, block, s, d, sa, sb, blocks, target_ids, dec_dynamic_out);
    return launch_status();
}

int tapenv_reward_mul(const tapenv_config *cfg, const void *state_a, const void *state_b, float *reward_out, void *stream) {
    int rc_ = check_cfg(cfg);
    if (rc_ != TAPENV_OK) return rc_;
    if (cfg->batch == 0) return TAPENV_OK;
    if (!state_a || !state_b || !reward_out) return TAPENV_EINVAL;
    const DevCfg d = devcfg_of(cfg);
    const StatePtrs sa = stateptrs_of(cfg, const_cast<void *>(state_a)), sb = stateptrs_of(cfg, const_cast<void *>(state_b));
    launch(reward_mul_kernel, (d.B + 127) / 128, 128, (cudaStream_t)stream, d, sa, sb, reward_out);
    return launch_status();
}

// ---- rolling window ---------------------------------------------------------------------------------------------
static int check_wcfg(const tapenv_window_config *w) {
    if (!w) return TAPENV_EINVAL;
    if (w->batch < 0 || w->total_blocks < 1 || w->window < 1 || w->window > w->total_blocks) return TAPENV_EINVAL;
    if (w->dim != 2 && w->dim != 3) return TAPENV_EINVAL;
    if (w->rotate_types != (w->dim == 2 ? 2 : 6)) return TAPENV_ESHAPE;      // InitialContainer: factorial(block_dim), generate.py:1610
    if (w->node_order != TAPENV_WINDOW_ORDER_REFERENCE && w->node_order != TAPENV_WINDOW_ORDER_SORTED) return TAPENV_EENUM;
    if (w->total_blocks > kWinMaxTotal || w->window > kWinMaxWindow) return TAPENV_ELIMIT;
    if (w->window * w->rotate_types > kMaxCandidates) return TAPENV_ELIMIT;
    return TAPENV_OK;
}

static WinCfg wincfg_of(const tapenv_window_config *w) {
    WinCfg d;
    d.B = w->batch; d.T = w->total_blocks; d.n = w->window; d.dim = w->dim; d.R = w->rotate_types; d.S = d.n * d.R;
    d.setorder = (w->node_order == TAPENV_WINDOW_ORDER_REFERENCE && 2 * d.n < d.T) ? 1 : 0;   // FilterAtlas.__iter__ (window.cuh)
    d.SV = d.S % 4 == 0 ? d.S / 4 : 1;
    d.RP = 32 / d.SV > 0 ? 32 / d.SV : 1;
    d.PB = (d.n + d.RP - 1) / d.RP;
    d.inv_n = (65536u + d.n - 1) / d.n;
    d.inv_SV = (65536u + d.SV - 1) / d.SV;
    // last axis of itertools.permutations(range(dim))[r]: 2D (0,1),(1,0); 3D (0,1,2),(0,2,1),(1,0,2),(1,2,0),(2,0,1),(2,1,0)
    // code 0 -> (left,right), 1 -> (forward,backward), 2 -> zeros; in 2D a last axis of 1 is the vertical one (generate.py:1802-1806)
    d.lastcodes = d.dim == 2 ? (2u | (0u << 2)) : (2u | (1u << 2) | (2u << 4) | (0u << 6) | (1u << 8) | (0u << 10));
    d.blocks_env = (unsigned)(d.R * d.T * d.dim);
    d.rotblocks = w->blocks_are_rotations ? 1 : 0;
    // perm_r[d] of itertools.permutations(range(dim)), 2 bits per (r, d), 6 bits per rotation
    static const int P2[2][3] = {{0, 1, 0}, {1, 0, 0}};
    static const int P3[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
    d.permcodes = 0ull;
    for (int r = 0; r < d.R; ++r)
        for (int k = 0; k < 3; ++k)
            d.permcodes |= (unsigned long long)(d.dim == 2 ? P2[r][k] : P3[r][k]) << (6 * r + 2 * k);
    d.mul_all = d.mul_c0 = d.mul_c1 = 0ull;
    for (int r = 0; r < d.R; ++r) {
        const unsigned code = (d.lastcodes >> (2 * r)) & 3u;
        const unsigned long long bit = 1ull << (r * d.n);
        d.mul_all |= bit;
        if (code == 0u) d.mul_c0 |= bit;
        if (code == 1u) d.mul_c1 |= bit;
    }
    return d;
}

size_t tapenv_window_state_bytes(const tapenv_window_config *wcfg) {
    if (check_wcfg(wcfg) != TAPENV_OK) return 0;
    return (size_t)wcfg->batch * kWinStateWords * sizeof(unsigned);
}

int tapenv_window_reset(const tapenv_window_config *wcfg, void *wstate, void *stream) {
    const int rc = check_wcfg(wcfg);
    if (rc != TAPENV_OK) return rc;
    if (wcfg->batch == 0) return TAPENV_OK;
    if (!wstate) return TAPENV_EINVAL;
    const int total = wcfg->batch * kWinStateWords;
    launch(window_reset_kernel, (total + 255) / 256, 256, (cudaStream_t)stream, (int)wcfg->batch, (int)wcfg->total_blocks, (unsigned *)wstate);
    return launch_status();
}

static bool win_fast_ok(const WinCfg &w, const void *a, const void *b, const void *c2) {
    const uintptr_t al = (uintptr_t)a | (uintptr_t)b | (uintptr_t)c2;
    return w.S % 4 == 0 && al % 16 == 0;
}

int tapenv_window_next(const tapenv_window_config *wcfg, void *wstate, const uint64_t *pred, const int32_t *blocks,
                       const int64_t *prev_ptr, float *static_out, float *dynamic_out, float *cur_mask_out,
                       float *mask_out, int32_t *nodes_out, int32_t *remaining_out, void *stream) {
    const int rc = check_wcfg(wcfg);
    if (rc != TAPENV_OK) return rc;
    if (wcfg->batch == 0) return TAPENV_OK;
    if (!wstate || !pred || !blocks || !static_out || !dynamic_out) return TAPENV_EINVAL;
    const WinCfg w = wincfg_of(wcfg);
    DevCfg c; memset(&c, 0, sizeof(c));
    StatePtrs st; memset(&st, 0, sizeof(st));
    const dim3 block(32 * kWarpsPerCta), grid((w.B + kWarpsPerCta - 1) / kWarpsPerCta);
    cudaStream_t s = (cudaStream_t)stream;
#define TAPENV_WIN_ARGS w, c, st, (unsigned *)wstate, (const unsigned long long *)pred, (const int *)blocks
    if (win_fast_ok(w, dynamic_out, cur_mask_out, mask_out))
        launch(window_kernel<-1, true>, grid, block, s, TAPENV_WIN_ARGS, prev_ptr, (float *)nullptr, (float *)nullptr, static_out,
               dynamic_out, cur_mask_out, mask_out, (int *)nodes_out, (int *)remaining_out);
    else
        launch(window_kernel<-1, false>, grid, block, s, TAPENV_WIN_ARGS, prev_ptr, (float *)nullptr, (float *)nullptr, static_out,
               dynamic_out, cur_mask_out, mask_out, (int *)nodes_out, (int *)remaining_out);
    return launch_status();
}

int tapenv_rolling_step(const tapenv_config *cfg, void *state, const tapenv_window_config *wcfg, void *wstate,
                        const uint64_t *pred, const int32_t *blocks, const int64_t *ptr, float *dec_static_out,
                        float *dec_dynamic_out, float *static_out, float *dynamic_out, float *cur_mask_out,
                        float *mask_out, int32_t *nodes_out, int32_t *remaining_out, void *stream) {
    TAPENV_PROLOGUE(cfg)
    const int rcw = check_wcfg(wcfg);
    if (rcw != TAPENV_OK) return rcw;
    if (cfg->batch != wcfg->batch || cfg->dim != wcfg->dim || cfg->blocks_num != wcfg->window ||
        cfg->rotate_types != wcfg->rotate_types || cfg->static_rows != 1 + cfg->dim) return TAPENV_ESHAPE;
    const int strat = strategy_kernel(cfg);
    if (strat < 0 || rot_old_layout(cfg)) return TAPENV_EUNSUPPORTED;
    if (d.B == 0) return TAPENV_OK;
    if (!state || !wstate || !pred || !blocks || !ptr || !static_out || !dynamic_out) return TAPENV_EINVAL;
    const WinCfg w = wincfg_of(wcfg);
    const DevCfg c = d;
    const StatePtrs st = stateptrs_of(cfg, state);
    const bool fast = win_fast_ok(w, dynamic_out, cur_mask_out, mask_out);
#define TAPENV_ROLL_ARGS TAPENV_WIN_ARGS, ptr, dec_static_out, dec_dynamic_out, static_out, dynamic_out, cur_mask_out, mask_out, (int *)nodes_out, (int *)remaining_out
#define TAPENV_ROLL(STRAT)                                                                                              \
    do {                                                                                                                \
        if (fast && w.n == 10 && w.R == 6) launch(window_kernel<STRAT, true, 10, 6>, grid, block, s, TAPENV_ROLL_ARGS);   \
        else if (fast && w.n == 10 && w.R == 2) launch(window_kernel<STRAT, true, 10, 2>, grid, block, s, TAPENV_ROLL_ARGS); \
        else if (fast) launch(window_kernel<STRAT, true>, grid, block, s, TAPENV_ROLL_ARGS);                            \
        else launch(window_kernel<STRAT, false>, grid, block, s, TAPENV_ROLL_ARGS);                                     \
    } while (0)
    if (strat == STRAT_LBG2D) TAPENV_ROLL(STRAT_LBG2D);
    else if (strat == STRAT_LBG3D) TAPENV_ROLL(STRAT_LBG3D);
    else if (strat == STRAT_LB) TAPENV_ROLL(STRAT_LB);
    else if (strat == STRAT_MACS3D) TAPENV_ROLL(STRAT_MACS3D);
    else TAPENV_ROLL(STRAT_MACS2D);
    return launch_status();
}

}  // extern "C"
