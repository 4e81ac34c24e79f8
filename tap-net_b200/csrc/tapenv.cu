// tapenv.cu -- kernels and C ABI of the B200-native TAP packing-environment step.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -shared -Xcompiler -fPIC
// No CPU fallback exists: every entry point launches a CUDA kernel or returns an error.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <math.h>

#include "../../include/tapenv.h"
#include "tapenv_common.cuh"
#include "dynmask.cuh"
#include "lbg2d.cuh"

namespace tapenv {

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
    const size_t B = (size_t)c->batch, n = (size_t)c->blocks_num, dim = (size_t)c->dim;
    size_t off = 0;
    L->scalars = off;   off = align_up(off + B * 4 * sizeof(int32_t), 256);
    L->heightmap = off; off = align_up(off + B * (size_t)cells_of(c) * sizeof(int32_t), 256);
    L->positions = off; off = align_up(off + B * n * dim * sizeof(int32_t), 256);
    L->blocks = off;    off = align_up(off + B * n * dim * sizeof(int32_t), 256);
    L->stable = off;    off = align_up(off + B * n, 256);
    L->flags = off;     off = align_up(off + B * sizeof(int32_t), 256);
    L->total = off;
}

static DevCfg devcfg_of(const tapenv_config *c) {
    DevCfg d;
    d.B = c->batch; d.n = c->blocks_num; d.dim = c->dim; d.R = c->rotate_types;
    d.W = c->width; d.L = c->length; d.H = c->height; d.S = c->blocks_num * c->rotate_types;
    d.strategy = c->strategy; d.hm_type = c->heightmap_type; d.flags = c->reward_flags; d.ratio_mode = c->ratio_mode;
    d.static_rows = c->static_rows; d.dyn_rows = c->dyn_rows; d.update_time = c->update_time;
    d.enc_len = enc_len_of(c);
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
    return s;
}

static int check_cfg(const tapenv_config *c) {
    if (!c) return TAPENV_EINVAL;
    if (c->batch < 0 || c->blocks_num < 1 || c->width < 1 || c->height < 1) return TAPENV_EINVAL;
    if (c->dim != 2 && c->dim != 3) return TAPENV_EINVAL;
    if (c->dim == 2 && c->length != 1) return TAPENV_ESHAPE;
    if (c->rotate_types < 1) return TAPENV_EINVAL;
    if (c->strategy != TAPENV_LB_GREEDY && c->strategy != TAPENV_MACS) return TAPENV_EENUM;
    if (c->heightmap_type < 0 || c->heightmap_type > 2) return TAPENV_EENUM;
    if (c->ratio_mode < 0 || c->ratio_mode > TAPENV_RATIO_CP_HALF) return TAPENV_EENUM;
    if (c->static_rows < 1 + c->dim) return TAPENV_ESHAPE;
    if (c->dyn_rows != c->blocks_num && c->dyn_rows != 3 * c->blocks_num) return TAPENV_ESHAPE;
    if (c->update_time != 1 && c->update_time != 3) return TAPENV_ESHAPE;
    if (c->update_time * c->blocks_num > c->dyn_rows) return TAPENV_ESHAPE;
    if (c->dim == 2 && c->width > kMaxWidth2D) return TAPENV_ELIMIT;
    if (c->dim == 3 && c->width * c->length > kMaxCells3D) return TAPENV_ELIMIT;
    if (c->blocks_num * c->rotate_types > kMaxCandidates) return TAPENV_ELIMIT;
    if (c->blocks_num > kMaxBlocks) return TAPENV_ELIMIT;
    if ((long long)c->dyn_rows * c->blocks_num * c->rotate_types >= 65536) return TAPENV_ELIMIT;
    return TAPENV_OK;
}

static int check_strategy_built(const tapenv_config *c) {
    if (c->strategy == TAPENV_LB_GREEDY && c->dim == 2) return TAPENV_OK;
    return TAPENV_EUNSUPPORTED;
}

static int g_envs_per_cta = 0;   // tuning knob only (0 = auto); never changes results

static int pick_epc(int B) {
    if (g_envs_per_cta > 0) return g_envs_per_cta;
    return B > 32768 ? 4 : 1;    // one CTA per environment unless the grid would be many waves deep
}

static int launch_status() { return cudaGetLastError() == cudaSuccess ? TAPENV_OK : TAPENV_ECUDA; }

// ------------------------------------------------------------------------------------
// device: per-environment indexing
// ------------------------------------------------------------------------------------
__device__ __forceinline__ int env_index(int &lane) {
    lane = threadIdx.x & 31;
    return blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
}

__device__ __forceinline__ Scal load_scal(const StatePtrs &st, int b) {
    const int4 v = st.scal[b];
    Scal s; s.valid = v.x; s.empty = v.y; s.nstable = v.z; s.k = v.w;
    return s;
}

// Container.add_new_block for one environment (tools.py:3663-3744): placement,
// commit, k += 1 even when the placement failed (tools.py:3713), heightmap encoding.
__device__ __forceinline__ void container_add_block_2d(const DevCfg &c, const StatePtrs &st, int b, int lane,
                                                       int bx, int bz, float *dec_dyn) {
    int h = lane < c.W ? st.heightmap[(size_t)b * c.W + lane] : 0;
    Scal sc = load_scal(st, b);
    if (sc.k >= c.n) {   // the reference raises IndexError (rotate_state[k], tools.py:3677)
        if (lane == 0) st.flags[b] |= 2;
    } else {
        const PlaceResult r = lbg2d_place(c, lane, bx, bz, h, sc);
        if (lane < c.W) st.heightmap[(size_t)b * c.W + lane] = h;
        if (lane == 0) {
            const size_t o = ((size_t)b * c.n + sc.k) * 2;
            st.blocks[o] = bx; st.blocks[o + 1] = bz;
            if (r.placed) { st.positions[o] = r.x; st.positions[o + 1] = r.z; }
            st.stable[(size_t)b * c.n + sc.k] = (unsigned char)r.stable;
            if (r.placed && r.z + bz > c.H) st.flags[b] |= 1;
            st.scal[b] = make_int4(sc.valid, sc.empty, sc.nstable, sc.k + 1);
        }
    }
    if (dec_dyn) encode_heightmap_2d(c, lane, h, dec_dyn + (size_t)b * c.enc_len);
}

// ------------------------------------------------------------------------------------
// K0 reset: Container.__init__ / clear_container + initial accessibility mask
// ------------------------------------------------------------------------------------
template <int VW, int CH>
__global__ void reset_kernel(DevCfg c, StatePtrs st, const float *__restrict__ dynamic,
                             float *__restrict__ cur_mask, float *__restrict__ mask, unsigned sv_magic) {
    int lane; const int b = env_index(lane);
    if (b >= c.B) return;
    const int cells = c.dim == 2 ? c.W : c.W * c.L;
    for (int i = lane; i < cells; i += 32) st.heightmap[(size_t)b * cells + i] = 0;
    for (int i = lane; i < c.n * c.dim; i += 32) { st.positions[(size_t)b * c.n * c.dim + i] = 0; st.blocks[(size_t)b * c.n * c.dim + i] = 0; }
    for (int i = lane; i < c.n; i += 32) st.stable[(size_t)b * c.n + i] = 0;
    if (lane == 0) { st.scal[b] = make_int4(0, 0, 0, 0); st.flags[b] = 0; }
    if (dynamic == nullptr) return;
    const int SV = c.S / VW, total = c.dyn_rows * SV;
    const float *din = dynamic + (size_t)b * c.dyn_rows * c.S;
    BandBits bits; bits.clear();
    DynTile<VW, CH> tile;
    for (int base = 0; base < total; base += 32 * CH) {
        tile.load(din, base, lane, total);
        tile.process(c, sv_magic, SV, nullptr, base, lane, total, -1, bits);
    }
    bits.combine(c.S);
    mask_pass(c, lane, nullptr, -1, bits.blocked(), cur_mask + (size_t)b * c.S, mask ? mask + (size_t)b * c.S : nullptr);
}

// ------------------------------------------------------------------------------------
// unfused pieces (signature parity with pack.update_dynamic / pack.update_mask / add_new_block)
// ------------------------------------------------------------------------------------
template <int VW, int CH>
__global__ void update_dynamic_kernel(DevCfg c, const float *__restrict__ dynamic, const float *__restrict__ static_,
                                      const int64_t *__restrict__ ptr, float *__restrict__ out, unsigned sv_magic) {
    int lane; const int b = env_index(lane);
    if (b >= c.B) return;
    const int SV = c.S / VW, total = c.dyn_rows * SV;
    const float *din = dynamic + (size_t)b * c.dyn_rows * c.S;
    float *dout = out + (size_t)b * c.dyn_rows * c.S;
    DynTile<VW, CH> tile;
    tile.load(din, 0, lane, total);
    const long long p = ptr[b];
    const int real = (int)static_[(size_t)b * c.static_rows * c.S + p];    // pack.py:347 (.long() truncates)
    BandBits bits; bits.clear();
    tile.process(c, sv_magic, SV, dout, 0, lane, total, real, bits);
    for (int base = 32 * CH; base < total; base += 32 * CH) {
        tile.load(din, base, lane, total);
        tile.process(c, sv_magic, SV, dout, base, lane, total, real, bits);
    }
}

template <int VW, int CH>
__global__ void update_mask_kernel(DevCfg c, const float *__restrict__ mask, const float *__restrict__ dynamic,
                                   const int64_t *__restrict__ ptr, float *__restrict__ new_mask,
                                   float *__restrict__ chosen_mask, unsigned sv_magic) {
    int lane; const int b = env_index(lane);
    if (b >= c.B) return;
    const int SV = c.S / VW, total = c.dyn_rows * SV;
    const float *din = dynamic + (size_t)b * c.dyn_rows * c.S;
    BandBits bits; bits.clear();
    DynTile<VW, CH> tile;
    for (int base = 0; base < total; base += 32 * CH) {
        tile.load(din, base, lane, total);
        tile.process(c, sv_magic, SV, nullptr, base, lane, total, -1, bits);
    }
    bits.combine(c.S);
    const int realm = (int)(ptr[b] % c.n);                                 // pack.py:314-316
    mask_pass(c, lane, mask + (size_t)b * c.S, realm, bits.blocked(), new_mask + (size_t)b * c.S,
              chosen_mask + (size_t)b * c.S);
}

__global__ void add_blocks_lbg2d_kernel(DevCfg c, StatePtrs st, const float *__restrict__ blocks,
                                        float *__restrict__ dec_dyn) {
    int lane; const int b = env_index(lane);
    if (b >= c.B) return;
    const int bx = (int)blocks[(size_t)b * 2], bz = (int)blocks[(size_t)b * 2 + 1];   // .astype(int) tools.py:3689
    container_add_block_2d(c, st, b, lane, bx, bz, dec_dyn);
}

// ------------------------------------------------------------------------------------
// fused decode-step kernel, LB_GREEDY 2D  (K1 + K2)
// ------------------------------------------------------------------------------------
template <int VW, int CH>
__global__ void step_lbg2d_kernel(DevCfg c, StatePtrs st, const int64_t *__restrict__ ptr,
                                  const float *__restrict__ static_, const float *__restrict__ dynamic_in,
                                  const float *__restrict__ mask_in, float *__restrict__ dynamic_out,
                                  float *__restrict__ cur_mask_out, float *__restrict__ mask_out,
                                  float *__restrict__ dec_static, float *__restrict__ dec_dyn, unsigned sv_magic) {
    int lane; const int b = env_index(lane);
    if (b >= c.B) return;
    const int SV = c.S / VW, total = c.dyn_rows * SV;
    const float *din = dynamic_in + (size_t)b * c.dyn_rows * c.S;
    float *dout = dynamic_out + (size_t)b * c.dyn_rows * c.S;

    // (1) put the precedence tensor in flight first; everything below up to (3) overlaps with it
    DynTile<VW, CH> tile;
    tile.load(din, 0, lane, total);

    // (2) environment transition on the tiny state
    const long long p = ptr[b];
    const float *srow = static_ + (size_t)b * c.static_rows * c.S + p;
    const int real = (int)srow[0];                                         // pack.py:347
    const float fx = srow[(size_t)c.S], fz = srow[(size_t)2 * c.S];        // model.py:404-406
    if (dec_static && lane < c.static_rows - 1) dec_static[(size_t)b * (c.static_rows - 1) + lane] = srow[(size_t)(1 + lane) * c.S];
    const float mval0 = lane < c.S ? mask_in[(size_t)b * c.S + lane] : 0.f;   // early issue; re-read is avoided for S<=32
    container_add_block_2d(c, st, b, lane, (int)fx, (int)fz, dec_dyn);

    // (3) masked copy + column reductions
    BandBits bits; bits.clear();
    tile.process(c, sv_magic, SV, dout, 0, lane, total, real, bits);
    for (int base = 32 * CH; base < total; base += 32 * CH) {
        tile.load(din, base, lane, total);
        tile.process(c, sv_magic, SV, dout, base, lane, total, real, bits);
    }
    bits.combine(c.S);

    // (4) masks (pack.py:318-331)
    const int realm = (int)(p % c.n);
    const unsigned long long blocked = bits.blocked();
    if (lane < c.S) {
        float m = mval0;
        if ((lane % c.n) == realm) m = 0.f;
        mask_out[(size_t)b * c.S + lane] = m;
        cur_mask_out[(size_t)b * c.S + lane] = ((blocked >> lane) & 1ull) ? 0.f : m;
    }
    for (int j = lane + 32; j < c.S; j += 32) {
        float m = mask_in[(size_t)b * c.S + j];
        if ((j % c.n) == realm) m = 0.f;
        mask_out[(size_t)b * c.S + j] = m;
        cur_mask_out[(size_t)b * c.S + j] = ((blocked >> j) & 1ull) ? 0.f : m;
    }
}

// ------------------------------------------------------------------------------------
// K6 reward: Container.calc_CPS / calc_ratio (tools.py:3887-3966), one thread per env
// ------------------------------------------------------------------------------------
__global__ void reward_kernel(DevCfg c, StatePtrs st, float *__restrict__ reward) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= c.B) return;
    const int4 s = st.scal[b];
    const int cells = c.dim == 2 ? c.W : c.W * c.L;
    double ratio;
    if (s.w == 0) {                                // current_blocks_num == 0 -> C=P=S=0 (tools.py:3888-3889)
        ratio = 0.0;
        if (c.ratio_mode == TAPENV_RATIO_CP_HALF) ratio = 0.0 / 2; else ratio = 0.0 / 3;
    } else {
        int height = 0;
        for (int i = 0; i < cells; ++i) height = max(height, st.heightmap[(size_t)b * cells + i]);
        const double C = (double)s.x / (double)((long long)cells * height);
        const double P = (double)s.x / (double)(s.y + s.x);
        const double S = (double)s.z / (double)s.w;
        switch (c.ratio_mode) {
            case TAPENV_RATIO_C: ratio = C / 3; break;
            case TAPENV_RATIO_CS: ratio = (C * S) / 3; break;
            case TAPENV_RATIO_C_P: ratio = (C + P) / 3; break;
            case TAPENV_RATIO_CP_S: ratio = ((C + P) * S) / 3; break;
            case TAPENV_RATIO_2C_SUM: ratio = (2 * C + P + S) / 3; break;
            case TAPENV_RATIO_CPS: ratio = (C * P * S) / 3; break;
            case TAPENV_RATIO_CP_HALF: ratio = (C + P) / 2; break;
            default: ratio = (C + P + S) / 3; break;
        }
    }
    reward[b] = (float)ratio;                      // scores[batch_index] = ... (model.py:510), fp32 tensor
}

// deterministic (fixed-order) reduction of (sum r, sum r^2, B) in fp64, single CTA
__global__ void reward_sums_kernel(int B, const float *__restrict__ reward, double *__restrict__ out) {
    __shared__ double s1[1024], s2[1024];
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
// launch helpers
// ------------------------------------------------------------------------------------
struct VecPlan { int vw; unsigned sv_magic; };

static VecPlan plan_vec(const DevCfg &d, const void *a, const void *b) {
    VecPlan p;
    const uintptr_t al = (uintptr_t)a | (uintptr_t)b;
    if (d.S % 4 == 0 && al % 16 == 0) p.vw = 4;
    else if (d.S % 2 == 0 && al % 8 == 0) p.vw = 2;
    else p.vw = 1;
    const unsigned SV = (unsigned)(d.S / p.vw);
    p.sv_magic = (unsigned)((0x100000000ull + SV - 1) / SV);
    return p;
}

#define TAPENV_DISPATCH_VW(plan, KERNEL, grid, block, stream, ...)                                   \
    do {                                                                                             \
        if ((plan).vw == 4) KERNEL<4, 5><<<grid, block, 0, stream>>>(__VA_ARGS__, (plan).sv_magic);   \
        else if ((plan).vw == 2) KERNEL<2, 5><<<grid, block, 0, stream>>>(__VA_ARGS__, (plan).sv_magic); \
        else KERNEL<1, 5><<<grid, block, 0, stream>>>(__VA_ARGS__, (plan).sv_magic);                  \
    } while (0)

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

void tapenv_set_envs_per_cta(int epc) { g_envs_per_cta = (epc == 1 || epc == 2 || epc == 4 || epc == 8) ? epc : 0; }

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
    cfg->batch = batch; cfg->blocks_num = blocks_num; cfg->dim = dim;
    cfg->rotate_types = allow_rot ? (dim == 2 ? 2 : 6) : 1;              // pack.py:306-309 (dim!)
    cfg->width = container_size[0];
    cfg->length = dim == 3 ? container_size[1] : 1;
    cfg->height = container_size[dim - 1];

    // packing strategy, with the reward-type override of tools.py:3617-3620
    if (!strcmp(packing_strategy, "LB_GREEDY")) cfg->strategy = TAPENV_LB_GREEDY;
    else if (!strcmp(packing_strategy, "MACS") || !strcmp(packing_strategy, "MUL")) cfg->strategy = TAPENV_MACS;
    else if (!strcmp(packing_strategy, "LB")) return TAPENV_EUNSUPPORTED;
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
    } else if (!strcmp(input_type, "mul") || !strcmp(input_type, "mul-with") || !strcmp(input_type, "rot-old")) {
        return TAPENV_EUNSUPPORTED;   // two-container inputs / legacy layout: out of scope
    } else return TAPENV_EENUM;
    return check_cfg(cfg);
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
    if ((cfg)->batch == 0) return TAPENV_OK;         \
    const DevCfg d = devcfg_of(cfg);                 \
    const int epc = pick_epc(d.B);                   \
    const dim3 block(32 * epc), grid((d.B + epc - 1) / epc); \
    cudaStream_t s = (cudaStream_t)stream;

int tapenv_reset(const tapenv_config *cfg, void *state, const float *dynamic, float *cur_mask_out, float *mask_out,
                 void *stream) {
    TAPENV_PROLOGUE(cfg)
    if (!state) return TAPENV_EINVAL;
    if (dynamic && !cur_mask_out) return TAPENV_EINVAL;
    const StatePtrs st = stateptrs_of(cfg, state);
    const VecPlan plan = plan_vec(d, dynamic, nullptr);
    TAPENV_DISPATCH_VW(plan, reset_kernel, grid, block, s, d, st, dynamic, cur_mask_out, mask_out);
    return launch_status();
}

int tapenv_update_dynamic(const tapenv_config *cfg, const float *dynamic, const float *static_, const int64_t *ptr,
                          float *dynamic_out, void *stream) {
    TAPENV_PROLOGUE(cfg)
    if (!dynamic || !static_ || !ptr || !dynamic_out) return TAPENV_EINVAL;
    const VecPlan plan = plan_vec(d, dynamic, dynamic_out);
    TAPENV_DISPATCH_VW(plan, update_dynamic_kernel, grid, block, s, d, dynamic, static_, ptr, dynamic_out);
    return launch_status();
}

int tapenv_update_mask(const tapenv_config *cfg, const float *mask, const float *dynamic, const int64_t *ptr,
                       float *new_mask_out, float *chosen_mask_out, void *stream) {
    TAPENV_PROLOGUE(cfg)
    if (!mask || !dynamic || !ptr || !new_mask_out || !chosen_mask_out) return TAPENV_EINVAL;
    const VecPlan plan = plan_vec(d, dynamic, nullptr);
    TAPENV_DISPATCH_VW(plan, update_mask_kernel, grid, block, s, d, mask, dynamic, ptr, new_mask_out, chosen_mask_out);
    return launch_status();
}

int tapenv_add_blocks(const tapenv_config *cfg, void *state, const float *blocks, float *dec_dynamic_out, void *stream) {
    TAPENV_PROLOGUE(cfg)
    rc_ = check_strategy_built(cfg);
    if (rc_ != TAPENV_OK) return rc_;
    if (!state || !blocks) return TAPENV_EINVAL;
    const StatePtrs st = stateptrs_of(cfg, state);
    add_blocks_lbg2d_kernel<<<grid, block, 0, s>>>(d, st, blocks, dec_dynamic_out);
    return launch_status();
}

int tapenv_step(const tapenv_config *cfg, void *state, const int64_t *ptr, const float *static_,
                const float *dynamic_in, const float *mask_in, float *dynamic_out, float *cur_mask_out,
                float *mask_out, float *dec_static_out, float *dec_dynamic_out, void *stream) {
    TAPENV_PROLOGUE(cfg)
    if (!state || !ptr || !static_ || !dynamic_in || !mask_in || !dynamic_out || !cur_mask_out || !mask_out)
        return TAPENV_EINVAL;
    rc_ = check_strategy_built(cfg);
    if (rc_ != TAPENV_OK) return rc_;
    const StatePtrs st = stateptrs_of(cfg, state);
    const VecPlan plan = plan_vec(d, dynamic_in, dynamic_out);
    TAPENV_DISPATCH_VW(plan, step_lbg2d_kernel, grid, block, s, d, st, ptr, static_, dynamic_in, mask_in, dynamic_out,
                       cur_mask_out, mask_out, dec_static_out, dec_dynamic_out);
    return launch_status();
}

int tapenv_reward(const tapenv_config *cfg, const void *state, float *reward_out, double *partial_sums_out, void *stream) {
    int rc_ = check_cfg(cfg);
    if (rc_ != TAPENV_OK) return rc_;
    if (cfg->batch > 0 && (!state || !reward_out)) return TAPENV_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    const DevCfg d = devcfg_of(cfg);
    const StatePtrs st = stateptrs_of(cfg, const_cast<void *>(state));
    if (d.B > 0) reward_kernel<<<(d.B + 127) / 128, 128, 0, s>>>(d, st, reward_out);
    if (partial_sums_out) reward_sums_kernel<<<1, 1024, 0, s>>>(d.B, reward_out, partial_sums_out);
    return launch_status();
}

int tapenv_episode(const tapenv_config *cfg, void *state, const float *static_, const float *dynamic,
                   const int64_t *ptr_seq, int32_t steps, float *reward_out, float *cur_mask_out, float *mask_out,
                   float *dec_dynamic_out, void *stream) {
    (void)cfg; (void)state; (void)static_; (void)dynamic; (void)ptr_seq; (void)steps; (void)reward_out;
    (void)cur_mask_out; (void)mask_out; (void)dec_dynamic_out; (void)stream;
    return TAPENV_EUNSUPPORTED;
}

}  // extern "C"
