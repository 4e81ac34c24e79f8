// host_checks.cpp -- host build of the header-only logic the kernels share, so the CPU test-suite can check it against the
// oracle without a GPU: the integer geometry of is_stable (tests/test_stable3d_cpu.py) and the MACS 3D placement with ONE
// emulated lane (tests/test_macs3d_host_cpu.py).  Not part of libtapenv.so and never used to produce results: the product
// path is CUDA only.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "stable3d.cuh"
#include "place_macs3d.cuh"

extern "C" void tapenv_host_stable3d_masks(int bx, int by, const uint32_t *masks, int count, unsigned char *out) {
    for (int i = 0; i < count; ++i) out[i] = tapenv::stable3d_from_support(bx, by, masks[i]) ? 1 : 0;
}

// One environment through `steps` blocks [steps][3] with the state laid out as the device keeps it (reset as reset_env_state
// does).  Outputs after EVERY step: heightmap [steps][W*L], scal [steps][4]; at the end: positions [cap][3], stable [cap],
// voxels [W*L*H] (int16), lists [H*L*lcap] (int8).  Returns the OR of the anomaly bits.
extern "C" int tapenv_host_macs3d_episode(int W, int L, int H, int cap, int lcap, int flags, int steps, const int *blocks_in,
                                          int *heightmaps, int *scals, int *positions, unsigned char *stable, short *voxels,
                                          signed char *lists) {
    using namespace tapenv;
    const int cells = W * L;
    memset(voxels, 0, sizeof(short) * cells * H);
    for (int i = 0; i < H * L * lcap; ++i) { const int q = i % lcap; lists[i] = q == 0 ? 2 : (q == 2 ? (signed char)(W - 1) : 0); }
    int *h = (int *)calloc(cells, sizeof(int)), *blks = (int *)calloc(cap * 3, sizeof(int));
    memset(positions, 0, sizeof(int) * cap * 3); memset(stable, 0, cap);
    int scal[4] = {0, 0, 0, 0};
    M3Scratch *sm = (M3Scratch *)malloc(sizeof(M3Scratch));
    M3State s;
    s.vox = voxels; s.lists = lists; s.h = h; s.W = W; s.L = L; s.H = H; s.cells = cells; s.lcap = lcap; s.sm = sm; s.lane = 0; s.nl = 1;
    int anomaly = 0;
    for (int t = 0; t < steps; ++t) {
        anomaly |= macs3d_env_add_block(flags, cap, s, scal, positions, blks, stable, blocks_in[t * 3], blocks_in[t * 3 + 1], blocks_in[t * 3 + 2]);
        memcpy(heightmaps + (size_t)t * cells, h, sizeof(int) * cells);
        memcpy(scals + (size_t)t * 4, scal, sizeof(scal));
    }
    free(h); free(blks); free(sm);
    return anomaly;
}

#include "place_lb.cuh"

// The LB strategy, warp form with one emulated lane (warp_form != 0) or the one-thread walk (warp_form == 0); same outputs as
// tapenv_host_macs3d_episode, blocks_in / positions are [steps][dim] / [cap][dim], lists are unsigned bytes ([0] = length).
template <int DIM>
static int lb_host_episode(int W, int L, int H, int cap, int lcap, int flags, int steps, const int *blocks_in, int warp_form,
                           int *heightmaps, int *scals, int *positions, unsigned char *stable, short *voxels, unsigned char *lists) {
    using namespace tapenv;
    const int cells = W * L;
    memset(voxels, 0, sizeof(short) * cells * H);
    for (int i = 0; i < H * L * lcap; ++i) lists[i] = (i % lcap) == 0 ? 1 : 0;      // every x list = [0] (tools.py:3649-3653)
    int *h = (int *)calloc(cells, sizeof(int)), *blks = (int *)calloc(cap * DIM, sizeof(int));
    memset(positions, 0, sizeof(int) * cap * DIM); memset(stable, 0, cap);
    int scal[4] = {0, 0, 0, 0};
    LbScratch *sm = (LbScratch *)malloc(sizeof(LbScratch));
    LbState s;
    s.vox = voxels; s.lists = lists; s.h = h; s.W = W; s.L = L; s.H = H; s.cells = cells; s.lcap = lcap;
    LbWarp w; w.sm = sm; w.lane = 0; w.nl = 1;
    int anomaly = 0;
    for (int t = 0; t < steps; ++t) {
        const int bx = blocks_in[t * DIM], by = DIM == 3 ? blocks_in[t * DIM + 1] : 1, bz = blocks_in[t * DIM + DIM - 1];
        if (warp_form) anomaly |= lb_env_add_block_warp<DIM>(flags, cap, s, w, scal, positions, blks, stable, bx, by, bz);
        else {                                          // the adapter of tapenv.cu (lb_env_add_block), restated for the host
            const int k = scal[3];
            if (k >= cap) { anomaly |= 2; } else {
                blks[k * DIM] = bx; if (DIM == 3) blks[k * DIM + 1] = by; blks[k * DIM + DIM - 1] = bz;
                unsigned char stb = 0;
                int o0 = scal[0], o1 = scal[1], o2 = scal[2];
                if (bx >= 1 && by >= 1 && bz >= 1 && bx <= W && by <= L) {
                    const int vol = bx * by * bz;
                    int a = 0;
                    const LbBest best = lb_place<DIM>(flags, s, k, positions, blks, bx, by, bz, scal[0] + vol, scal[1], scal[2], a);
                    if (best.any) {
                        lb_commit<DIM>(s, k, best, bx, by, bz, a);
                        if (!(a & 1)) {
                            positions[k * DIM] = best.x; if (DIM == 3) positions[k * DIM + 1] = best.y; positions[k * DIM + DIM - 1] = best.z;
                            stb = (unsigned char)best.stable; o0 += vol; o1 += best.add; o2 += best.stable;
                        }
                    }
                    anomaly |= a;
                }
                stable[k] = stb; scal[0] = o0; scal[1] = o1; scal[2] = o2; scal[3] = k + 1;
            }
        }
        memcpy(heightmaps + (size_t)t * cells, h, sizeof(int) * cells);
        memcpy(scals + (size_t)t * 4, scal, sizeof(scal));
    }
    free(h); free(blks); free(sm);
    return anomaly;
}

extern "C" int tapenv_host_lb_episode(int dim, int W, int L, int H, int cap, int lcap, int flags, int steps, const int *blocks_in,
                                      int warp_form, int *heightmaps, int *scals, int *positions, unsigned char *stable,
                                      short *voxels, unsigned char *lists) {
    if (dim == 2) return lb_host_episode<2>(W, 1, H, cap, lcap, flags, steps, blocks_in, warp_form, heightmaps, scals, positions, stable, voxels, lists);
    return lb_host_episode<3>(W, L, H, cap, lcap, flags, steps, blocks_in, warp_form, heightmaps, scals, positions, stable, voxels, lists);
}
