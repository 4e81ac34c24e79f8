// host_checks.cpp -- host build of the header-only integer geometry used by the kernels, so the CPU
// test-suite can check it exhaustively (tests/test_stable3d_cpu.py).  Not part of libtapenv.so and never
// used to produce results: the product path is CUDA only.
#include <stdint.h>
#include "stable3d.cuh"

extern "C" void tapenv_host_stable3d_masks(int bx, int by, const uint32_t *masks, int count, unsigned char *out) {
    for (int i = 0; i < count; ++i) out[i] = tapenv::stable3d_from_support(bx, by, masks[i]) ? 1 : 0;
}
