// copy_probe.cu -- the C2 floor experiment (VERDICT r01 "weak" 5 / item 10): which data-movement instruction flavour moves the
// out-of-place copy of `dynamic` [4096, 30, 20] f32 (9.83 MB in, 9.83 MB out: 93 % of a C2 step's bytes) fastest?
//   v4      : ld.global.nc.L1::no_allocate.v4 / st.global.L1::no_allocate.v4     (what step_kernel uses), one warp per environment
//   v8      : sm_100 256-bit ld.global.nc.L1::no_allocate.v8.f32 / st.global.v8.f32, flat mapping
//   bulk    : cp.async.bulk global -> shared (mbarrier) then cp.async.bulk shared -> global, one 2400-byte row per environment
//   flat v4 : grid-stride float4 copy (no per-environment structure), the shape torch's copy kernel has
// Each variant: graph-free, 1 warm-up + 50 timed launches alternating over RING cold buffers (ring > L2), CUDA events.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o copy_probe scripts/copy_probe.cu && ./copy_probe [B]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int ENV_FLOATS = 600;                 // 30 x 20
constexpr int ENV_V4 = ENV_FLOATS / 4;          // 150
constexpr int ENV_V8 = ENV_FLOATS / 8;          // 75
constexpr int ENV_BYTES = ENV_FLOATS * 4;       // 2400

__device__ __forceinline__ uint4 ld4(const void *p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st4(void *p, const uint4 &v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
struct V8 { unsigned a[8]; };
__device__ __forceinline__ V8 ld8(const void *p) {
    V8 v;
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v.a[0]), "=r"(v.a[1]), "=r"(v.a[2]), "=r"(v.a[3]), "=r"(v.a[4]), "=r"(v.a[5]), "=r"(v.a[6]), "=r"(v.a[7]) : "l"(p));
    return v;
}
__device__ __forceinline__ void st8(void *p, const V8 &v) {
    asm volatile("st.global.L1::no_allocate.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "r"(v.a[0]), "r"(v.a[1]), "r"(v.a[2]), "r"(v.a[3]), "r"(v.a[4]), "r"(v.a[5]), "r"(v.a[6]), "r"(v.a[7]) : "memory");
}

// one warp per environment, 4 warps per CTA, all loads issued before the first store (step_kernel's shape)
__global__ void __launch_bounds__(128) copy_v4(const float *__restrict__ in, float *__restrict__ out, int B) {
    const int b = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= B) return;
    const uint4 *s = reinterpret_cast<const uint4 *>(in) + (size_t)b * ENV_V4 + lane;
    uint4 *d = reinterpret_cast<uint4 *>(out) + (size_t)b * ENV_V4 + lane;
    uint4 v[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) if (lane + 32 * i < ENV_V4) v[i] = ld4(s + 32 * i);
#pragma unroll
    for (int i = 0; i < 5; ++i) if (lane + 32 * i < ENV_V4) st4(d + 32 * i, v[i]);
}

__global__ void __launch_bounds__(128) copy_v8(const float *__restrict__ in, float *__restrict__ out, int B) {
    const int b = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= B) return;
    const char *s = reinterpret_cast<const char *>(in) + (size_t)b * ENV_BYTES + lane * 32;
    char *d = reinterpret_cast<char *>(out) + (size_t)b * ENV_BYTES + lane * 32;
    V8 v[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) if (lane + 32 * i < ENV_V8) v[i] = ld8(s + 1024 * i);
#pragma unroll
    for (int i = 0; i < 3; ++i) if (lane + 32 * i < ENV_V8) st8(d + 1024 * i, v[i]);
}

// grid-stride flat copies
__global__ void __launch_bounds__(256) copy_flat4(const uint4 *__restrict__ in, uint4 *__restrict__ out, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) st4(out + i, ld4(in + i));
}
__global__ void __launch_bounds__(256) copy_flat8(const char *__restrict__ in, char *__restrict__ out, size_t n32) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n32; i += stride) st8(out + i * 32, ld8(in + i * 32));
}

// TMA-style 1-D bulk copies: EPC environments per CTA, one elected thread issues global->shared bulk copies onto an mbarrier,
// waits, then issues shared->global bulk copies.  No register staging at all.
template <int EPC>
__global__ void __launch_bounds__(32) copy_bulk(const float *__restrict__ in, float *__restrict__ out, int B) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar;
    const int b0 = blockIdx.x * EPC;
    const int ne = min(EPC, B - b0);
    if (ne <= 0) return;
    const unsigned bar_a = (unsigned)__cvta_generic_to_shared(&bar);
    const unsigned sm_a = (unsigned)__cvta_generic_to_shared(smem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const unsigned bytes = (unsigned)ne * ENV_BYTES;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_a), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(sm_a), "l"(in + (size_t)b0 * ENV_FLOATS), "r"(bytes), "r"(bar_a) : "memory");
        unsigned done = 0;
        while (!done) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(bar_a) : "memory");
        }
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                     :: "l"(out + (size_t)b0 * ENV_FLOATS), "r"(sm_a), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

template <typename F>
static float time_it(F launch, int ring, int iters) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < ring; ++i) launch(i);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) launch(i % ring);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    return 1e3f * ms / iters;
}

int main(int argc, char **argv) {
    const int B = argc > 1 ? atoi(argv[1]) : 4096;
    const size_t bytes = (size_t)B * ENV_BYTES;
    const int ring = (int)(400e6 / (2.0 * bytes)) + 2;           // ring of distinct in/out pairs >> 126 MB L2
    float **in = new float *[ring], **out = new float *[ring];
    for (int i = 0; i < ring; ++i) { CK(cudaMalloc(&in[i], bytes)); CK(cudaMalloc(&out[i], bytes)); CK(cudaMemset(in[i], i + 1, bytes)); }
    const int iters = 200;
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("{\"B\": %d, \"MB_each_way\": %.2f, \"ring\": %d", B, bytes / 1e6, ring);
    float us;
    us = time_it([&](int i) { copy_v4<<<(B + 3) / 4, 128>>>(in[i], out[i], B); }, ring, iters);
    printf(", \"warp_per_env_v4_us\": %.2f", us);
    us = time_it([&](int i) { copy_v8<<<(B + 3) / 4, 128>>>(in[i], out[i], B); }, ring, iters);
    printf(", \"warp_per_env_v8_us\": %.2f", us);
    us = time_it([&](int i) { copy_flat4<<<sms * 8, 256>>>((const uint4 *)in[i], (uint4 *)out[i], bytes / 16); }, ring, iters);
    printf(", \"flat_v4_us\": %.2f", us);
    us = time_it([&](int i) { copy_flat8<<<sms * 8, 256>>>((const char *)in[i], (char *)out[i], bytes / 32); }, ring, iters);
    printf(", \"flat_v8_us\": %.2f", us);
    CK(cudaFuncSetAttribute(copy_bulk<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * ENV_BYTES));
    us = time_it([&](int i) { copy_bulk<4><<<(B + 3) / 4, 32, 4 * ENV_BYTES>>>(in[i], out[i], B); }, ring, iters);
    printf(", \"bulk_4env_per_cta_us\": %.2f", us);
    CK(cudaFuncSetAttribute(copy_bulk<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * ENV_BYTES));
    us = time_it([&](int i) { copy_bulk<16><<<(B + 15) / 16, 32, 16 * ENV_BYTES>>>(in[i], out[i], B); }, ring, iters);
    printf(", \"bulk_16env_per_cta_us\": %.2f", us);
    us = time_it([&](int i) { CK(cudaMemcpyAsync(out[i], in[i], bytes, cudaMemcpyDeviceToDevice)); }, ring, iters);
    printf(", \"cudaMemcpyAsync_us\": %.2f", us);
    printf(", \"note\": \"back-to-back eager launches on one stream: per-launch time includes launch gaps; 2*MB/us = GB/s\"}\n");
    // verify one variant actually copied
    unsigned char *h = (unsigned char *)malloc(16); CK(cudaMemcpy(h, out[1], 16, cudaMemcpyDeviceToHost));
    if (h[0] != 2) { printf("copy check failed\n"); return 1; }
    return 0;
}
