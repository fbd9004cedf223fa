// Micro-benchmark: gather of random 96-byte rows from a 387 MB table (the bilateral lattice values of
// a 32-image batch), (a) with LDG.128 by groups of 6 lanes, as the slice / splat kernels do, and
// (b) with one cp.async.bulk (TMA, UBLKCP) per row into shared memory followed by conflict-free
// LDS.128.  Question: can the TMA path deliver small rows faster than the LSU data pipe?
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int kRowF4 = 6;           // 96-byte rows
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

__device__ __forceinline__ uint32_t lcg(uint32_t x) { return x * 1664525u + 1013904223u; }

// (a) 5 rows per warp request (30 lanes), like the product kernels; STRIDE = row pitch in float4
// (6 = dense 96-byte rows, half of them straddle two 128-byte lines; 8 = one row per line)
template <int STRIDE>
__global__ void __launch_bounds__(kThreads) ldg_gather(const float4 *__restrict__ tab, uint32_t rows, int iters,
                                                       float *out) {
    const int lane = threadIdx.x & 31, sub = lane / kRowF4, c = lane % kRowF4;
    uint32_t seed = (blockIdx.x * kThreads + threadIdx.x - c) * 2654435761u + 12345u;  // same per group
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int it = 0; it < iters; it++) {
        float4 v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            seed = lcg(seed);
            const uint32_t r = (seed >> 4) % rows;
            v[k] = sub < 5 ? __ldg(tab + (size_t)r * STRIDE + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < 8; k++) {
            acc.x += v[k].x; acc.y += v[k].y; acc.z += v[k].z; acc.w += v[k].w;
        }
    }
    if (acc.x + acc.y + acc.z + acc.w == 12345.678f) out[0] = acc.x;
}

// (a2) narrower loads: W bytes per lane (8 or 4), 96 / W lanes per row, whole rows per warp request
template <int W>
__global__ void __launch_bounds__(kThreads) ldg_gather_narrow(const float *__restrict__ tab, uint32_t rows, int iters,
                                                              float *out) {
    constexpr int LPR = 96 / W;        // lanes per row: 12 or 24
    constexpr int RPW = 32 / LPR;      // rows per warp request: 2 or 1
    const int lane = threadIdx.x & 31, sub = lane / LPR, c = lane % LPR;
    uint32_t seed = (blockIdx.x * kThreads + threadIdx.x - c) * 2654435761u + 12345u;
    float acc = 0.f;
    for (int it = 0; it < iters; it++) {
        float v[8][W / 4];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            seed = lcg(seed);
            const uint32_t r = (seed >> 4) % rows;
            const float *p = tab + (size_t)r * 24 + c * (W / 4);
            if (sub < RPW) {
                if (W == 8) {
                    const float2 t = __ldg(reinterpret_cast<const float2 *>(p));
                    v[k][0] = t.x;
                    v[k][W / 4 - 1] = t.y;
                } else {
                    v[k][0] = __ldg(p);
                }
            } else {
                v[k][0] = 0.f;
                v[k][W / 4 - 1] = 0.f;
            }
        }
#pragma unroll
        for (int k = 0; k < 8; k++) acc += v[k][0] + v[k][W / 4 - 1];
    }
    if (acc == 12345.678f) out[0] = acc;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// (b) every lane issues one 96-byte bulk copy per round; ROUNDS rounds are in flight per warp
template <int ROUNDS>
__global__ void __launch_bounds__(kThreads) bulk_gather(const float4 *__restrict__ tab, uint32_t rows, int iters,
                                                        float *out) {
    extern __shared__ __align__(128) unsigned char dyn[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(dyn);
    float4(*stage)[ROUNDS][32 * kRowF4] = reinterpret_cast<float4(*)[ROUNDS][32 * kRowF4]>(dyn + 128);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t seed = (blockIdx.x * kThreads + threadIdx.x) * 2654435761u + 12345u;
    const uint32_t b = smem_u32(&bar[w]);
    if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    __syncwarp();
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t phase = 0;
    for (int it = 0; it < iters; it++) {
        if (lane == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b),
                         "r"(ROUNDS * 32 * kRowF4 * 16));
        __syncwarp();
#pragma unroll
        for (int k = 0; k < ROUNDS; k++) {
            seed = lcg(seed);
            const uint32_t r = (seed >> 4) % rows;
            const uint32_t dst = smem_u32(&stage[w][k][lane * kRowF4]);
            asm volatile(
                "cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                "l"(tab + (size_t)r * kRowF4), "r"(kRowF4 * 16), "r"(b)
                : "memory");
        }
        uint32_t done = 0;
        while (!done)
            asm volatile(
                "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                : "=r"(done)
                : "r"(b), "r"(phase)
                : "memory");
        phase ^= 1;
        // conflict-free read-out: lane l takes float4 l, l + 32, ... of the warp's contiguous staging
#pragma unroll
        for (int k = 0; k < ROUNDS; k++)
#pragma unroll
            for (int j = 0; j < kRowF4; j++) {
                const float4 v = stage[w][k][j * 32 + lane];
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        __syncwarp();
    }
    if (acc.x + acc.y + acc.z + acc.w == 12345.678f) out[0] = acc.x;
}

int main(int argc, char **argv) {
    const uint32_t rows = argc > 1 ? (uint32_t)atol(argv[1]) : (4u << 20);
    printf("table: %u rows of 96 B = %.1f MB\n", rows, rows * 96.0 / 1e6);
    float4 *tab;
    float *out;
    cudaMalloc(&tab, (size_t)rows * 8 * sizeof(float4));
    cudaMalloc(&out, 4);
    cudaMemset(tab, 0, (size_t)rows * 8 * sizeof(float4));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int grid = 148 * 8;
    cudaFuncSetAttribute(bulk_gather<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 + kWarps * 2 * 32 * kRowF4 * 16);
    float ms;
    for (int rep = 0; rep < 4; rep++) {
        const int iters = 64;
        cudaEventRecord(e0);
        if (rep < 2) ldg_gather<6><<<grid, kThreads>>>(tab, rows, iters, out);
        else ldg_gather<8><<<grid, kThreads>>>(tab, rows, iters, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double n = (double)grid * kWarps * 5 * 8 * iters;
        printf("LDG.128 groups of 6 lanes, pitch %d B: %.3f ms, %.1f G rows/s, %.0f GB/s\n", rep < 2 ? 96 : 128, ms,
               n / ms / 1e6, n * 96 / ms / 1e6);
    }
    for (int rep = 0; rep < 4; rep++) {
        const int iters = 64;
        cudaEventRecord(e0);
        if (rep < 2) ldg_gather_narrow<8><<<grid, kThreads>>>(reinterpret_cast<const float *>(tab), rows, iters, out);
        else ldg_gather_narrow<4><<<grid, kThreads>>>(reinterpret_cast<const float *>(tab), rows, iters, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double n = (double)grid * kWarps * (rep < 2 ? 2 : 1) * 8 * iters;
        printf("LDG.%d, %d lanes per row: %.3f ms, %.1f G rows/s, %.0f GB/s\n", rep < 2 ? 64 : 32, rep < 2 ? 12 : 24, ms,
               n / ms / 1e6, n * 96 / ms / 1e6);
    }
    for (int rep = 0; rep < 2; rep++) {
        const int iters = 64;
        cudaEventRecord(e0);
        bulk_gather<1><<<grid, kThreads, 128 + kWarps * 1 * 32 * kRowF4 * 16>>>(tab, rows, iters, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double n = (double)grid * kThreads * 1 * iters;
        printf("cp.async.bulk 96 B, 1 round : %.3f ms, %.1f G rows/s, %.0f GB/s\n", ms, n / ms / 1e6, n * 96 / ms / 1e6);
    }
    for (int rep = 0; rep < 2; rep++) {
        const int iters = 32;
        cudaEventRecord(e0);
        bulk_gather<2><<<grid, kThreads, 128 + kWarps * 2 * 32 * kRowF4 * 16>>>(tab, rows, iters, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double n = (double)grid * kThreads * 2 * iters;
        printf("cp.async.bulk 96 B, 2 rounds: %.3f ms, %.1f G rows/s, %.0f GB/s\n", ms, n / ms / 1e6, n * 96 / ms / 1e6);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
