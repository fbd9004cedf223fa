// Micro-benchmark: Blackwell TMA row gather (cp.async.bulk.tensor.2d ... tile::gather4: four rows of a
// 2-D tensor map per instruction, SASS UTMALDG) of random 96-byte rows into shared memory, against
// the LDG.128 lane-group gather the slice / splat kernels use (tools/micro/bulk_gather.cu: 106.7 G
// rows/s L2-resident) and the one-row-per-instruction bulk copy (65 G rows/s).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather4 gather4.cu   (no -lcuda: the
//   tensor-map encoder is fetched with cudaGetDriverEntryPoint)
//   ./gather4 [rows]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

constexpr int kRowF4 = 6;  // 96-byte rows
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

__device__ __forceinline__ uint32_t lcg(uint32_t x) { return x * 1664525u + 1013904223u; }
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// every warp runs its own pipeline: per iteration ROUNDS x 32 rows = ROUNDS x 8 gather4 instructions
// (lanes 0..7 issue one each per round), one mbarrier per warp, then a conflict-free LDS read-out.
template <int ROUNDS>
__global__ void __launch_bounds__(kThreads) tma_gather4(const __grid_constant__ CUtensorMap map, uint32_t rows,
                                                        int iters, float *out) {
    extern __shared__ __align__(128) unsigned char dyn[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(dyn);
    float4(*stage)[ROUNDS][32 * kRowF4] = reinterpret_cast<float4(*)[ROUNDS][32 * kRowF4]>(dyn + 128);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t seed = (blockIdx.x * kThreads + threadIdx.x) * 2654435761u + 12345u;
    const uint32_t b = smem_u32(&bar[w]);
    if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t phase = 0;
    for (int it = 0; it < iters; it++) {
        if (lane == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b),
                         "r"(ROUNDS * 32 * kRowF4 * 16));
        __syncwarp();
        if (lane < 8) {
#pragma unroll
            for (int k = 0; k < ROUNDS; k++) {
                int r[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    seed = lcg(seed);
                    r[j] = (int)((seed >> 4) % rows);
                }
                const uint32_t dst = smem_u32(&stage[w][k][lane * 4 * kRowF4]);
                asm volatile(
                    "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes"
                    " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
                    "l"(&map), "r"(b), "r"(0), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3])
                    : "memory");
            }
        }
        uint32_t done = 0;
        while (!done)
            asm volatile(
                "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                : "=r"(done)
                : "r"(b), "r"(phase)
                : "memory");
        phase ^= 1;
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

// producer / consumer split: warp 0 issues gather4 for a ring of STAGES stages of 32 rows, the other
// warps read the rows out; keeps the TMA queue full while rows are being consumed.
template <int STAGES>
__global__ void __launch_bounds__(kThreads) tma_gather4_ring(const __grid_constant__ CUtensorMap map, uint32_t rows,
                                                             int iters, float *out) {
    extern __shared__ __align__(128) unsigned char dyn[];
    uint64_t *full = reinterpret_cast<uint64_t *>(dyn);
    uint64_t *empty = full + STAGES;
    float4(*stage)[32 * kRowF4] = reinterpret_cast<float4(*)[32 * kRowF4]>(dyn + 1024);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    constexpr int kConsumers = kThreads - 32;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&empty[s])), "r"(kConsumers));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto wait = [](uint32_t bar, uint32_t ph) {
        uint32_t done = 0;
        while (!done)
            asm volatile(
                "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                : "=r"(done)
                : "r"(bar), "r"(ph)
                : "memory");
    };
    if (w == 0) {
        uint32_t seed = (blockIdx.x * 32 + lane) * 2654435761u + 12345u;
        for (int it = 0; it < iters; it++) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            if (it >= STAGES) wait(smem_u32(&empty[s]), ph ^ 1);
            if (lane == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])),
                             "r"(32 * kRowF4 * 16));
            __syncwarp();
            if (lane < 8) {
                int r[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    seed = lcg(seed);
                    r[j] = (int)((seed >> 4) % rows);
                }
                asm volatile(
                    "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes"
                    " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(&stage[s][lane * 4 * kRowF4])),
                    "l"(&map), "r"(smem_u32(&full[s])), "r"(0), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3])
                    : "memory");
            }
        }
    } else {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const int t = threadIdx.x - 32;  // 224 consumer threads, 192 float4 per stage
        for (int it = 0; it < iters; it++) {
            const int s = it % STAGES;
            wait(smem_u32(&full[s]), (it / STAGES) & 1);
            if (t < 32 * kRowF4) {
                const float4 v = stage[s][t];
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
        }
        if (acc.x + acc.y + acc.z + acc.w == 12345.678f) out[0] = acc.x;
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                             const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

#define CK(x)                                                                             \
    do {                                                                                  \
        cudaError_t e_ = (x);                                                             \
        if (e_ != cudaSuccess) {                                                          \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                      \
        }                                                                                 \
    } while (0)

template <typename F>
static double time_ms(F f) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    f();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    f();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    return ms;
}

static void run_table(EncodeFn encode, uint32_t rows, int pitch_f4) {
    float4 *tab;
    const size_t bytes = (size_t)rows * pitch_f4 * 16;
    CK(cudaMalloc(&tab, bytes));
    CK(cudaMemset(tab, 0, bytes));
    float *out;
    CK(cudaMalloc(&out, 4));
    CUtensorMap map;
    const cuuint64_t gdim[2] = {24, rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)pitch_f4 * 16};
    const cuuint32_t box[2] = {24, 1};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, tab, gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
        exit(1);
    }
    const int iters = 200;
    printf("table %.1f MB (%u rows, pitch %d B)\n", bytes / 1e6, rows, pitch_f4 * 16);
#define RUN_WARP(R, BPS)                                                                                    \
    {                                                                                                       \
        const size_t smem = 128 + (size_t)kWarps * R * 32 * kRowF4 * 16;                                    \
        CK(cudaFuncSetAttribute(tma_gather4<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
        const int grid = 148 * BPS;                                                                         \
        const double ms = time_ms([&] { tma_gather4<R><<<grid, kThreads, smem>>>(map, rows, iters, out); }); \
        CK(cudaGetLastError());                                                                             \
        const double n = (double)grid * kWarps * iters * R * 32;                                            \
        printf("  gather4 per-warp pipeline, %d rounds in flight, %d CTAs/SM : %7.1f G rows/s\n", R, BPS,   \
               n / (ms * 1e-3) / 1e9);                                                                      \
    }
    RUN_WARP(1, 8)
    RUN_WARP(2, 4)
    RUN_WARP(4, 2)
    RUN_WARP(4, 4)
#define RUN_RING(S, BPS)                                                                                        \
    {                                                                                                           \
        const size_t smem = 1024 + (size_t)S * 32 * kRowF4 * 16;                                                \
        CK(cudaFuncSetAttribute(tma_gather4_ring<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
        const int grid = 148 * BPS;                                                                             \
        const int it2 = iters * 8;                                                                              \
        const double ms = time_ms([&] { tma_gather4_ring<S><<<grid, kThreads, smem>>>(map, rows, it2, out); }); \
        CK(cudaGetLastError());                                                                                 \
        const double n = (double)grid * it2 * 32;                                                               \
        printf("  gather4 producer warp + ring of %2d stages, %d CTAs/SM      : %7.1f G rows/s\n", S, BPS,      \
               n / (ms * 1e-3) / 1e9);                                                                          \
    }
    RUN_RING(8, 4)
    RUN_RING(16, 4)
    RUN_RING(16, 8)
    RUN_RING(32, 2)
    CK(cudaFree(tab));
    CK(cudaFree(out));
}

int main(int argc, char **argv) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn) {
        printf("cuTensorMapEncodeTiled not available\n");
        return 1;
    }
    EncodeFn encode = (EncodeFn)fn;
    const uint32_t sizes[3] = {131072u, 1024u, 4194304u};
    for (int i = 0; i < 3; i++) {
        run_table(encode, argc > 1 ? (uint32_t)atoi(argv[1]) : sizes[i], 6);
        if (argc > 1) break;
    }
    run_table(encode, 131072u, 8);  // one row per 128-byte line
    return 0;
}
