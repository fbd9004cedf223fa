// Micro-benchmark: streaming copy of rows of 6 float4 (96 B) stored at a stride of 6 (dense) or
// 8 float4 (one row per 128-byte line, last sector never touched).  Answers: does HBM traffic grow
// when only 3 of the 4 sectors of every line are touched?
#include <cstdio>
#include <cuda_runtime.h>
template <int STRIDE>
__global__ void copy_rows(const float4 *__restrict__ in, float4 *__restrict__ out, long rows) {
    long e = (long)blockIdx.x * blockDim.x + threadIdx.x;  // element = (row, col<6)
    for (; e < rows * 6; e += (long)gridDim.x * blockDim.x) {
        long r = e / 6;
        int c = (int)(e - r * 6);
        float4 v = in[r * STRIDE + c];
        v.x += 1.f;
        out[r * STRIDE + c] = v;
    }
}
int main() {
    const long rows = 4L << 20;  // 4 M rows: 403 MB dense, 537 MB strided
    float4 *a, *b;
    cudaMalloc(&a, rows * 8 * sizeof(float4));
    cudaMalloc(&b, rows * 8 * sizeof(float4));
    cudaMemset(a, 0, rows * 8 * sizeof(float4));
    cudaMemset(b, 0, rows * 8 * sizeof(float4));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int variant = 0; variant < 2; variant++) {
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0);
            for (int i = 0; i < 10; i++) {
                if (variant == 0) copy_rows<6><<<148 * 16, 256>>>(a, b, rows);
                else copy_rows<8><<<148 * 16, 256>>>(a, b, rows);
            }
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            printf("stride %d float4: %.1f us per copy, %.0f GB/s useful\n", variant ? 8 : 6, ms * 100,
                   2.0 * rows * 96 / (ms / 10 * 1e-3) / 1e9);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
