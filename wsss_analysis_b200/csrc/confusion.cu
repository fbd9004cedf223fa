// confusion.cu -- integer confusion-matrix accumulation behind mIoU (sm_100a).
//
// Replaces chainercv.calc_semantic_segmentation_confusion as called at
// /root/reference/03b_irn/step/eval_sem_seg.py:41 (int64 C x C, bincount(C*gt + pred) over gt >= 0)
// and the per-class mask loops of /root/reference/03a_sec-dsrg/model.py:698-719.
// Pure integer counting: shared-memory histogram per CTA, one 64-bit global atomic per non-zero bin.
// Integer addition is order independent, so the result is bit-exact however the pixels are sharded
// over CTAs, streams or GPUs; the cross-GPU sum is a torch.distributed all_reduce(SUM, int64) over
// NCCL done by the host layer (wsss_analysis_b200/evaluation.py).
#include "common.cuh"

namespace dcrf {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxSmemBins = 8192;  // (C+1)*C <= 8192  <=>  C <= 90

__global__ void __launch_bounds__(kThreads) confusion_kernel(const int32_t *__restrict__ gt,
                                                             const int32_t *__restrict__ pred, int64_t n,
                                                             int C, unsigned long long *__restrict__ conf,
                                                             unsigned long long *__restrict__ n_bad,
                                                             int use_smem) {
    extern __shared__ unsigned int bins[];
    const int nbins = (C + 1) * C;
    if (use_smem) {
        for (int i = threadIdx.x; i < nbins; i += kThreads) bins[i] = 0;
        __syncthreads();
    }
    unsigned int bad = 0;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
        const int g = gt[i], p = pred[i];
        if (p < 0 || p >= C) {
            bad++;
            continue;
        }
        const int row = (g >= 0 && g < C) ? g : C;
        if (use_smem) atomicAdd(&bins[row * C + p], 1u);
        else atomicAdd(&conf[(int64_t)row * C + p], 1ull);
    }
    if (bad && n_bad) atomicAdd(n_bad, (unsigned long long)bad);
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < nbins; i += kThreads) {
            const unsigned int c = bins[i];
            if (c) atomicAdd(&conf[i], (unsigned long long)c);
        }
    }
}

}  // namespace
}  // namespace dcrf

using namespace dcrf;

extern "C" int dcrf_confusion_accumulate(const int32_t *gt, const int32_t *pred, int64_t n, int n_classes,
                                         int64_t *conf, int64_t *n_bad_pred, int device, void *stream) {
    try {
        DCRF_REQUIRE(gt && pred && conf, DCRF_EINVAL, "NULL argument");
        DCRF_REQUIRE(n_classes >= 1 && n_classes <= 4096, DCRF_EINVAL, "n_classes must be in [1, 4096]");
        DCRF_REQUIRE(n >= 0, DCRF_EINVAL, "n must be >= 0");
        if (n == 0) return DCRF_OK;
        int prev = -1;
        DCRF_CUDA(cudaGetDevice(&prev));
        if (device >= 0 && device != prev) DCRF_CUDA(cudaSetDevice(device));
        const int nbins = (n_classes + 1) * n_classes;
        const int use_smem = nbins <= kMaxSmemBins;
        int64_t want = (n + kThreads * 8 - 1) / (kThreads * 8);
        int nb = (int)std::min<int64_t>(std::max<int64_t>(want, 1), (int64_t)kNumSMs * 8);
        confusion_kernel<<<nb, kThreads, use_smem ? sizeof(unsigned int) * nbins : 0, (cudaStream_t)stream>>>(
            gt, pred, n, n_classes, (unsigned long long *)conf, (unsigned long long *)n_bad_pred, use_smem);
        cudaError_t e = cudaGetLastError();
        g_launches.fetch_add(1);
        if (device >= 0 && device != prev) cudaSetDevice(prev);
        DCRF_CUDA(e);
        return DCRF_OK;
    } catch (const Error &e) {
        set_error(e.msg);
        return e.code;
    }
}
