// resize.cu -- the two cv2.resize flavours the reference applies on either side of the CRF (sm_100a).
//
//   INTER_LINEAR on float32 HxWxC maps: network output -> ground-truth size BEFORE the CRF
//       (/root/reference/03a_sec-dsrg/model.py:686-687, :696)
//   INTER_NEAREST on label maps: CRF / PNG labels -> evaluation size BEFORE the confusion matrix
//       (/root/reference/03b_irn/step/eval_sem_seg.py:36, 03c_hsn/demo.py:181-183)
//
// Index / weight arithmetic restates OpenCV's resize.cpp (4.x): scale = 1 / ((double)dst / src);
// nearest: s = min(floor(d * scale), src - 1) -- no half-pixel shift;
// linear : f = (float)((d + 0.5) * scale - 0.5), s = floor(f), f -= s, clamped at both borders,
//          horizontal pass then vertical pass in float32.
// cv2 is available in the build image, so tests/golden/resize_*.npz are REAL cv2 outputs
// (tools/make_golden_resize.py): nearest is bit-exact; linear agrees to <= 1e-4 absolute on N(0,1)
// data -- OpenCV's vectorised float path rounds sample positions / weights in single precision
// differently from its own scalar formula, which is what is restated here.
#include <math.h>

#include "common.cuh"

namespace dcrf {
namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) resize_nearest_kernel(const int32_t *__restrict__ src, int sh, int sw,
                                                                  int32_t *__restrict__ dst, int dh, int dw,
                                                                  double ifx, double ify) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= (int64_t)dh * dw) return;
    const int y = (int)(i / dw), x = (int)(i - (int64_t)y * dw);
    const int sx = min((int)floor(x * ifx), sw - 1);
    const int sy = min((int)floor(y * ify), sh - 1);
    dst[i] = src[(int64_t)sy * sw + sx];
}

__device__ __forceinline__ void linear_coeff(int d, double scale, int ssize, int *s0, int *s1, float *w1) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= (float)s;
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }
    *s0 = s;
    *s1 = min(s + 1, ssize - 1);
    *w1 = f;
}

// one thread per destination (pixel, channel)
__global__ void __launch_bounds__(kThreads) resize_linear_kernel(const float *__restrict__ src, int sh, int sw, int C,
                                                                 float *__restrict__ dst, int dh, int dw,
                                                                 double scale_x, double scale_y) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= (int64_t)dh * dw * C) return;
    const int c = (int)(i % C);
    const int64_t pix = i / C;
    const int y = (int)(pix / dw), x = (int)(pix - (int64_t)y * dw);
    int x0, x1, y0, y1;
    float fx, fy;
    linear_coeff(x, scale_x, sw, &x0, &x1, &fx);
    linear_coeff(y, scale_y, sh, &y0, &y1, &fy);
    const float a0 = 1.f - fx, a1 = fx, b0 = 1.f - fy, b1 = fy;
    const float *r0 = src + ((int64_t)y0 * sw) * C + c, *r1 = src + ((int64_t)y1 * sw) * C + c;
    const float t0 = __fadd_rn(__fmul_rn(r0[(int64_t)x0 * C], a0), __fmul_rn(r0[(int64_t)x1 * C], a1));
    const float t1 = __fadd_rn(__fmul_rn(r1[(int64_t)x0 * C], a0), __fmul_rn(r1[(int64_t)x1 * C], a1));
    dst[i] = __fadd_rn(__fmul_rn(t0, b0), __fmul_rn(t1, b1));
}

struct DevGuard {
    int prev = -1;
    explicit DevGuard(int dev) {
        DCRF_CUDA(cudaGetDevice(&prev));
        if (dev >= 0 && dev != prev) DCRF_CUDA(cudaSetDevice(dev));
        else prev = -1;
    }
    ~DevGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

}  // namespace
}  // namespace dcrf

using namespace dcrf;

extern "C" int dcrf_resize_nearest_i32(const int32_t *src, int sh, int sw, int32_t *dst, int dh, int dw, int device,
                                       void *stream) {
    try {
        DCRF_REQUIRE(src && dst, DCRF_EINVAL, "NULL argument");
        DCRF_REQUIRE(sh >= 1 && sw >= 1 && dh >= 1 && dw >= 1, DCRF_EINVAL, "sizes must be >= 1");
        DevGuard guard(device);
        const double ifx = 1.0 / ((double)dw / sw), ify = 1.0 / ((double)dh / sh);
        resize_nearest_kernel<<<ceil_div((int64_t)dh * dw, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
            src, sh, sw, dst, dh, dw, ifx, ify);
        DCRF_LAUNCHED();
        return DCRF_OK;
    } catch (const Error &e) {
        set_error(e.msg);
        return e.code;
    }
}

extern "C" int dcrf_resize_bilinear_f32(const float *src, int sh, int sw, int channels, float *dst, int dh, int dw,
                                        int device, void *stream) {
    try {
        DCRF_REQUIRE(src && dst, DCRF_EINVAL, "NULL argument");
        DCRF_REQUIRE(sh >= 1 && sw >= 1 && dh >= 1 && dw >= 1 && channels >= 1, DCRF_EINVAL, "sizes must be >= 1");
        DevGuard guard(device);
        const double scale_x = 1.0 / ((double)dw / sw), scale_y = 1.0 / ((double)dh / sh);
        resize_linear_kernel<<<ceil_div((int64_t)dh * dw * channels, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
            src, sh, sw, channels, dst, dh, dw, scale_x, scale_y);
        DCRF_LAUNCHED();
        return DCRF_OK;
    } catch (const Error &e) {
        set_error(e.msg);
        return e.code;
    }
}
