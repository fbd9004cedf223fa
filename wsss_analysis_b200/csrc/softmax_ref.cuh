// softmax_ref.cuh -- the per-pixel softmax of `expAndNormalize` [EXT] with
// the float association of a sequential CPU evaluation, so that the marginals are bit-identical to
// oracle/densecrf_oracle.c (orc_exp_and_normalize):
//     mx = max_l b[l];  o[l] = expf(b[l] - mx);  s = ((o[0] + o[1]) + o[2]) + ...;  Q[l] = o[l] / s
//
//  * expf: the oracle calls the host libm.  glibc (>= 2.27) evaluates expf in double precision --
//    x * 32/ln2 = k + r, 2^(k/32) from a 32-entry table, a cubic in r, one rounding to float at the
//    end -- which is restated here operation for operation.  tests/test_oracle.py checks the same
//    restatement (oracle/expf_ref.c) against the host's expf over every float in [-104, 0]: identical
//    for all but one of 1.12e9 inputs (x = -0x1.f8cbb2p+5, a 1-ulp difference on a value of 1e-28).
//    CUDA's own expf differs from glibc's by an ulp on a fraction of the inputs, and at bistable
//    pixels the mean-field map amplifies such differences ~2.5x per iteration.
//  * the table lives in registers: lane i of every warp holds entry i, a look-up is two shuffles
//    (the slice kernels are LSU-bound; 24 table loads per pixel would cost more than the shuffles).
//  * the sum runs in label order through the lanes of the pixel's group (a chain of shuffles).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace dcrf {

// asuint64(2^(i/32)) - (i << 47), i = 0..31 (2^(i/32) correctly rounded to double)
__device__ const unsigned long long kExp2fTab[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull,
};

struct ExpfRef {
    unsigned lo, hi;  // this lane's table entry
    __device__ __forceinline__ ExpfRef() {
        const unsigned long long t = kExp2fTab[threadIdx.x & 31];
        lo = (unsigned)t;
        hi = (unsigned)(t >> 32);
    }
    // expf(x) for x <= 0 (and -inf).  Warp-collective: all 32 lanes must call it together.
    __device__ __forceinline__ float operator()(float x) const {
        constexpr unsigned FULL = 0xffffffffu;
        constexpr double kInvLn2N = 0x1.71547652b82fep+0 * 32.0;
        constexpr double kShift = 0x1.8p+52;
        constexpr double kC0 = 0x1.c6af84b912394p-5 / 32.0 / 32.0 / 32.0;
        constexpr double kC1 = 0x1.ebfce50fac4f3p-3 / 32.0 / 32.0;
        constexpr double kC2 = 0x1.62e42ff0c52d6p-1 / 32.0;
        const double z = __dmul_rn(kInvLn2N, (double)x);
        const double ks = __dadd_rn(z, kShift);          // round to integer, ties to even
        const unsigned klo = (unsigned)__double2loint(ks);  // low bits = k (two's complement)
        const double r = __dsub_rn(z, __dsub_rn(ks, kShift));
        const unsigned tlo = __shfl_sync(FULL, lo, klo & 31);
        const unsigned thi = __shfl_sync(FULL, hi, klo & 31) + (klo << 15);  // += k << 47
        const double s = __hiloint2double((int)thi, (int)tlo);               // 2^(k/32)
        const double r2 = __dmul_rn(r, r);
        const double p = __fma_rn(kC0, r, kC1);
        double y = __fma_rn(kC2, r, 1.0);
        y = __fma_rn(p, r2, y);
        y = __dmul_rn(y, s);
        const float out = __double2float_rn(y);
        return x < -0x1.9fe368p6f ? 0.f : out;  // underflow (and -inf / garbage lanes) -> +0
    }
};

// Q row = softmax over the L valid labels of `t` (4 labels per lane, G = g lanes per pixel starting at
// lane `gbase`, this lane owning labels 4c..4c+3).  Warp-collective.  Padding labels come out as 0.
template <int G>
__device__ __forceinline__ float4 softmax_ref_row(float4 t, int L, int c, int g_rt, int gbase, const ExpfRef &ex) {
    constexpr unsigned FULL = 0xffffffffu;
    const int g = G ? G : g_rt;
    const int l0 = c * 4;
    const float NEG = -INFINITY;
    if (l0 + 0 >= L) t.x = NEG;
    if (l0 + 1 >= L) t.y = NEG;
    if (l0 + 2 >= L) t.z = NEG;
    if (l0 + 3 >= L) t.w = NEG;
    const float m = fmaxf(fmaxf(t.x, t.y), fmaxf(t.z, t.w));
    float mx = NEG;
    for (int i = 0; i < g; i++) mx = fmaxf(mx, __shfl_sync(FULL, m, (gbase + i) & 31));
    float4 e;
    e.x = ex(__fsub_rn(t.x, mx));  // padding: expf(-inf) = 0
    e.y = ex(__fsub_rn(t.y, mx));
    e.z = ex(__fsub_rn(t.z, mx));
    e.w = ex(__fsub_rn(t.w, mx));
    // s = (((0 + e[0]) + e[1]) + ...) in label order; x + (+0) = x, so the padding does not matter
    float run = 0.f;
    for (int i = 0; i < g; i++) {
        const float prev = __shfl_sync(FULL, run, (gbase + i - 1) & 31);
        if (c == i) {
            const float b = i == 0 ? 0.f : prev;
            run = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(b, e.x), e.y), e.z), e.w);
        }
    }
    const float sum = __shfl_sync(FULL, run, (gbase + g - 1) & 31);
    return make_float4(__fdiv_rn(e.x, sum), __fdiv_rn(e.y, sum), __fdiv_rn(e.z, sum), __fdiv_rn(e.w, sum));
}

}  // namespace dcrf
