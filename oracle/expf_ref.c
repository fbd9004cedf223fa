/* expf_ref.c -- TEST INFRASTRUCTURE (see densecrf_oracle.c's header).
 *
 * Restatement of the double-precision expf algorithm of glibc >= 2.27 (sysdeps/ieee754/flt-32/
 * e_expf.c, from the ARM optimized-routines): x * 32/ln2 = k + r, 2^(k/32) from a 32-entry table, a
 * cubic in r, one final rounding to float.  The CUDA softmax of the reference-arithmetic mode
 * (wsss_analysis_b200/csrc/softmax_ref.cuh) evaluates exactly these operations; this file exists so
 * that the restatement can be checked against the host's own expf (which the oracle's softmax
 * calls) on the CPU, for every float in the softmax's input range [-104, 0].
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static const uint64_t kTab[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull,
};

/* expf(x) for x <= 0 */
float orc_expf_ref(float x) {
    const double inv_ln2_n = 0x1.71547652b82fep+0 * 32.0, shift = 0x1.8p+52;
    const double c0 = 0x1.c6af84b912394p-5 / 32.0 / 32.0 / 32.0, c1 = 0x1.ebfce50fac4f3p-3 / 32.0 / 32.0,
                 c2 = 0x1.62e42ff0c52d6p-1 / 32.0;
    if (x < -0x1.9fe368p6f) return 0.0f;
    double z = inv_ln2_n * (double)x;
    double kd = z + shift;
    uint64_t ki, t;
    memcpy(&ki, &kd, 8);
    kd -= shift;
    double r = z - kd;
    t = kTab[ki % 32] + (ki << 47);
    double s;
    memcpy(&s, &t, 8);
    double r2 = r * r;
    double p = fma(c0, r, c1);
    double y = fma(c2, r, 1.0);
    y = fma(p, r2, y);
    y = y * s;
    return (float)y;
}

/* compare with the host libm over the float bit patterns u_lo, u_lo + stride, ... <= u_hi;
 * returns the number of mismatches and the first mismatching pattern */
int64_t orc_expf_ref_check(uint32_t u_lo, uint32_t u_hi, uint32_t stride, uint32_t *first_bad) {
    int64_t bad = 0;
    for (uint64_t u = u_lo; u <= u_hi; u += stride) {
        uint32_t b = (uint32_t)u;
        float x, a, e;
        memcpy(&x, &b, 4);
        a = orc_expf_ref(x);
        e = expf(x);
        if (memcmp(&a, &e, 4)) {
            if (!bad && first_bad) *first_bad = b;
            bad++;
        }
    }
    return bad;
}

/* y[i] = host libm expf(x[i]) -- what the oracle's softmax calls (numpy's own exp is a different routine) */
void orc_expf_host(const float *x, float *y, int64_t n) {
    for (int64_t i = 0; i < n; i++) y[i] = expf(x[i]);
}
