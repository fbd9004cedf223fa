// filter.cu -- per-iteration hot loop of the DenseCRF mean field on sm_100a:
//   splat (deterministic gather over the transposed incidence rows), d+1 directional blurs,
//   and slice fused with normalisation, label compatibility, unary add and the softmax over labels.
//
// Replaces `Permutohedral::compute`, `DenseKernel::filter`, `PottsCompatibility::apply` and
// `expAndNormalize` inside pydensecrf's `inference(n)` [EXT], reached
// from /root/reference/03c_hsn/utilities.py:442.
//
// Data layout: every value matrix is "row-major with Lp floats per row", Lp = L rounded up to a
// multiple of 4, pad lanes always 0.  A row is handled by G = Lp/4 adjacent lanes, one float4 each,
// so every global access is a 128-bit access and a warp touches 32/G consecutive rows.
// HBM-bound byte work: no tensor cores.
#include <math.h>

#include "common.cuh"

namespace dcrf {

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

// lane -> (row within warp, float4 column); rows_per_warp = 32 / g
template <int G>
struct RowMap {
    int g, rpw;
    __device__ __forceinline__ explicit RowMap(int g_rt) {
        g = G ? G : g_rt;
        rpw = 32 / g;
    }
    __device__ __forceinline__ int sub() const { return (threadIdx.x & 31) / g; }
    __device__ __forceinline__ int col() const { return (threadIdx.x & 31) % g; }
    __device__ __forceinline__ int64_t row() const {
        return ((int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5)) * rpw + sub();
    }
    __device__ __forceinline__ bool lane_active() const { return sub() < rpw; }
};

static inline int rows_per_block(int g) { return kWarps * (32 / g); }

__device__ __forceinline__ float4 ldg4(const float *p) {
    return __ldg(reinterpret_cast<const float4 *>(p));
}
__device__ __forceinline__ void st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }

// acc += w * q with separately rounded multiply and add (the CPU specification has no FMA)
__device__ __forceinline__ void mul_add(float4 &acc, float w, const float4 q) {
    acc.x = __fadd_rn(acc.x, __fmul_rn(w, q.x));
    acc.y = __fadd_rn(acc.y, __fmul_rn(w, q.y));
    acc.z = __fadd_rn(acc.z, __fmul_rn(w, q.z));
    acc.w = __fadd_rn(acc.w, __fmul_rn(w, q.w));
}
__device__ __forceinline__ float4 scale4(const float4 q, float n) {
    return make_float4(__fmul_rn(q.x, n), __fmul_rn(q.y, n), __fmul_rn(q.z, n), __fmul_rn(q.w, n));
}

// ---------------------------------------------------------------------------------------------
// splat: val[v] = sum over the row's entries, ascending entry order, of w * (norm[p] * Q[p])
// (same summation order as the sequential pixel scan of A.4 => bit-identical lattice values)
// ---------------------------------------------------------------------------------------------
template <int G, bool PRE>
__global__ void __launch_bounds__(kThreads) splat_kernel(
    const int32_t *__restrict__ csr_start, const int32_t *__restrict__ csr_pix,
    const float *__restrict__ csr_w, const float *__restrict__ Q, const float *__restrict__ norm,
    float *__restrict__ val, int64_t M, int g_rt) {
    const RowMap<G> rm(g_rt);
    const int64_t v = rm.row();
    if (!rm.lane_active() || v >= M) return;
    const int g = rm.g, c = rm.col();
    int s = csr_start[v];
    const int s1 = csr_start[v + 1];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (; s + 4 <= s1; s += 4) {
        int p[4];
        float w[4];
        float4 q[4];
        float n[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            p[i] = csr_pix[s + i];
            w[i] = csr_w[s + i];
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            q[i] = ldg4(Q + ((int64_t)p[i] * g + c) * 4);
            if (PRE) n[i] = norm[p[i]];
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (PRE) q[i] = scale4(q[i], n[i]);
            mul_add(acc, w[i], q[i]);
        }
    }
    for (; s < s1; s++) {
        const int p = csr_pix[s];
        const float w = csr_w[s];
        float4 q = ldg4(Q + ((int64_t)p * g + c) * 4);
        if (PRE) q = scale4(q, norm[p]);
        mul_add(acc, w, q);
    }
    st4(val + (v * g + c) * 4, acc);
}

// ---------------------------------------------------------------------------------------------
// blur along one axis: out[v] = in[v] + 0.5 * (in[n1] + in[n2]); absent neighbour = zero row.
// SEQ reproduces the value_size<=2 association (sum in float, 0.5* and outer add in double).
// ---------------------------------------------------------------------------------------------
template <bool SEQ>
__device__ __forceinline__ float blur1(float o, float a, float b) {
    if (SEQ) return (float)((double)o + 0.5 * (double)__fadd_rn(a, b));
    return __fadd_rn(o, __fmul_rn(0.5f, __fadd_rn(a, b)));
}

template <int G, bool SEQ>
__global__ void __launch_bounds__(kThreads) blur_kernel(const int2 *__restrict__ neigh,
                                                        const float *__restrict__ in,
                                                        float *__restrict__ out, int64_t M, int g_rt) {
    const RowMap<G> rm(g_rt);
    const int64_t v = rm.row();
    if (!rm.lane_active() || v >= M) return;
    const int g = rm.g, c = rm.col();
    const int2 nb = neigh[v];
    const float4 o = ldg4(in + (v * g + c) * 4);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (nb.x >= 0) a = ldg4(in + ((int64_t)nb.x * g + c) * 4);
    if (nb.y >= 0) b = ldg4(in + ((int64_t)nb.y * g + c) * 4);
    float4 r;
    r.x = blur1<SEQ>(o.x, a.x, b.x);
    r.y = blur1<SEQ>(o.y, a.y, b.y);
    r.z = blur1<SEQ>(o.z, a.z, b.z);
    r.w = blur1<SEQ>(o.w, a.w, b.w);
    st4(out + (v * g + c) * 4, r);
}

// ---------------------------------------------------------------------------------------------
// slice helpers
// ---------------------------------------------------------------------------------------------
// non-SEQ: acc += (w*alpha) * v ; SEQ: acc += (w * v) * alpha        (A.4 slice)
template <bool SEQ>
__device__ __forceinline__ float4 slice_row(const int32_t *__restrict__ offset,
                                            const float *__restrict__ bary,
                                            const float *__restrict__ val, int64_t p, int d, float alpha,
                                            int g, int c) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int64_t base = p * (d + 1);
    for (int r = 0; r <= d; r++) {
        const int o = offset[base + r];
        const float w = bary[base + r];
        const float4 v = ldg4(val + ((int64_t)o * g + c) * 4);
        if (SEQ) {
            acc.x = __fadd_rn(acc.x, __fmul_rn(__fmul_rn(w, v.x), alpha));
            acc.y = __fadd_rn(acc.y, __fmul_rn(__fmul_rn(w, v.y), alpha));
            acc.z = __fadd_rn(acc.z, __fmul_rn(__fmul_rn(w, v.z), alpha));
            acc.w = __fadd_rn(acc.w, __fmul_rn(__fmul_rn(w, v.w), alpha));
        } else {
            mul_add(acc, __fmul_rn(w, alpha), v);
        }
    }
    return acc;
}

// compat( norm (.) sliced ): Potts -> (-w) * x ; diagonal -> c[l] * x[l] ; matrix -> C x
// (matrix rows are padded to Lp with zeros; all lanes of the warp must call this)
__device__ __forceinline__ float4 apply_compat(const SliceTerm &t, float4 x, int g, int c, int Lp) {
    if (t.compat_kind == DCRF_COMPAT_POTTS) {
        const float w = -t.potts_w;
        return make_float4(__fmul_rn(w, x.x), __fmul_rn(w, x.y), __fmul_rn(w, x.z), __fmul_rn(w, x.w));
    }
    if (t.compat_kind == DCRF_COMPAT_DIAGONAL) {
        const float4 cc = ldg4(t.compat + c * 4);
        return make_float4(__fmul_rn(x.x, cc.x), __fmul_rn(x.y, cc.y), __fmul_rn(x.z, cc.z),
                           __fmul_rn(x.w, cc.w));
    }
    // matrix: out[a] = sum_b C[a][b] x[b], b ascending; x[b] fetched from the row's other lanes
    const int lane = threadIdx.x & 31;
    const int gbase = lane - c;
    float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
    const float *C0 = t.compat + (int64_t)(c * 4) * Lp;
    for (int bl = 0; bl < g; bl++) {
        const int src = (gbase + bl) & 31;
        float xb[4];
        xb[0] = __shfl_sync(0xffffffffu, x.x, src);
        xb[1] = __shfl_sync(0xffffffffu, x.y, src);
        xb[2] = __shfl_sync(0xffffffffu, x.z, src);
        xb[3] = __shfl_sync(0xffffffffu, x.w, src);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int b = bl * 4 + i;
            out.x = __fadd_rn(out.x, __fmul_rn(C0[0 * Lp + b], xb[i]));
            out.y = __fadd_rn(out.y, __fmul_rn(C0[1 * Lp + b], xb[i]));
            out.z = __fadd_rn(out.z, __fmul_rn(C0[2 * Lp + b], xb[i]));
            out.w = __fadd_rn(out.w, __fmul_rn(C0[3 * Lp + b], xb[i]));
        }
    }
    return out;
}

// ---------------------------------------------------------------------------------------------
// fused: slice every pairwise term, normalise, compat, unary add, softmax over labels  (A.7)
//   t = -U ; for k: t -= compat_k( norm_k * slice_k ) ; Q = softmax_L(t)
// ---------------------------------------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(kThreads) slice_softmax_kernel(const SliceArgs a,
                                                                 const float *__restrict__ unary,
                                                                 float *__restrict__ Q, int64_t Ntot,
                                                                 int L, int g_rt) {
    const RowMap<G> rm(g_rt);
    const int g = rm.g, c = rm.col();
    const int64_t p = rm.row();
    const bool act = rm.lane_active() && p < Ntot;
    const int64_t pc = act ? p : 0;  // inactive lanes shadow pixel 0 so that shuffles stay uniform
    const float4 u = ldg4(unary + (pc * g + c) * 4);
    float4 t = make_float4(-u.x, -u.y, -u.z, -u.w);
    const int Lp = g * 4;
    for (int k = 0; k < a.n_terms; k++) {
        const SliceTerm &tm = a.term[k];
        float4 x = a.seq ? slice_row<true>(tm.offset, tm.bary, tm.val, pc, tm.d, tm.alpha, g, c)
                         : slice_row<false>(tm.offset, tm.bary, tm.val, pc, tm.d, tm.alpha, g, c);
        if (tm.norm) x = scale4(x, tm.norm[pc]);
        const float4 y = apply_compat(tm, x, g, c, Lp);
        t.x = __fsub_rn(t.x, y.x);
        t.y = __fsub_rn(t.y, y.y);
        t.z = __fsub_rn(t.z, y.z);
        t.w = __fsub_rn(t.w, y.w);
    }
    // softmax over the L valid labels of the row (max-subtracted)
    const int l0 = c * 4;
    const float NEG = -INFINITY;
    float m = NEG;
    if (l0 + 0 < L) m = fmaxf(m, t.x);
    if (l0 + 1 < L) m = fmaxf(m, t.y);
    if (l0 + 2 < L) m = fmaxf(m, t.z);
    if (l0 + 3 < L) m = fmaxf(m, t.w);
    const int lane = threadIdx.x & 31;
    const int gbase = lane - c;
    float mx = NEG;
    for (int i = 0; i < g; i++) mx = fmaxf(mx, __shfl_sync(0xffffffffu, m, (gbase + i) & 31));
    float4 e;
    e.x = (l0 + 0 < L) ? expf(__fsub_rn(t.x, mx)) : 0.f;
    e.y = (l0 + 1 < L) ? expf(__fsub_rn(t.y, mx)) : 0.f;
    e.z = (l0 + 2 < L) ? expf(__fsub_rn(t.z, mx)) : 0.f;
    e.w = (l0 + 3 < L) ? expf(__fsub_rn(t.w, mx)) : 0.f;
    const float ls = __fadd_rn(__fadd_rn(__fadd_rn(e.x, e.y), e.z), e.w);
    float sum = 0.f;
    for (int i = 0; i < g; i++) sum = __fadd_rn(sum, __shfl_sync(0xffffffffu, ls, (gbase + i) & 31));
    if (act) {
        float4 q;
        q.x = __fdiv_rn(e.x, sum);
        q.y = __fdiv_rn(e.y, sum);
        q.z = __fdiv_rn(e.z, sum);
        q.w = __fdiv_rn(e.w, sum);
        st4(Q + (p * g + c) * 4, q);
    }
}

// slice of one lattice without any epilogue (norm construction, test hook)
template <int G, bool SEQ>
__global__ void __launch_bounds__(kThreads) slice_plain_kernel(
    const int32_t *__restrict__ offset, const float *__restrict__ bary, const float *__restrict__ val,
    float *__restrict__ out, int64_t Ntot, int d, float alpha, int g_rt) {
    const RowMap<G> rm(g_rt);
    const int64_t p = rm.row();
    if (!rm.lane_active() || p >= Ntot) return;
    const int g = rm.g, c = rm.col();
    const float4 x = slice_row<SEQ>(offset, bary, val, p, d, alpha, g, c);
    st4(out + (p * g + c) * 4, x);
}

// pairwise_out = compat( norm * slice ) for a single term (klDivergence)
template <int G>
__global__ void __launch_bounds__(kThreads) slice_pairwise_kernel(const SliceTerm tm,
                                                                  float *__restrict__ out,
                                                                  int64_t Ntot, int g_rt, int seq) {
    const RowMap<G> rm(g_rt);
    const int g = rm.g, c = rm.col();
    const int64_t p = rm.row();
    const bool act = rm.lane_active() && p < Ntot;
    const int64_t pc = act ? p : 0;
    float4 x = seq ? slice_row<true>(tm.offset, tm.bary, tm.val, pc, tm.d, tm.alpha, g, c)
                   : slice_row<false>(tm.offset, tm.bary, tm.val, pc, tm.d, tm.alpha, g, c);
    if (tm.norm) x = scale4(x, tm.norm[pc]);
    const float4 y = apply_compat(tm, x, g, c, g * 4);
    if (act) st4(out + (p * g + c) * 4, y);
}

// norm[p] from the sliced all-ones filter (column 0 of an Lp-wide buffer)      (A.5)
__global__ void __launch_bounds__(kThreads) norm_finalize_kernel(const float *__restrict__ sliced,
                                                                 int Lp, float *__restrict__ norm,
                                                                 int64_t Ntot, int ntype) {
    const int64_t p = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (p >= Ntot) return;
    const float x = sliced[p * Lp];
    float r;
    if (ntype == DCRF_NORMALIZE_SYMMETRIC) r = (float)(1.0 / sqrt((double)x + 1e-20));
    else r = (float)(1.0 / ((double)x + 1e-20));
    norm[p] = r;
}

__global__ void __launch_bounds__(kThreads) fill_ones_col0_kernel(float *__restrict__ buf, int64_t n4) {
    // buf viewed as float4 rows of width Lp = 4: (1, 0, 0, 0)
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i < n4) st4(buf + i * 4, make_float4(1.f, 0.f, 0.f, 0.f));
}

// ---------------------------------------------------------------------------------------------
// layout changes at the API boundary: (L, N_b) row-major blocks <-> (Ntot, Lp) pixel-major
// ---------------------------------------------------------------------------------------------
constexpr int kTP = 32;  // pixels per tile
// grid (ceil(maxN/32), B), block (32, 8)
__global__ void ln_to_pm_kernel(const float *__restrict__ ln, float *__restrict__ pm,
                                const int *__restrict__ pix_start, int L, int Lp) {
    extern __shared__ float tile[];  // [Lp][33]
    const int b = blockIdx.y;
    const int64_t ps = pix_start[b];
    const int Nb = (int)(pix_start[b + 1] - ps);
    const int p0 = blockIdx.x * kTP;
    if (p0 >= Nb) return;
    const float *src = ln + ps * L;  // image block (L, Nb)
    for (int l = threadIdx.y; l < Lp; l += blockDim.y) {
        const int p = p0 + threadIdx.x;
        tile[l * (kTP + 1) + threadIdx.x] = (l < L && p < Nb) ? src[(int64_t)l * Nb + p] : 0.f;
    }
    __syncthreads();
    const int np = min(kTP, Nb - p0);
    float *dst = pm + (ps + p0) * Lp;
    for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < np * Lp; i += blockDim.x * blockDim.y) {
        const int pp = i / Lp, l = i - pp * Lp;
        dst[i] = tile[l * (kTP + 1) + pp];
    }
}

__global__ void pm_to_ln_kernel(const float *__restrict__ pm, float *__restrict__ ln,
                                const int *__restrict__ pix_start, int L, int Lp) {
    extern __shared__ float tile[];  // [Lp][33]
    const int b = blockIdx.y;
    const int64_t ps = pix_start[b];
    const int Nb = (int)(pix_start[b + 1] - ps);
    const int p0 = blockIdx.x * kTP;
    if (p0 >= Nb) return;
    const int np = min(kTP, Nb - p0);
    const float *src = pm + (ps + p0) * Lp;
    for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < np * Lp; i += blockDim.x * blockDim.y) {
        const int pp = i / Lp, l = i - pp * Lp;
        tile[l * (kTP + 1) + pp] = src[i];
    }
    __syncthreads();
    float *dst = ln + ps * L;
    for (int l = threadIdx.y; l < L; l += blockDim.y) {
        const int p = p0 + threadIdx.x;
        if (p < Nb) dst[(int64_t)l * Nb + p] = tile[l * (kTP + 1) + threadIdx.x];
    }
}

// first maximum wins, like np.argmax
__global__ void __launch_bounds__(kThreads) argmax_kernel(const float *__restrict__ pm,
                                                          int32_t *__restrict__ labels, int64_t Ntot,
                                                          int L, int Lp) {
    const int64_t p = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (p >= Ntot) return;
    const float *row = pm + p * Lp;
    float best = row[0];
    int bi = 0;
    for (int l4 = 0; l4 < Lp; l4 += 4) {
        const float4 v = ldg4(row + l4);
        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int l = l4 + i;
            if (l < L && vv[i] > best) { best = vv[i]; bi = l; }
        }
    }
    labels[p] = bi;
}

// ---------------------------------------------------------------------------------------------
// KL divergence terms: fixed-shape double reduction (deterministic)
// ---------------------------------------------------------------------------------------------
constexpr int kKlBlocks = 1024;
__global__ void __launch_bounds__(kThreads) kl_partial_kernel(
    const float *__restrict__ Q, const float *__restrict__ unary, const float *__restrict__ p0,
    const float *__restrict__ p1, const float *__restrict__ p2, const float *__restrict__ p3, int n_pair,
    int64_t Ntot, int L, int Lp, double *__restrict__ partial) {
    __shared__ double sh[kThreads];
    const float *pp[4] = {p0, p1, p2, p3};
    double acc = 0.0;
    const int64_t total = Ntot * Lp;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total;
         i += (int64_t)kKlBlocks * kThreads) {
        const int l = (int)(i % Lp);
        if (l >= L) continue;
        const float q = Q[i];
        const float qc = q > 1e-20f ? q : 1e-20f;
        acc += (double)q * log((double)qc);
        acc += (double)unary[i] * (double)q;
        for (int k = 0; k < n_pair; k++) acc += (double)__fmul_rn(q, pp[k][i]);
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = kThreads / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
__global__ void kl_final_kernel(const double *__restrict__ partial, double *__restrict__ out) {
    __shared__ double sh[kKlBlocks];
    for (int i = threadIdx.x; i < kKlBlocks; i += blockDim.x) sh[i] = partial[i];
    __syncthreads();
    for (int s = kKlBlocks / 2; s > 0; s >>= 1) {
        for (int i = threadIdx.x; i < s; i += blockDim.x) sh[i] += sh[i + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sh[0];
}

// dispatch on G = Lp/4 (1..8 specialised, anything else through the runtime-g instantiation)
#define DCRF_DISPATCH_G(g, ...)                                  \
    switch (g) {                                                 \
        case 1: { constexpr int G = 1; __VA_ARGS__; } break;     \
        case 2: { constexpr int G = 2; __VA_ARGS__; } break;     \
        case 3: { constexpr int G = 3; __VA_ARGS__; } break;     \
        case 4: { constexpr int G = 4; __VA_ARGS__; } break;     \
        case 5: { constexpr int G = 5; __VA_ARGS__; } break;     \
        case 6: { constexpr int G = 6; __VA_ARGS__; } break;     \
        case 7: { constexpr int G = 7; __VA_ARGS__; } break;     \
        case 8: { constexpr int G = 8; __VA_ARGS__; } break;     \
        default: { constexpr int G = 0; __VA_ARGS__; } break;    \
    }

}  // namespace

void launch_splat(const Lattice &lat, const float *Q, const float *norm_pre, float *val, int Lp,
                  cudaStream_t s) {
    if (lat.M == 0) return;
    const int g = Lp / 4;
    const int nb = ceil_div(lat.M, rows_per_block(g));
    ProfScope prof(DCRF_K_SPLAT, lat.d, s);
    DCRF_DISPATCH_G(g, {
        if (norm_pre)
            splat_kernel<G, true><<<nb, kThreads, 0, s>>>(lat.csr_start.p, lat.csr_pix.p, lat.csr_w.p, Q,
                                                         norm_pre, val, lat.M, g);
        else
            splat_kernel<G, false><<<nb, kThreads, 0, s>>>(lat.csr_start.p, lat.csr_pix.p, lat.csr_w.p,
                                                          Q, nullptr, val, lat.M, g);
    });
    DCRF_LAUNCHED();
}

void launch_blur(const Lattice &lat, int axis, const float *in, float *out, int Lp, bool seq,
                 cudaStream_t s) {
    if (lat.M == 0) return;
    const int g = Lp / 4;
    const int nb = ceil_div(lat.M, rows_per_block(g));
    const int2 *nbr = lat.neigh.p + (int64_t)axis * lat.M;
    ProfScope prof(DCRF_K_BLUR, lat.d, s);
    DCRF_DISPATCH_G(g, {
        if (seq) blur_kernel<G, true><<<nb, kThreads, 0, s>>>(nbr, in, out, lat.M, g);
        else blur_kernel<G, false><<<nb, kThreads, 0, s>>>(nbr, in, out, lat.M, g);
    });
    DCRF_LAUNCHED();
}

void launch_slice_softmax(const SliceArgs &a, const float *unary, float *Q, int64_t Ntot, int L, int Lp,
                          cudaStream_t s) {
    if (Ntot == 0) return;
    const int g = Lp / 4;
    const int nb = ceil_div(Ntot, rows_per_block(g));
    ProfScope prof(DCRF_K_SLICE, a.n_terms, s);
    DCRF_DISPATCH_G(g, { slice_softmax_kernel<G><<<nb, kThreads, 0, s>>>(a, unary, Q, Ntot, L, g); });
    DCRF_LAUNCHED();
}

void launch_slice_plain(const Lattice &lat, const float *val, float *out, int64_t Ntot, int Lp, bool seq,
                        cudaStream_t s) {
    if (Ntot == 0) return;
    const int g = Lp / 4;
    const int nb = ceil_div(Ntot, rows_per_block(g));
    const float alpha = 1.0f / (1.0f + powf(2.0f, (float)-lat.d));
    DCRF_DISPATCH_G(g, {
        if (seq)
            slice_plain_kernel<G, true><<<nb, kThreads, 0, s>>>(lat.offset.p, lat.bary.p, val, out, Ntot,
                                                               lat.d, alpha, g);
        else
            slice_plain_kernel<G, false><<<nb, kThreads, 0, s>>>(lat.offset.p, lat.bary.p, val, out, Ntot,
                                                                lat.d, alpha, g);
    });
    DCRF_LAUNCHED();
}

void launch_slice_pairwise_only(const SliceTerm &t, float *out, int64_t Ntot, int L, int Lp,
                                cudaStream_t s) {
    if (Ntot == 0) return;
    const int g = Lp / 4;
    const int nb = ceil_div(Ntot, rows_per_block(g));
    DCRF_DISPATCH_G(g, { slice_pairwise_kernel<G><<<nb, kThreads, 0, s>>>(t, out, Ntot, g, L <= 2 ? 1 : 0); });
    DCRF_LAUNCHED();
}

void launch_norm_finalize(const float *sliced, int Lp, float *norm, int64_t Ntot, int ntype,
                          cudaStream_t s) {
    if (Ntot == 0) return;
    norm_finalize_kernel<<<ceil_div(Ntot, kThreads), kThreads, 0, s>>>(sliced, Lp, norm, Ntot, ntype);
    DCRF_LAUNCHED();
}

void launch_fill_ones_col0(float *buf, int64_t Ntot, int Lp, cudaStream_t s) {
    DCRF_REQUIRE(Lp == 4, DCRF_EINVAL, "fill_ones_col0 expects Lp == 4");
    if (Ntot == 0) return;
    fill_ones_col0_kernel<<<ceil_div(Ntot, kThreads), kThreads, 0, s>>>(buf, Ntot);
    DCRF_LAUNCHED();
}

static int max_image_pixels(const BatchGeom &g) {
    int64_t m = 0;
    for (int b = 0; b < g.B; b++) m = std::max<int64_t>(m, g.pix_start[b + 1] - g.pix_start[b]);
    return (int)m;
}

void launch_ln_to_pm(const float *ln, float *pm, const BatchGeom &g, int L, int Lp, cudaStream_t s) {
    if (g.Ntot == 0) return;
    dim3 grid(ceil_div(max_image_pixels(g), kTP), g.B), block(kTP, 8);
    ln_to_pm_kernel<<<grid, block, sizeof(float) * Lp * (kTP + 1), s>>>(ln, pm, g.d_pix_start, L, Lp);
    DCRF_LAUNCHED();
}

void launch_pm_to_ln(const float *pm, float *ln, const BatchGeom &g, int L, int Lp, cudaStream_t s) {
    if (g.Ntot == 0) return;
    dim3 grid(ceil_div(max_image_pixels(g), kTP), g.B), block(kTP, 8);
    pm_to_ln_kernel<<<grid, block, sizeof(float) * Lp * (kTP + 1), s>>>(pm, ln, g.d_pix_start, L, Lp);
    DCRF_LAUNCHED();
}

void launch_argmax(const float *pm, int32_t *labels, int64_t Ntot, int L, int Lp, cudaStream_t s) {
    if (Ntot == 0) return;
    argmax_kernel<<<ceil_div(Ntot, kThreads), kThreads, 0, s>>>(pm, labels, Ntot, L, Lp);
    DCRF_LAUNCHED();
}

void launch_kl(const float *Q, const float *unary, const float *const *pair_out, int n_pair, int64_t Ntot,
               int L, int Lp, double *out, cudaStream_t s) {
    DevBuf<double> partial;
    partial.alloc(kKlBlocks, s);
    const float *p[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int k = 0; k < n_pair && k < 4; k++) p[k] = pair_out[k];
    kl_partial_kernel<<<kKlBlocks, kThreads, 0, s>>>(Q, unary, p[0], p[1], p[2], p[3], n_pair, Ntot, L, Lp,
                                                    partial.p);
    DCRF_LAUNCHED();
    kl_final_kernel<<<1, 256, 0, s>>>(partial.p, out);
    DCRF_LAUNCHED();
}

}  // namespace dcrf
