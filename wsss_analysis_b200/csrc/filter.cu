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
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"
#include "softmax_ref.cuh"

namespace dcrf {

namespace {

#ifndef DCRF_TUNE_THREADS
#define DCRF_TUNE_THREADS 256
#endif
constexpr int kThreads = DCRF_TUNE_THREADS;
constexpr int kWarps = kThreads / 32;

// lane -> (row within warp, float4 column); rows_per_warp = 32 / g
template <int G>
struct RowMap {
    int g, rpw, vb;
    // vblock: the block index this CTA plays (persistent kernels loop over virtual blocks)
    __device__ __forceinline__ explicit RowMap(int g_rt, int vblock = blockIdx.x) {
        g = G ? G : g_rt;
        rpw = 32 / g;
        vb = vblock;
    }
    __device__ __forceinline__ int sub() const { return (threadIdx.x & 31) / g; }
    __device__ __forceinline__ int col() const { return (threadIdx.x & 31) % g; }
    __device__ __forceinline__ int64_t row() const {
        return ((int64_t)vb * kWarps + (threadIdx.x >> 5)) * rpw + sub();
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
// (same summation order as the sequential pixel scan of A.4 => bit-identical lattice values).
//
// Flat mapping: thread t owns float4 column c = t % g of the rows v = t / g, t/g + n_groups, ...
// (persistent loop).  Rows differ a lot in length (1 .. thousands of entries), so the loop is
// flattened: every trip handles one batch of <= SB entries of the thread's CURRENT row and moves on
// to its next row when the row is finished.  Lanes of a warp therefore never wait for the longest
// row of the warp, and every trip keeps SB independent gathers in flight.
// ---------------------------------------------------------------------------------------------
constexpr int kSplatBatch = 4;
constexpr int kSplatBlocksPerSM = 6;

template <int G, bool PRE>
__global__ void __launch_bounds__(kThreads) splat_kernel(
    const int32_t *__restrict__ csr_start, const int32_t *__restrict__ csr_pix,
    const float *__restrict__ csr_w, const float *__restrict__ Q, const float *__restrict__ norm,
    float *__restrict__ val, int64_t M, int g_rt) {
    constexpr int SB = kSplatBatch;
    const int g = G ? G : g_rt;
    const int64_t tid = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int64_t n_groups = ((int64_t)gridDim.x * kThreads) / g;
    int64_t v = tid / g;
    const int c = (int)(tid - v * g);
    if (v >= n_groups || v >= M) return;
    int s = csr_start[v], s1 = csr_start[v + 1];
    // row bounds of the NEXT row are fetched one row ahead so that a row switch costs no extra trip
    int64_t vn = v + n_groups;
    int ns = 0, ns1 = 0;
    if (vn < M) { ns = csr_start[vn]; ns1 = csr_start[vn + 1]; }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (;;) {
        int p[SB];
        float w[SB];
        float4 q[SB];
        float n[SB];
        const int cnt = min(SB, s1 - s);
#pragma unroll
        for (int i = 0; i < SB; i++) {
            const bool ok = i < cnt;
            p[i] = ok ? csr_pix[s + i] : 0;
            w[i] = ok ? csr_w[s + i] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < SB; i++) {
            if (i < cnt) {
                q[i] = ldg4(Q + ((int64_t)p[i] * g + c) * 4);
                if (PRE) n[i] = norm[p[i]];
            }
        }
#pragma unroll
        for (int i = 0; i < SB; i++) {
            if (i < cnt) {
                if (PRE) q[i] = scale4(q[i], n[i]);
                mul_add(acc, w[i], q[i]);
            }
        }
        s += cnt;
        if (s >= s1) {
            st4(val + (v * g + c) * 4, acc);
            acc = make_float4(0.f, 0.f, 0.f, 0.f);
            v = vn;
            if (v >= M) break;
            s = ns;
            s1 = ns1;
            vn = v + n_groups;
            if (vn < M) { ns = csr_start[vn]; ns1 = csr_start[vn + 1]; }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Fast path (default): same algorithm, leaner instruction stream.  The exact kernels above spend
// most of their issue slots on separately rounded multiplies/adds, 64-bit index math, per-entry
// norm gathers and libm-accurate exp/div; ncu showed them issue-bound (IPC 0.6/SMSP at < 30 % of HBM
// peak).  Here: (pixel, weight) and (vertex, weight) pairs are packed into one 64-bit load, the
// pre-normalisation is folded into the splat weight at build time, accumulation uses FMA, indices
// are 32-bit.  Results differ from the exact path by float rounding only (~1e-7 relative on the
// lattice values); run-to-run determinism is unchanged (fixed summation order, no atomics).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) pack_fast_tables_kernel(
    const int32_t *__restrict__ offset, const float *__restrict__ bary, const int32_t *__restrict__ csr_pix,
    const float *__restrict__ csr_w, const float *__restrict__ norm_pre, const float *__restrict__ norm_post,
    int d1, int2 *__restrict__ ent, int2 *__restrict__ csr_ent, int64_t E) {
    const int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (e >= E) return;
    float b = bary[e];
    if (norm_post) b = __fmul_rn(b, norm_post[e / d1]);  // post-normalisation folded into the slice weight
    ent[e] = make_int2(offset[e], __float_as_int(b));
    const int p = csr_pix[e];
    float w = csr_w[e];
    if (norm_pre) w = __fmul_rn(w, norm_pre[p]);
    csr_ent[e] = make_int2(p, __float_as_int(w));
}

// Reference-association tables: nothing is folded that the sequential evaluation rounds separately.
//   ent      = (vertex id, bary * alpha)            -- A.4 slice: wa = w * alpha; acc += wa * v
//   csr_ent4 = (pixel, bary, pre-norm[pixel] | 1, 0) -- A.5 / A.4 splat: val += w * (norm * Q)
// (multiplying by 1.0f is exact, so kernels without a pre-normalisation use the same code path)
__global__ void __launch_bounds__(kThreads) pack_ref_tables_kernel(
    const int32_t *__restrict__ offset, const float *__restrict__ bary, const int32_t *__restrict__ csr_pix,
    const float *__restrict__ csr_w, const float *__restrict__ norm_pre, float alpha, int2 *__restrict__ ent,
    int4 *__restrict__ csr_ent4, int64_t E) {
    const int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (e >= E) return;
    ent[e] = make_int2(offset[e], __float_as_int(__fmul_rn(bary[e], alpha)));
    const int p = csr_pix[e];
    const float n = norm_pre ? norm_pre[p] : 1.0f;
    csr_ent4[e] = make_int4(p, __float_as_int(csr_w[e]), __float_as_int(n), 0);
}

// packed CSR entry of the splat: FMA tables (pixel, w * norm) or reference tables (pixel, w, norm, -)
template <bool REF> struct CsrEnt { typedef int2 type; };
template <> struct CsrEnt<true> { typedef int4 type; };
__device__ __forceinline__ int2 zero_ent(int2) { return make_int2(0, 0); }
__device__ __forceinline__ int4 zero_ent(int4) { return make_int4(0, 0, 0, 0); }

__device__ __forceinline__ void fma4(float4 &acc, float w, const float4 q) {
    acc.x = fmaf(w, q.x, acc.x);
    acc.y = fmaf(w, q.y, acc.y);
    acc.z = fmaf(w, q.z, acc.z);
    acc.w = fmaf(w, q.w, acc.w);
}

template <bool REF>
__device__ __forceinline__ void splat_acc(float4 &acc, float w, float n, const float4 q) {
    if (REF) mul_add(acc, w, scale4(q, n));  // val += w * (norm * Q), every product rounded (A.4, A.5)
    else fma4(acc, w, q);
}
__device__ __forceinline__ float ent_norm(const int2) { return 1.0f; }
__device__ __forceinline__ float ent_norm(const int4 e) { return __int_as_float(e.z); }

// Work distribution: rows differ wildly in length (bilateral lattice: median 6, p99 46, max > 150
// entries), so rows are handed out dynamically.  A warp claims chunks of 32 consecutive rows with
// one atomicAdd, keeps the chunk's 33 row bounds in registers (one coalesced load), and its
// 32/g lane groups pull the next unclaimed row of the chunk whenever they finish one.  Every trip
// of the (warp-uniform) loop handles one batch of <= SB entries of each group's current row.  Each
// row is still summed by ONE group in ascending entry order, so the result does not depend on the
// schedule: bit-deterministic run to run.
constexpr int kSplatChunk = 32;
// Rows longer than this (flat image regions collapse the bilateral lattice to a few vertices with
// thousands of entries each) leave the queue: whole warps sum them, one row per warp (splat_long_rows_body).
constexpr int kSplatLongRow = 256;
constexpr int kSmallLatticeEntries = 4000000;  // below this many entries a lattice counts as "small" (see launch_pack_fast_tables)
#ifndef DCRF_TUNE_SPLAT_REF_MINB
#define DCRF_TUNE_SPLAT_REF_MINB 0  // 0: same __launch_bounds__ minimum as the FMA kernel of that G
#endif
constexpr int kSplatRefMinBlocks = DCRF_TUNE_SPLAT_REF_MINB;

template <int G, int SB, bool REF>
__device__ __forceinline__ void splat_fast_body(const int32_t *__restrict__ csr_start,
                                                const typename CsrEnt<REF>::type *__restrict__ csr_ent,
                                                const float4 *__restrict__ Q4, float4 *__restrict__ val4, int M,
                                                int g_rt, int *__restrict__ row_counter, int long_cap,
                                                int chunk = kSplatChunk) {
    typedef typename CsrEnt<REF>::type Ent;
    constexpr unsigned FULL = 0xffffffffu;
    const int g = G ? G : g_rt;
    const int lane = threadIdx.x & 31;
    const int gpw = 32 / g;             // lane groups per warp
    const int sub = lane / g;           // my group
    const int c = lane - sub * g;       // my float4 column
    const bool lane_on = sub < gpw;
    const unsigned leader_bit = 1u << (sub * g);
    // warp-uniform queue state
    int q_base = 0, q_next = 0, q_end = 0;  // rows [q_next, q_end) of the chunk starting at q_base
    int bounds = 0, bound_last = 0;         // lane i holds csr_start[q_base + i]; bound_last = csr_start[q_base + 32]
    bool exhausted = false;
    // per-group state: current row v with entries [s, s1); e_cur = its current batch, already loaded
    int v = -1, s = 0, s1 = 0, cnt_cur = 0;
    Ent e_cur[SB];
#pragma unroll
    for (int i = 0; i < SB; i++) e_cur[i] = zero_ent(Ent());
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    // Software pipeline: in every trip the row gathers of batch k and the entry loads of batch k+1
    // (possibly of the group's NEXT row, claimed one trip ahead) are in flight together.
    for (;;) {
        // 1. gathers of the current batch
        float4 q[SB];
#pragma unroll
        for (int i = 0; i < SB; i++) {
            q[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < cnt_cur) q[i] = __ldg(Q4 + ((unsigned)e_cur[i].x * g + c));
        }
        // 2. choose the next batch: same row, or claim a new row from the warp's chunk
        const bool row_ends = v < 0 || s + SB >= s1;
        int nv = v, ns = s + SB, ns1 = s1;
        const bool need = lane_on && row_ends;
        const unsigned need_mask = __ballot_sync(FULL, need && c == 0);
        if (need_mask) {
            if (q_next >= q_end && !exhausted) {
                // one global atomic per 32 rows; measured alternatives that were slower on B200
                // (batch of 32 VOC images, d = 5): static striding of rows 842 us, CTA-local chunk
                // dealing 519 us, rows pre-sorted by length 597-943 us; this queue 456 us.
                int base = 0;
                if (lane == 0) base = atomicAdd(row_counter, chunk);
                base = __shfl_sync(FULL, base, 0);
                if (base >= M) {
                    exhausted = true;
                } else {
                    q_base = q_next = base;
                    q_end = min(base + chunk, M);
                    bounds = csr_start[min(base + lane, M)];
                    bound_last = csr_start[q_end];
                }
            }
            const int rank = __popc(need_mask & (leader_bit - 1));
            const int row = q_next + rank;
            const bool take = need && row < q_end;
            const int rel = take ? row - q_base : 0;
            const int b0 = __shfl_sync(FULL, bounds, rel & 31);
            int b1 = __shfl_sync(FULL, bounds, (rel + 1) & 31);
            if (rel + 1 == q_end - q_base) b1 = bound_last;
            if (need) {
                nv = (take && b1 - b0 <= long_cap) ? row : -1;  // long rows: summed by whole warps (splat_long_rows_body)
                ns = b0;
                ns1 = b1;
            }
            q_next = min(q_end, q_next + __popc(need_mask));
        }
        // 3. entry loads of the next batch
        const int ncnt = nv >= 0 ? min(SB, ns1 - ns) : 0;
        Ent e_next[SB];
#pragma unroll
        for (int i = 0; i < SB; i++) e_next[i] = (i < ncnt) ? __ldg(csr_ent + ns + i) : zero_ent(Ent());
        // 4. consume the current batch (padded slots add 0 * 0)
#pragma unroll
        for (int i = 0; i < SB; i++) splat_acc<REF>(acc, __int_as_float(e_cur[i].y), ent_norm(e_cur[i]), q[i]);
        if (v >= 0 && row_ends) {
            val4[(unsigned)v * g + c] = acc;
            acc = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // 5. rotate
        v = nv;
        s = ns;
        s1 = ns1;
        cnt_cur = ncnt;
#pragma unroll
        for (int i = 0; i < SB; i++) e_cur[i] = e_next[i];
        if (!__ballot_sync(FULL, v >= 0) && exhausted) break;
    }
}

// Same schedule with COOPERATIVE entry loads: every warp-level load request costs the LSU a fixed
// >= 6 cycles however few bytes it moves (tools/micro/bulk_gather.cu), and above each lane loads all
// SB entries of its group's batch itself -- SB broadcast requests per SB gathers.  Here lane c of a
// group loads entry c of the batch (one request per G entries) and the (pixel, weight) pairs are
// broadcast inside the group with shuffles.  A trip handles NB * G entries.  Summation order and
// arithmetic are unchanged, so the result is bit-identical to splat_fast_kernel.
template <int G, int NB, bool REF>
__device__ __forceinline__ void splat_coop_body(const int32_t *__restrict__ csr_start,
                                                const typename CsrEnt<REF>::type *__restrict__ csr_ent,
                                                const float4 *__restrict__ Q4, float4 *__restrict__ val4, int M,
                                                int *__restrict__ row_counter, int long_cap,
                                                int chunk = kSplatChunk) {
    typedef typename CsrEnt<REF>::type Ent;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int SB = G * NB;
    constexpr int gpw = 32 / G;
    const int lane = threadIdx.x & 31;
    const int sub = lane / G;
    const int c = lane - sub * G;
    const bool lane_on = sub < gpw;
    const int gbase = (sub * G) & 31;
    const unsigned leader_bit = 1u << gbase;
    int q_base = 0, q_next = 0, q_end = 0;
    int bounds = 0, bound_last = 0;
    bool exhausted = false;
    int v = -1, s = 0, s1 = 0, cnt_cur = 0;
    Ent e_cur[NB];
#pragma unroll
    for (int b = 0; b < NB; b++) e_cur[b] = zero_ent(Ent());
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (;;) {
        // 1. broadcast the batch's entries inside the group, request the row gathers
        float4 q[SB];
#pragma unroll
        for (int b = 0; b < NB; b++)
#pragma unroll
            for (int i = 0; i < G; i++) {
                const int k = b * G + i;
                const int px = __shfl_sync(FULL, e_cur[b].x, (gbase + i) & 31);
                q[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (k < cnt_cur) q[k] = __ldg(Q4 + ((unsigned)px * G + c));
            }
        // 2. next batch: same row, or a new row from the warp's chunk
        const bool row_ends = v < 0 || s + SB >= s1;
        int nv = v, ns = s + SB, ns1 = s1;
        const bool need = lane_on && row_ends;
        const unsigned need_mask = __ballot_sync(FULL, need && c == 0);
        if (need_mask) {
            if (q_next >= q_end && !exhausted) {
                int base = 0;
                if (lane == 0) base = atomicAdd(row_counter, chunk);
                base = __shfl_sync(FULL, base, 0);
                if (base >= M) {
                    exhausted = true;
                } else {
                    q_base = q_next = base;
                    q_end = min(base + chunk, M);
                    bounds = csr_start[min(base + lane, M)];
                    bound_last = csr_start[q_end];
                }
            }
            const int rank = __popc(need_mask & (leader_bit - 1));
            const int row = q_next + rank;
            const bool take = need && row < q_end;
            const int rel = take ? row - q_base : 0;
            const int b0 = __shfl_sync(FULL, bounds, rel & 31);
            int b1 = __shfl_sync(FULL, bounds, (rel + 1) & 31);
            if (rel + 1 == q_end - q_base) b1 = bound_last;
            if (need) {
                nv = (take && b1 - b0 <= long_cap) ? row : -1;  // long rows: summed by whole warps (splat_long_rows_body)
                ns = b0;
                ns1 = b1;
            }
            q_next = min(q_end, q_next + __popc(need_mask));
        }
        // 3. cooperative entry loads of the next batch: lane c takes entries c, c + G, ...
        const int ncnt = nv >= 0 ? min(SB, ns1 - ns) : 0;
        Ent e_next[NB];
#pragma unroll
        for (int b = 0; b < NB; b++)
            e_next[b] = (b * G + c < ncnt) ? __ldg(csr_ent + ns + b * G + c) : zero_ent(Ent());
        // 4. consume (padded slots add 0 * 0)
#pragma unroll
        for (int b = 0; b < NB; b++)
#pragma unroll
            for (int i = 0; i < G; i++) {
                const float w = __int_as_float(__shfl_sync(FULL, e_cur[b].y, (gbase + i) & 31));
                const float n = REF ? __shfl_sync(FULL, ent_norm(e_cur[b]), (gbase + i) & 31) : 1.0f;
                splat_acc<REF>(acc, w, n, q[b * G + i]);
            }
        if (v >= 0 && row_ends) {
            val4[(unsigned)v * G + c] = acc;
            acc = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        v = nv;
        s = ns;
        s1 = ns1;
        cnt_cur = ncnt;
#pragma unroll
        for (int b = 0; b < NB; b++) e_cur[b] = e_next[b];
        if (!__ballot_sync(FULL, v >= 0) && exhausted) break;
    }
}

// rows with more than long_cap entries (found at build time, any order: rows are independent)
__global__ void __launch_bounds__(kThreads) find_long_rows_kernel(const int32_t *__restrict__ csr_start, int64_t M,
                                                                  int32_t *__restrict__ long_rows,
                                                                  int *__restrict__ n_long, int long_cap) {
    const int64_t v = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (v >= M) return;
    if (csr_start[v + 1] - csr_start[v] > long_cap) long_rows[atomicAdd(n_long, 1)] = (int32_t)v;
}

// Lattices whose rows are SHORT on average (fewer than 4 entries per vertex: a Gaussian lattice with
// sxy <= 1, a bilateral lattice over a noisy image with a narrow colour bandwidth -- ADP 1088^2 with
// sxy = 1: 2.6 entries per row, DeepGlobe 612^2 with srgb = 5: 1.5) leave the row queue of the kernels
// above mostly idle: every trip reserves G gather slots per lane group and pays the queue's ballots and
// shuffles for rows that end after one or two entries.  Here a lane group simply owns row v (static
// mapping, no counter, no tail kernel) and sums its entries four per round, entry loads cooperative;
// the same ascending order of additions, so the result is bit-identical to the queue kernels'.
// Measured (B200): bilateral splat of 8 DeepGlobe 612^2 images 373 -> 268 us (step 19.3 -> 17.5 ms);
// ADP-morph Gaussian 400 -> 386; at 5.8 entries per row (HistoSegNet Gaussian) a wash, at 4.2 with
// outliers of 40+ (ADP-func bilateral) 500 -> 652: the queue kernel stays the default from 4 entries
// per row upwards (DCRF_SPLAT_SHORT_ROWS moves the threshold).
template <int G, bool REF>
__device__ __forceinline__ void splat_short_body(const int32_t *__restrict__ csr_start,
                                                 const typename CsrEnt<REF>::type *__restrict__ csr_ent,
                                                 const float4 *__restrict__ Q4, float4 *__restrict__ val4, int M,
                                                 int g_rt, int vblock) {
    typedef typename CsrEnt<REF>::type Ent;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int U = 4;                                  // entries per round
    const RowMap<G> rm(g_rt, vblock);
    const int64_t v64 = rm.row();
    const bool act = rm.lane_active() && v64 < M;
    const unsigned g = rm.g, c = rm.col();
    const unsigned v = act ? (unsigned)v64 : 0u;
    const int gbase = (int)(threadIdx.x & 31) - (int)c;
    int s = act ? csr_start[v] : 0;
    const int s1 = act ? csr_start[v + 1] : 0;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    // cooperative entry loads: lane c of the group loads entries c, c + g, ... of the round (one request
    // per group instead of U broadcast requests), pairs travel by shuffle; warp-uniform loop
    const int per_lane = (U + (int)g - 1) / (int)g;       // 1 for g >= 4
    while (__any_sync(FULL, s < s1)) {
        Ent e[U];
#pragma unroll
        for (int j = 0; j < U; j++) {
            const int idx = s + (int)c + j * (int)g;
            e[j] = (j < per_lane && (int)c + j * (int)g < U && idx < s1) ? __ldg(csr_ent + idx) : zero_ent(Ent());
        }
        float4 q[U];
        float w[U], nrm[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int src = (gbase + u % (int)g) & 31, slot = u / (int)g;
            int px = 0, wy = 0;
            float nn = 1.0f;
#pragma unroll
            for (int j = 0; j < U; j++)
                if (j == slot) {
                    px = e[j].x;
                    wy = e[j].y;
                    nn = ent_norm(e[j]);
                }
            px = __shfl_sync(FULL, px, src);
            w[u] = __int_as_float(__shfl_sync(FULL, wy, src));
            nrm[u] = REF ? __shfl_sync(FULL, nn, src) : 1.0f;
            q[u] = (s + u < s1) ? __ldg(Q4 + ((unsigned)px * g + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < U; u++) splat_acc<REF>(acc, (s + u < s1) ? w[u] : 0.f, nrm[u], q[u]);
        s += U;
    }
    if (act) val4[v * g + c] = acc;
}

template <int G, bool REF>
__global__ void __launch_bounds__(kThreads) splat_short_kernel(const int32_t *__restrict__ csr_start,
                                                               const typename CsrEnt<REF>::type *__restrict__ csr_ent,
                                                               const float4 *__restrict__ Q4,
                                                               float4 *__restrict__ val4, int M, int g_rt) {
    splat_short_body<G, REF>(csr_start, csr_ent, Q4, val4, M, g_rt, blockIdx.x);
}

// Rows with more than long_cap entries (flat image regions collapse the bilateral lattice to a few vertices
// with thousands of entries each; histology: 30 % of the entries sit in rows of 256-1600) are summed by
// whole WARPS, one row per warp, claimed from a counter: per round the warp loads 32/g * g consecutive
// entries with ONE coalesced request, lane group k sums entries [k g, (k + 1) g) of the round (g gathers in
// flight), and the groups' partial sums are added in group order -- a fixed association, so the result is
// deterministic (though not the sequential order of the specification; DCRF_ARITH_STRICT has no long rows).
// Every warp of the splat kernel does this FIRST (longest work first), then joins the row queue, which
// skips these rows -- round 2 ran a separate tail kernel after the queue kernel for the entries beyond
// the cut (a ~7 us launch on the critical path of every iteration of a small problem).
// History of the cut: with whole CTAs per row and a shared-memory tree (round 1) cap 256 / 64 / 32 gave
// 369 / 615 / 782 us for the bilateral splat of 16 HistoSegNet 321^2 images; with warps per row the cut can
// sit where the queue's one-row-per-lane-group schedule stays balanced (see profiles/README.md).
template <int G, bool REF>
__device__ __forceinline__ void splat_long_rows_body(
    const int32_t *__restrict__ csr_start, const typename CsrEnt<REF>::type *__restrict__ csr_ent,
    const float4 *__restrict__ Q4, float4 *__restrict__ val4, const int32_t *__restrict__ long_rows,
    const int *__restrict__ n_long, int g_rt, int long_cap, int *__restrict__ counter) {
    typedef typename CsrEnt<REF>::type Ent;
    constexpr unsigned FULL = 0xffffffffu;
    const int g = G ? G : g_rt;
    const int lane = threadIdx.x & 31, gpw = 32 / g;
    const int sub = lane / g, c = lane - sub * g;
    const bool on = sub < gpw;
    const int per_round = gpw * g;  // entries per warp round: group k takes entries [k g, (k + 1) g) of the round
    const int gbase = (sub * g) & 31;
    const int n = *n_long;
    if (n == 0) return;
    for (;;) {
        int i = 0;
        if (lane == 0) i = atomicAdd(counter, 1);
        i = __shfl_sync(FULL, i, 0);
        if (i >= n) break;
        const int v = long_rows[i];
        const int s0 = csr_start[v], s1 = csr_start[v + 1];
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        // one coalesced entry load per round (lane l: entry l of the round), pairs broadcast inside the group
        Ent e_cur = (on && s0 + lane < s1) ? __ldg(csr_ent + s0 + lane) : zero_ent(Ent());
        for (int s = s0; s < s1; s += per_round) {   // warp-uniform loop
            const int sn = s + per_round;
            const Ent e_next = (on && sn + lane < s1) ? __ldg(csr_ent + sn + lane) : zero_ent(Ent());
            const int cnt = min(g, s1 - (s + sub * g));   // entries of my group in this round (may be <= 0)
            float4 q[8];
            if (G) {
#pragma unroll
                for (int k = 0; k < (G ? G : 1); k++) {
                    const int px = __shfl_sync(FULL, e_cur.x, (gbase + k) & 31);
                    q[k] = (on && k < cnt) ? __ldg(Q4 + ((unsigned)px * g + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int k = 0; k < (G ? G : 1); k++) {
                    const float w = __int_as_float(__shfl_sync(FULL, e_cur.y, (gbase + k) & 31));
                    const float nrm = REF ? __shfl_sync(FULL, ent_norm(e_cur), (gbase + k) & 31) : 1.0f;
                    splat_acc<REF>(acc, w, nrm, q[k]);   // padded slots add 0 * 0
                }
            } else {
                for (int k = 0; k < g; k++) {
                    const int px = __shfl_sync(FULL, e_cur.x, (gbase + k) & 31);
                    const float w = __int_as_float(__shfl_sync(FULL, e_cur.y, (gbase + k) & 31));
                    const float nrm = REF ? __shfl_sync(FULL, ent_norm(e_cur), (gbase + k) & 31) : 1.0f;
                    const float4 qq = (on && k < cnt) ? __ldg(Q4 + ((unsigned)px * g + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    splat_acc<REF>(acc, w, nrm, qq);
                }
            }
            e_cur = e_next;
        }
        // partial sums of the groups, added in group order
        float4 tot;
        tot.x = __shfl_sync(FULL, acc.x, c);
        tot.y = __shfl_sync(FULL, acc.y, c);
        tot.z = __shfl_sync(FULL, acc.z, c);
        tot.w = __shfl_sync(FULL, acc.w, c);
        for (int k = 1; k < gpw; k++) {
            tot.x += __shfl_sync(FULL, acc.x, k * g + c);
            tot.y += __shfl_sync(FULL, acc.y, k * g + c);
            tot.z += __shfl_sync(FULL, acc.z, k * g + c);
            tot.w += __shfl_sync(FULL, acc.w, k * g + c);
        }
        if (on && sub == 0) val4[(unsigned)v * g + c] = tot;
    }
}

// The splat kernels.  The first kLongRowBlocks CTAs of the grid (scheduled first) sum the long rows, whole
// warps per row; all others run the row queue, which skips those rows: one launch, the longest work
// starts first and runs next to the queue.  With no long rows the extra CTAs exit at once.  (Letting
// the queue CTAs help with leftover long rows afterwards costs the queue loop its spill-free 48 registers.)
constexpr int kLongRowBlocks = 4 * kNumSMs;
template <int G, int SB, bool REF>
__global__ void __launch_bounds__(kThreads) splat_fast_kernel(const int32_t *__restrict__ csr_start,
                                                              const typename CsrEnt<REF>::type *__restrict__ csr_ent,
                                                              const float4 *__restrict__ Q4,
                                                              float4 *__restrict__ val4, int M, int g_rt,
                                                              int *__restrict__ row_counter, int long_cap, int chunk,
                                                              const int32_t *__restrict__ long_rows,
                                                              const int *__restrict__ n_long) {
    if (blockIdx.x < kLongRowBlocks) {
        splat_long_rows_body<G, REF>(csr_start, csr_ent, Q4, val4, long_rows, n_long, g_rt, long_cap, row_counter + 1);
        return;
    }
    splat_fast_body<G, SB, REF>(csr_start, csr_ent, Q4, val4, M, g_rt, row_counter, long_cap, chunk);
}

template <int G, int NB, int MINB, bool REF>
__global__ void __launch_bounds__(kThreads, MINB) splat_coop_kernel(const int32_t *__restrict__ csr_start,
                                                              const typename CsrEnt<REF>::type *__restrict__ csr_ent,
                                                              const float4 *__restrict__ Q4,
                                                              float4 *__restrict__ val4, int M,
                                                              int *__restrict__ row_counter, int long_cap, int chunk,
                                                              const int32_t *__restrict__ long_rows,
                                                              const int *__restrict__ n_long) {
    if (blockIdx.x < kLongRowBlocks) {
        splat_long_rows_body<G, REF>(csr_start, csr_ent, Q4, val4, long_rows, n_long, G, long_cap, row_counter + 1);
        return;
    }
    splat_coop_body<G, NB, REF>(csr_start, csr_ent, Q4, val4, M, row_counter, long_cap, chunk);
}

// ---------------------------------------------------------------------------------------------
// blur along one axis: out[v] = in[v] + 0.5 * (in[n1] + in[n2]); absent neighbour = zero row.
// SEQ reproduces the value_size<=2 association (sum in float, 0.5* and outer add in double).
// Flat mapping over the M*g float4 elements, BU elements per thread: all neighbour ids and own
// rows are requested first, then the 2*BU dependent gathers, so 3*BU 128-bit loads are in flight.
// ---------------------------------------------------------------------------------------------
template <bool SEQ>
__device__ __forceinline__ float blur1(float o, float a, float b) {
    if (SEQ) return (float)((double)o + 0.5 * (double)__fadd_rn(a, b));
    return __fadd_rn(o, __fmul_rn(0.5f, __fadd_rn(a, b)));
}

#ifndef DCRF_TUNE_BLUR_UNROLL
#define DCRF_TUNE_BLUR_UNROLL 2  // measured on B200 (batch of 32 VOC images): 1 -> 162 us, 2 -> 129 us, 3 -> 144, 4 -> 142, 8 -> 228
#endif
constexpr int kBlurUnroll = DCRF_TUNE_BLUR_UNROLL;
#ifndef DCRF_TUNE_BLUR_THREADS
#define DCRF_TUNE_BLUR_THREADS 128  // 128 -> 129 us, 256 -> 133 us, 512 -> 139 us (unroll 2)
#endif
constexpr int kBlurThreads = DCRF_TUNE_BLUR_THREADS;

// one tile of THREADS * BU float4 elements (tile index = blockIdx.x in the stand-alone kernel)
template <int G, bool SEQ, int THREADS>
__device__ __forceinline__ void blur_tile(const int2 *__restrict__ neigh, const float *__restrict__ in,
                                          float *__restrict__ out, int64_t M, int g_rt, int64_t tile) {
    constexpr int BU = kBlurUnroll;
    const int g = G ? G : g_rt;
    const int64_t total = M * g;
    const int64_t base = tile * (THREADS * BU) + threadIdx.x;
    int64_t idx[BU];
    int c[BU];
    int2 nb[BU];
    float4 o[BU], a[BU], b[BU];
#pragma unroll
    for (int u = 0; u < BU; u++) {
        idx[u] = base + (int64_t)u * THREADS;
        const bool ok = idx[u] < total;
        const int64_t v = ok ? idx[u] / g : 0;
        c[u] = (int)(idx[u] - v * g);
        nb[u] = ok ? neigh[v] : make_int2(-1, -1);
        o[u] = ok ? ldg4(in + idx[u] * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < BU; u++) {
        a[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        b[u] = a[u];
        if (nb[u].x >= 0) a[u] = ldg4(in + ((int64_t)nb[u].x * g + c[u]) * 4);
        if (nb[u].y >= 0) b[u] = ldg4(in + ((int64_t)nb[u].y * g + c[u]) * 4);
    }
#pragma unroll
    for (int u = 0; u < BU; u++) {
        if (idx[u] < total) {
            float4 r;
            r.x = blur1<SEQ>(o[u].x, a[u].x, b[u].x);
            r.y = blur1<SEQ>(o[u].y, a[u].y, b[u].y);
            r.z = blur1<SEQ>(o[u].z, a[u].z, b[u].z);
            r.w = blur1<SEQ>(o[u].w, a[u].w, b[u].w);
            st4(out + idx[u] * 4, r);
        }
    }
}

template <int G, bool SEQ>
__global__ void __launch_bounds__(kBlurThreads) blur_kernel(const int2 *__restrict__ neigh,
                                                        const float *__restrict__ in,
                                                        float *__restrict__ out, int64_t M, int g_rt) {
    blur_tile<G, SEQ, kBlurThreads>(neigh, in, out, M, g_rt, blockIdx.x);
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

// compile-time-D slice split in two phases so that callers can put the loads of SEVERAL terms in
// flight before consuming any of them: gather() issues the D+1 index/weight loads and then the D+1
// row gathers, reduce() accumulates in r order with the non-SEQ association (w*alpha)*v.
template <int D>
struct SliceGather {
    float4 v[D + 1];
    float w[D + 1];
    __device__ __forceinline__ void gather(const int32_t *__restrict__ offset, const float *__restrict__ bary,
                                           const float *__restrict__ val, int64_t p, int g, int c) {
        int o[D + 1];
        const int64_t base = p * (D + 1);
#pragma unroll
        for (int r = 0; r <= D; r++) {
            o[r] = offset[base + r];
            w[r] = bary[base + r];
        }
#pragma unroll
        for (int r = 0; r <= D; r++) v[r] = ldg4(val + ((int64_t)o[r] * g + c) * 4);
    }
    __device__ __forceinline__ float4 reduce(float alpha) const {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r <= D; r++) mul_add(acc, __fmul_rn(w[r], alpha), v[r]);
        return acc;
    }
};

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
// FAST25: exactly the reference configuration -- term 0 has d = 2, term 1 has d = 5 (Gaussian +
// bilateral), not the value_size <= 2 association: the 9 index loads, then the 9 row gathers of both
// terms, the unary row and both norms are all in flight before the first use.
template <int G, bool FAST25>
__global__ void __launch_bounds__(kThreads) slice_softmax_kernel(const SliceArgs a,
                                                                 const float *__restrict__ unary,
                                                                 float *__restrict__ Q, int64_t Ntot,
                                                                 int L, int g_rt) {
    const RowMap<G> rm(g_rt);
    const int g = rm.g, c = rm.col();
    const int64_t p = rm.row();
    const bool act = rm.lane_active() && p < Ntot;
    const int64_t pc = act ? p : 0;  // inactive lanes shadow pixel 0 so that shuffles stay uniform
    const int Lp = g * 4;
    float4 t;
    if (FAST25) {
        const SliceTerm &t0 = a.term[0];
        const SliceTerm &t1 = a.term[1];
        SliceGather<2> g0;
        SliceGather<5> g1;
        g0.gather(t0.offset, t0.bary, t0.val, pc, g, c);
        g1.gather(t1.offset, t1.bary, t1.val, pc, g, c);
        const float4 u = ldg4(unary + (pc * g + c) * 4);
        const float n0 = t0.norm ? t0.norm[pc] : 1.f;
        const float n1 = t1.norm ? t1.norm[pc] : 1.f;
        t = make_float4(-u.x, -u.y, -u.z, -u.w);
        float4 x = g0.reduce(t0.alpha);
        if (t0.norm) x = scale4(x, n0);
        float4 y = apply_compat(t0, x, g, c, Lp);
        t.x = __fsub_rn(t.x, y.x); t.y = __fsub_rn(t.y, y.y); t.z = __fsub_rn(t.z, y.z); t.w = __fsub_rn(t.w, y.w);
        x = g1.reduce(t1.alpha);
        if (t1.norm) x = scale4(x, n1);
        y = apply_compat(t1, x, g, c, Lp);
        t.x = __fsub_rn(t.x, y.x); t.y = __fsub_rn(t.y, y.y); t.z = __fsub_rn(t.z, y.z); t.w = __fsub_rn(t.w, y.w);
    } else {
        const float4 u = ldg4(unary + (pc * g + c) * 4);
        t = make_float4(-u.x, -u.y, -u.z, -u.w);
    }
    for (int k = 0; k < (FAST25 ? 0 : a.n_terms); k++) {
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
    // softmax over the L valid labels of the row: libm-identical expf, sum in label order (A.7)
    const ExpfRef ex;
    const float4 q = softmax_ref_row<G>(t, L, c, g, (int)(threadIdx.x & 31) - c, ex);
    if (act) st4(Q + (p * g + c) * 4, q);
}

// Fast fused slice for the reference configuration: term 0 with d = DA, term 1 with d = DB, both
// Potts.  t = -U + sum_k (w_k alpha_k n_k[p]) * sum_r bary_r v_r ; Q = softmax(t) with ex2.approx.
template <int D>
__device__ __forceinline__ float4 slice_fast_row(const int2 *__restrict__ ent, const float4 *__restrict__ val4,
                                                 unsigned p, unsigned g, unsigned c) {
    int2 e[D + 1];
    float4 v[D + 1];
    const int2 *ep = ent + (size_t)p * (D + 1);
    if (((D + 1) & 1) == 0) {  // even entry count: 128-bit loads of two entries
        const int4 *ep4 = reinterpret_cast<const int4 *>(ep);
#pragma unroll
        for (int r = 0; r < (D + 1) / 2; r++) {
            const int4 t = __ldg(ep4 + r);
            e[2 * r] = make_int2(t.x, t.y);
            e[2 * r + 1] = make_int2(t.z, t.w);
        }
    } else {
#pragma unroll
        for (int r = 0; r <= D; r++) e[r] = __ldg(ep + r);
    }
#pragma unroll
    for (int r = 0; r <= D; r++) v[r] = __ldg(val4 + ((unsigned)e[r].x * g + c));
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r <= D; r++) fma4(acc, __int_as_float(e[r].y), v[r]);
    return acc;
}

// Cooperative form of slice_fast_row: lane c of the pixel's group loads entries c, c + G, ... (one load
// request for the group instead of one broadcast request per 16 bytes of entries) and the (vertex,
// weight) pairs travel by shuffle.  Same values, same order of additions.
template <int D, int G>
__device__ __forceinline__ float4 slice_fast_row_coop(const int2 *__restrict__ ent, const float4 *__restrict__ val4,
                                                      unsigned p, unsigned c, int gbase) {
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int NL = (D + G) / G;
    int2 e[NL];
    const int2 *ep = ent + (size_t)p * (D + 1);
#pragma unroll
    for (int j = 0; j < NL; j++) {
        const int idx = (int)c + j * G;
        e[j] = idx <= D ? __ldg(ep + idx) : make_int2(0, 0);
    }
    float4 v[D + 1];
#pragma unroll
    for (int r = 0; r <= D; r++) {
        const unsigned vx = (unsigned)__shfl_sync(FULL, e[r / G].x, (gbase + r % G) & 31);
        v[r] = __ldg(val4 + (vx * G + c));
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r <= D; r++)
        fma4(acc, __int_as_float(__shfl_sync(FULL, e[r / G].y, (gbase + r % G) & 31)), v[r]);
    return acc;
}

// (register budget: the default heuristic's 32 registers / full occupancy is fastest -- 590 us;
// forcing 40 / 48 / 64 registers to keep more gathers in flight per thread gave 596 / 602 / 654 us)
// __launch_bounds__(.., 8): 32 registers = full occupancy; at 34 (6 resident CTAs) it takes 639 us.
template <int G, int DA, int DB>
__global__ void __launch_bounds__(kThreads, G ? 2048 / kThreads : 1) slice_softmax_fast_kernel(const SliceArgs a,
                                                                      const float4 *__restrict__ unary4,
                                                                      float4 *__restrict__ Q4, unsigned Ntot,
                                                                      int L, int g_rt) {
    const RowMap<G> rm(g_rt);
    const unsigned g = rm.g, c = rm.col();
    const int64_t p64 = rm.row();
    const bool act = rm.lane_active() && p64 < (int64_t)Ntot;
    const unsigned p = act ? (unsigned)p64 : 0u;
    const SliceTerm &t0 = a.term[0];
    const SliceTerm &t1 = a.term[1];
    float4 x0, x1;
    if (G >= 3) {  // cooperative entry loads (580 -> 567 us at G = 6, bit-identical)
        const int gb = (int)(threadIdx.x & 31) - (int)c;
        x0 = slice_fast_row_coop<DA, (G >= 3 ? G : 3)>(t0.ent, reinterpret_cast<const float4 *>(t0.val), p, c, gb);
        x1 = slice_fast_row_coop<DB, (G >= 3 ? G : 3)>(t1.ent, reinterpret_cast<const float4 *>(t1.val), p, c, gb);
    } else {
        x0 = slice_fast_row<DA>(t0.ent, reinterpret_cast<const float4 *>(t0.val), p, g, c);
        x1 = slice_fast_row<DB>(t1.ent, reinterpret_cast<const float4 *>(t1.val), p, g, c);
    }
    const float4 u = __ldg(unary4 + (p * g + c));
    // (the per-pixel post-normalisation is already inside the packed entry weights)
    const float w0 = t0.potts_w * t0.alpha, w1 = t1.potts_w * t1.alpha;
    float4 t;
    t.x = fmaf(w1, x1.x, fmaf(w0, x0.x, -u.x));
    t.y = fmaf(w1, x1.y, fmaf(w0, x0.y, -u.y));
    t.z = fmaf(w1, x1.z, fmaf(w0, x0.z, -u.z));
    t.w = fmaf(w1, x1.w, fmaf(w0, x0.w, -u.w));
    const int l0 = c * 4;
    const float NEG = -INFINITY;
    if (l0 + 0 >= L) t.x = NEG;
    if (l0 + 1 >= L) t.y = NEG;
    if (l0 + 2 >= L) t.z = NEG;
    if (l0 + 3 >= L) t.w = NEG;
    const float m = fmaxf(fmaxf(t.x, t.y), fmaxf(t.z, t.w));
    const int lane = threadIdx.x & 31;
    const int gbase = lane - c;
    float mx = NEG;
    for (int i = 0; i < (int)g; i++) mx = fmaxf(mx, __shfl_sync(0xffffffffu, m, (gbase + i) & 31));
    float4 e;
    e.x = expf(t.x - mx);  // expf(-inf) = 0 for the padding lanes
    e.y = expf(t.y - mx);
    e.z = expf(t.z - mx);
    e.w = expf(t.w - mx);
    const float ls = (e.x + e.y) + (e.z + e.w);
    float sum = 0.f;
    for (int i = 0; i < (int)g; i++) sum += __shfl_sync(0xffffffffu, ls, (gbase + i) & 31);
    if (act) {
        const float inv = 1.0f / sum;
        Q4[p * g + c] = make_float4(e.x * inv, e.y * inv, e.z * inv, e.w * inv);
    }
}

// Fast fused slice for any other combination of Potts terms (1..4 terms, any dimensions): same
// arithmetic as above with run-time loops.
template <int G>
__device__ __forceinline__ void slice_softmax_fast_generic_body(const SliceArgs &a, const float4 *__restrict__ unary4,
                                                                float4 *__restrict__ Q4, unsigned Ntot, int L,
                                                                int g_rt, int vblock) {
    const RowMap<G> rm(g_rt, vblock);
    const unsigned g = rm.g, c = rm.col();
    const int64_t p64 = rm.row();
    const bool act = rm.lane_active() && p64 < (int64_t)Ntot;
    const unsigned p = act ? (unsigned)p64 : 0u;
    const float4 u = __ldg(unary4 + (p * g + c));
    float4 t = make_float4(-u.x, -u.y, -u.z, -u.w);
    for (int k = 0; k < a.n_terms; k++) {
        const SliceTerm &tm = a.term[k];
        const int2 *ep = tm.ent + (size_t)p * (tm.d + 1);
        const float4 *val4 = reinterpret_cast<const float4 *>(tm.val);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r <= tm.d; r++) {
            const int2 e = __ldg(ep + r);
            fma4(acc, __int_as_float(e.y), __ldg(val4 + ((unsigned)e.x * g + c)));
        }
        const float w = tm.potts_w * tm.alpha;  // post-normalisation: inside the entry weights
        t.x = fmaf(w, acc.x, t.x);
        t.y = fmaf(w, acc.y, t.y);
        t.z = fmaf(w, acc.z, t.z);
        t.w = fmaf(w, acc.w, t.w);
    }
    const int l0 = c * 4;
    const float NEG = -INFINITY;
    if (l0 + 0 >= L) t.x = NEG;
    if (l0 + 1 >= L) t.y = NEG;
    if (l0 + 2 >= L) t.z = NEG;
    if (l0 + 3 >= L) t.w = NEG;
    const float m = fmaxf(fmaxf(t.x, t.y), fmaxf(t.z, t.w));
    const int lane = threadIdx.x & 31;
    const int gbase = lane - c;
    float mx = NEG;
    for (int i = 0; i < (int)g; i++) mx = fmaxf(mx, __shfl_sync(0xffffffffu, m, (gbase + i) & 31));
    float4 e;
    e.x = expf(t.x - mx);
    e.y = expf(t.y - mx);
    e.z = expf(t.z - mx);
    e.w = expf(t.w - mx);
    const float ls = (e.x + e.y) + (e.z + e.w);
    float sum = 0.f;
    for (int i = 0; i < (int)g; i++) sum += __shfl_sync(0xffffffffu, ls, (gbase + i) & 31);
    if (act) {
        const float inv = 1.0f / sum;
        Q4[p * g + c] = make_float4(e.x * inv, e.y * inv, e.z * inv, e.w * inv);
    }
}

template <int G>
__global__ void __launch_bounds__(kThreads) slice_softmax_fast_generic_kernel(const SliceArgs a,
                                                                              const float4 *__restrict__ unary4,
                                                                              float4 *__restrict__ Q4, unsigned Ntot,
                                                                              int L, int g_rt) {
    slice_softmax_fast_generic_body<G>(a, unary4, Q4, Ntot, L, g_rt, blockIdx.x);
}

// Q0 = softmax(-U) (startInference), fast-math variant of slice_softmax_kernel with no terms
template <int G>
__global__ void __launch_bounds__(kThreads) softmax_unary_fast_kernel(const float4 *__restrict__ unary4,
                                                                      float4 *__restrict__ Q4, unsigned Ntot,
                                                                      int L, int g_rt) {
    const RowMap<G> rm(g_rt);
    const unsigned g = rm.g, c = rm.col();
    const int64_t p64 = rm.row();
    const bool act = rm.lane_active() && p64 < (int64_t)Ntot;
    const unsigned p = act ? (unsigned)p64 : 0u;
    const float4 u = __ldg(unary4 + (p * g + c));
    float4 t = make_float4(-u.x, -u.y, -u.z, -u.w);
    const int l0 = c * 4;
    const float NEG = -INFINITY;
    if (l0 + 0 >= L) t.x = NEG;
    if (l0 + 1 >= L) t.y = NEG;
    if (l0 + 2 >= L) t.z = NEG;
    if (l0 + 3 >= L) t.w = NEG;
    const float m = fmaxf(fmaxf(t.x, t.y), fmaxf(t.z, t.w));
    const int lane = threadIdx.x & 31;
    const int gbase = lane - c;
    float mx = NEG;
    for (int i = 0; i < (int)g; i++) mx = fmaxf(mx, __shfl_sync(0xffffffffu, m, (gbase + i) & 31));
    float4 e;
    e.x = expf(t.x - mx);
    e.y = expf(t.y - mx);
    e.z = expf(t.z - mx);
    e.w = expf(t.w - mx);
    const float ls = (e.x + e.y) + (e.z + e.w);
    float sum = 0.f;
    for (int i = 0; i < (int)g; i++) sum += __shfl_sync(0xffffffffu, ls, (gbase + i) & 31);
    if (act) {
        const float inv = 1.0f / sum;
        Q4[p * g + c] = make_float4(e.x * inv, e.y * inv, e.z * inv, e.w * inv);
    }
}

// ---------------------------------------------------------------------------------------------
// Reference-association fused slice (kSliceRef): the packed-table / cooperative-load structure of the
// FMA kernels above with the arithmetic of the sequential CPU evaluation, operation for operation:
//   x_k = sum_r (bary_r * alpha_k) * v_r   separately rounded products and sums, r ascending  (A.4)
//   y_k = (-w_k) * (x_k * norm_k[p])                                                          (A.5, A.6)
//   t   = ((-U) - y_0) - y_1 ...;   Q = softmax_ref_row(t)                                    (A.7)
// so Q is bit-identical to oracle/densecrf_oracle.c whenever the lattice values are.
// ---------------------------------------------------------------------------------------------
template <int D, int G>
__device__ __forceinline__ float4 slice_ref_row_coop(const int2 *__restrict__ ent, const float4 *__restrict__ val4,
                                                     unsigned p, unsigned c, int gbase) {
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int NL = (D + G) / G;
    int2 e[NL];
    const int2 *ep = ent + (size_t)p * (D + 1);
#pragma unroll
    for (int j = 0; j < NL; j++) {
        const int idx = (int)c + j * G;
        e[j] = idx <= D ? __ldg(ep + idx) : make_int2(0, 0);
    }
    float4 v[D + 1];
#pragma unroll
    for (int r = 0; r <= D; r++) {
        const unsigned vx = (unsigned)__shfl_sync(FULL, e[r / G].x, (gbase + r % G) & 31);
        v[r] = __ldg(val4 + (vx * G + c));
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r <= D; r++)
        mul_add(acc, __int_as_float(__shfl_sync(FULL, e[r / G].y, (gbase + r % G) & 31)), v[r]);
    return acc;
}

__device__ __forceinline__ float4 slice_ref_row_rt(const int2 *__restrict__ ent, const float4 *__restrict__ val4,
                                                   unsigned p, int d, unsigned g, unsigned c) {
    const int2 *ep = ent + (size_t)p * (d + 1);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r0 = 0; r0 <= d; r0 += 4) {  // 4 gathers in flight
        int2 e[4];
        float4 v[4];
#pragma unroll
        for (int i = 0; i < 4; i++) e[i] = r0 + i <= d ? __ldg(ep + r0 + i) : make_int2(0, 0);
#pragma unroll
        for (int i = 0; i < 4; i++)
            v[i] = r0 + i <= d ? __ldg(val4 + ((unsigned)e[i].x * g + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (r0 + i <= d) mul_add(acc, __int_as_float(e[i].y), v[i]);
    }
    return acc;
}

// t -= (-w) * (x * n)
__device__ __forceinline__ void sub_potts(float4 &t, const float4 x, float n, float w) {
    const float nw = -w;
    t.x = __fsub_rn(t.x, __fmul_rn(nw, __fmul_rn(x.x, n)));
    t.y = __fsub_rn(t.y, __fmul_rn(nw, __fmul_rn(x.y, n)));
    t.z = __fsub_rn(t.z, __fmul_rn(nw, __fmul_rn(x.z, n)));
    t.w = __fsub_rn(t.w, __fmul_rn(nw, __fmul_rn(x.w, n)));
}

#ifndef DCRF_TUNE_REF_MINB
#define DCRF_TUNE_REF_MINB 6
#endif
template <int G, int DA, int DB>
__global__ void __launch_bounds__(kThreads, G ? DCRF_TUNE_REF_MINB : 1) slice_softmax_ref_kernel(
    const SliceArgs a, const float4 *__restrict__ unary4, float4 *__restrict__ Q4, unsigned Ntot, int L, int g_rt) {
    const RowMap<G> rm(g_rt);
    const unsigned g = rm.g, c = rm.col();
    const int64_t p64 = rm.row();
    const bool act = rm.lane_active() && p64 < (int64_t)Ntot;
    const unsigned p = act ? (unsigned)p64 : 0u;
    const SliceTerm &t0 = a.term[0];
    const SliceTerm &t1 = a.term[1];
    const int gb = (int)(threadIdx.x & 31) - (int)c;
    float4 x0, x1;
    if (G >= 3) {
        x0 = slice_ref_row_coop<DA, (G >= 3 ? G : 3)>(t0.ent, reinterpret_cast<const float4 *>(t0.val), p, c, gb);
        x1 = slice_ref_row_coop<DB, (G >= 3 ? G : 3)>(t1.ent, reinterpret_cast<const float4 *>(t1.val), p, c, gb);
    } else {
        x0 = slice_ref_row_rt(t0.ent, reinterpret_cast<const float4 *>(t0.val), p, DA, g, c);
        x1 = slice_ref_row_rt(t1.ent, reinterpret_cast<const float4 *>(t1.val), p, DB, g, c);
    }
    const float4 u = __ldg(unary4 + (p * g + c));
    const float n0 = t0.norm ? __ldg(t0.norm + p) : 1.0f;  // x * 1.0f is exact
    const float n1 = t1.norm ? __ldg(t1.norm + p) : 1.0f;
    float4 t = make_float4(-u.x, -u.y, -u.z, -u.w);
    sub_potts(t, x0, n0, t0.potts_w);
    sub_potts(t, x1, n1, t1.potts_w);
    const ExpfRef ex;
    const float4 q = softmax_ref_row<G>(t, L, (int)c, (int)g, gb, ex);
    if (act) Q4[p * g + c] = q;
}

// any other combination of Potts terms (0..4 terms, any dimensions; n_terms = 0 is startInference)
template <int G>
__device__ __forceinline__ void slice_softmax_ref_generic_body(const SliceArgs &a, const float4 *__restrict__ unary4,
                                                               float4 *__restrict__ Q4, unsigned Ntot, int L,
                                                               int g_rt, int vblock) {
    const RowMap<G> rm(g_rt, vblock);
    const unsigned g = rm.g, c = rm.col();
    const int64_t p64 = rm.row();
    const bool act = rm.lane_active() && p64 < (int64_t)Ntot;
    const unsigned p = act ? (unsigned)p64 : 0u;
    const int gb = (int)(threadIdx.x & 31) - (int)c;
    const float4 u = __ldg(unary4 + (p * g + c));
    float4 t = make_float4(-u.x, -u.y, -u.z, -u.w);
    for (int k = 0; k < a.n_terms; k++) {
        const SliceTerm &tm = a.term[k];
        const float4 x = slice_ref_row_rt(tm.ent, reinterpret_cast<const float4 *>(tm.val), p, tm.d, g, c);
        const float n = tm.norm ? __ldg(tm.norm + p) : 1.0f;
        sub_potts(t, x, n, tm.potts_w);
    }
    const ExpfRef ex;
    const float4 q = softmax_ref_row<G>(t, L, (int)c, (int)g, gb, ex);
    if (act) Q4[p * g + c] = q;
}

template <int G>
__global__ void __launch_bounds__(kThreads) slice_softmax_ref_generic_kernel(
    const SliceArgs a, const float4 *__restrict__ unary4, float4 *__restrict__ Q4, unsigned Ntot, int L, int g_rt) {
    slice_softmax_ref_generic_body<G>(a, unary4, Q4, Ntot, L, g_rt, blockIdx.x);
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

// ---------------------------------------------------------------------------------------------
// Kernel normalisation (A.5): norm = f( filter(all-ones) ) through the value_size = 1 association
// of A.4 (splat w * 1, blur through a double 0.5, slice (w * v) * alpha).  Scalar kernels: one float
// per vertex / pixel instead of a padded float4 row.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) norm_splat_kernel(const int32_t *__restrict__ csr_start,
                                                              const float *__restrict__ csr_w,
                                                              float *__restrict__ val, int64_t M) {
    const int64_t v = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (v >= M) return;
    float acc = 0.f;
    const int s1 = csr_start[v + 1];
    for (int s = csr_start[v]; s < s1; s++) acc = __fadd_rn(acc, __fmul_rn(csr_w[s], 1.0f));
    val[v] = acc;
}
__global__ void __launch_bounds__(kThreads) norm_blur_kernel(const int2 *__restrict__ neigh,
                                                             const float *__restrict__ in,
                                                             float *__restrict__ out, int64_t M) {
    const int64_t v = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (v >= M) return;
    const int2 nb = neigh[v];
    const float a = nb.x >= 0 ? in[nb.x] : 0.f, b = nb.y >= 0 ? in[nb.y] : 0.f;
    out[v] = (float)((double)in[v] + 0.5 * (double)__fadd_rn(a, b));
}
__global__ void __launch_bounds__(kThreads) norm_slice_kernel(const int32_t *__restrict__ offset,
                                                              const float *__restrict__ bary,
                                                              const float *__restrict__ val, int d, float alpha,
                                                              float *__restrict__ norm, int64_t N, int ntype) {
    const int64_t p = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (p >= N) return;
    float acc = 0.f;
    const int64_t base = p * (d + 1);
    for (int r = 0; r <= d; r++)
        acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(bary[base + r], val[offset[base + r]]), alpha));
    float res;
    if (ntype == DCRF_NORMALIZE_SYMMETRIC) res = (float)(1.0 / sqrt((double)acc + 1e-20));
    else res = (float)(1.0 / ((double)acc + 1e-20));
    norm[p] = res;
}

// ---------------------------------------------------------------------------------------------
// layout changes at the API boundary: (L, N_b) row-major blocks <-> (Ntot, Lp) pixel-major
// ---------------------------------------------------------------------------------------------
constexpr int kTP = 128;  // pixels per tile; 256 threads per CTA; tile[Lp][kTP + 1] floats of shared memory
// grid (ceil(maxN / 128), B).  Label rows are read 128 pixels at a time (4 independent coalesced
// loads per lane), the pixel-major side moves as float4.
__global__ void __launch_bounds__(kThreads) ln_to_pm_kernel(const float *__restrict__ ln,
                                                            float *__restrict__ pm,
                                                            const int *__restrict__ pix_start, int L, int Lp) {
    extern __shared__ float tile[];
    const int b = blockIdx.y;
    const int64_t ps = pix_start[b];
    const int Nb = (int)(pix_start[b + 1] - ps);
    const int p0 = blockIdx.x * kTP;
    if (p0 >= Nb) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float *src = ln + ps * L + p0;  // image block (L, Nb)
    const int np = min(kTP, Nb - p0);
    for (int l = warp; l < Lp; l += kWarps) {
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int pp = lane + 32 * k;
            v[k] = (l < L && pp < np) ? src[(int64_t)l * Nb + pp] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 4; k++) tile[l * (kTP + 1) + lane + 32 * k] = v[k];
    }
    __syncthreads();
    float4 *dst = reinterpret_cast<float4 *>(pm + (ps + p0) * Lp);
    const int g = Lp >> 2;
    for (int i = threadIdx.x; i < np * g; i += kThreads) {
        const int pp = i / g, l = (i - pp * g) * 4;
        dst[i] = make_float4(tile[l * (kTP + 1) + pp], tile[(l + 1) * (kTP + 1) + pp],
                             tile[(l + 2) * (kTP + 1) + pp], tile[(l + 3) * (kTP + 1) + pp]);
    }
}

__global__ void __launch_bounds__(kThreads) pm_to_ln_kernel(const float *__restrict__ pm,
                                                            float *__restrict__ ln,
                                                            const int *__restrict__ pix_start, int L, int Lp) {
    extern __shared__ float tile[];
    const int b = blockIdx.y;
    const int64_t ps = pix_start[b];
    const int Nb = (int)(pix_start[b + 1] - ps);
    const int p0 = blockIdx.x * kTP;
    if (p0 >= Nb) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int np = min(kTP, Nb - p0);
    const float4 *src = reinterpret_cast<const float4 *>(pm + (ps + p0) * Lp);
    const int g = Lp >> 2;
    for (int i = threadIdx.x; i < np * g; i += kThreads) {
        const int pp = i / g, l = (i - pp * g) * 4;
        const float4 v = __ldg(src + i);
        tile[l * (kTP + 1) + pp] = v.x;
        tile[(l + 1) * (kTP + 1) + pp] = v.y;
        tile[(l + 2) * (kTP + 1) + pp] = v.z;
        tile[(l + 3) * (kTP + 1) + pp] = v.w;
    }
    __syncthreads();
    float *dst = ln + ps * L + p0;
    for (int l = warp; l < L; l += kWarps) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int pp = lane + 32 * k;
            if (pp < np) dst[(int64_t)l * Nb + pp] = tile[l * (kTP + 1) + pp];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Unary construction on the GPU (the NumPy glue of the reference's wrappers):
// results go straight into the pixel-major (Ntot, Lp) unary buffer.
// ---------------------------------------------------------------------------------------------
// unary_from_softmax (pydensecrf.utils [EXT], called at 03c_hsn/utilities.py:431):
//   U = -log(clip(scale * p + (1 - scale) / L, clip, 1)) computed in double like NumPy, stored as f32.
// probs: concatenated (L, N_b) blocks, float64 or float32.  grid (ceil(maxN/128), B).
template <typename T>
__global__ void __launch_bounds__(kThreads) unary_from_probs_kernel(const T *__restrict__ probs,
                                                                    float *__restrict__ pm,
                                                                    const int *__restrict__ pix_start, int L, int Lp,
                                                                    double scale, double clip_lo, int has_clip) {
    extern __shared__ float tile[];
    const int b = blockIdx.y;
    const int64_t ps = pix_start[b];
    const int Nb = (int)(pix_start[b + 1] - ps);
    const int p0 = blockIdx.x * kTP;
    if (p0 >= Nb) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const T *src = probs + ps * L + p0;
    const int np = min(kTP, Nb - p0);
    const double uni = (1.0 - scale) / (double)L;
    for (int l = warp; l < Lp; l += kWarps) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int pp = lane + 32 * k;
            float u = 0.f;
            if (l < L && pp < np) {
                double v = (double)src[(int64_t)l * Nb + pp];
                if (scale != 1.0) v = scale * v + uni;
                if (has_clip) v = fmin(fmax(v, clip_lo), 1.0);
                u = (float)(-log(v));
            }
            tile[l * (kTP + 1) + pp] = u;
        }
    }
    __syncthreads();
    float4 *dst = reinterpret_cast<float4 *>(pm + (ps + p0) * Lp);
    const int g = Lp >> 2;
    for (int i = threadIdx.x; i < np * g; i += kThreads) {
        const int pp = i / g, l = (i - pp * g) * 4;
        dst[i] = make_float4(tile[l * (kTP + 1) + pp], tile[(l + 1) * (kTP + 1) + pp],
                             tile[(l + 2) * (kTP + 1) + pp], tile[(l + 3) * (kTP + 1) + pp]);
    }
}

// crf_inference(use_log = True) of SEC/DSRG ([EXT] lib/crf.py; call sites SEC.py:275, model.py:689):
//   p = softmax over the channel axis of an (H, W, C) feature map, U = -log p  (float32 math).
// feat: concatenated (N_b, C) blocks = already pixel-major.  One lane group per pixel.
template <int G>
__global__ void __launch_bounds__(kThreads) unary_from_logits_kernel(const float *__restrict__ feat,
                                                                     float *__restrict__ pm, int64_t Ntot, int L,
                                                                     int g_rt, int use_log) {
    const RowMap<G> rm(g_rt);
    const int g = rm.g, c = rm.col();
    const int64_t p64 = rm.row();
    const bool act = rm.lane_active() && p64 < Ntot;
    const int64_t p = act ? p64 : 0;
    float f[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int l = c * 4 + i;
        f[i] = l < L ? feat[p * L + l] : -INFINITY;
    }
    const int lane = threadIdx.x & 31, gbase = lane - c;
    float u[4];
    if (use_log) {
        float m = fmaxf(fmaxf(f[0], f[1]), fmaxf(f[2], f[3])), mx = -INFINITY;
        for (int i = 0; i < g; i++) mx = fmaxf(mx, __shfl_sync(0xffffffffu, m, (gbase + i) & 31));
        float e[4], ls = 0.f;
#pragma unroll
        for (int i = 0; i < 4; i++) { e[i] = expf(f[i] - mx); ls += e[i]; }
        float sum = 0.f;
        for (int i = 0; i < g; i++) sum += __shfl_sync(0xffffffffu, ls, (gbase + i) & 31);
#pragma unroll
        for (int i = 0; i < 4; i++) u[i] = -logf(e[i] / sum);
    } else {
#pragma unroll
        for (int i = 0; i < 4; i++) u[i] = -logf(f[i]);
    }
    if (act) {
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (c * 4 + i >= L) u[i] = 0.f;
        st4(pm + (p * g + c) * 4, make_float4(u[0], u[1], u[2], u[3]));
    }
}

// unary_from_labels (pydensecrf.utils [EXT]; used by crf_inference_label, cam_to_ir_label.py:35)
__global__ void __launch_bounds__(kThreads) unary_from_labels_kernel(const int32_t *__restrict__ labels,
                                                                     float *__restrict__ pm, int64_t Ntot, int L,
                                                                     int Lp, float n_energy, float p_energy,
                                                                     float unsure_energy, int zero_unsure,
                                                                     int *__restrict__ bad) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= Ntot * Lp) return;
    const int64_t p = i / Lp;
    const int l = (int)(i - p * Lp);
    const int lab = labels[p];
    float u = 0.f;
    if (l < L) {
        if (zero_unsure) {
            // classes are 1-based, 0 = unsure (uniform); NumPy's U[labels - 1] wraps label 0 to L - 1
            // before the unsure columns are overwritten, so only the final state matters
            if (lab == 0) u = unsure_energy;
            else u = (l == lab - 1) ? p_energy : n_energy;
            if (lab < 0 || lab > L) atomicExch(bad, 1);
        } else {
            u = (l == lab) ? p_energy : n_energy;
            if (lab < 0 || lab >= L) atomicExch(bad, 1);
        }
    }
    pm[i] = u;
}

// first maximum wins, like np.argmax; T = int32_t or uint8_t (L <= 256: the form label consumers download)
template <typename T>
__global__ void __launch_bounds__(kThreads) argmax_kernel(const float *__restrict__ pm, T *__restrict__ labels,
                                                          int64_t Ntot, int L, int Lp) {
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
    labels[p] = (T)bi;
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

// ---------------------------------------------------------------------------------------------
// Persistent mean field for SMALL problems (one VOC image, a batch of 41x41 SEC maps): the whole of
// `inference(n)` -- Q0 = softmax(-U), then n x [splat, d+1 blurs, fused slice] -- as
// ONE cooperative launch with grid barriers between the phases.  With a kernel per phase such
// problems are launch-latency bound (~16 launches of 5-20 us per iteration); here an iteration costs
// its work plus ~9 grid barriers.  The phases run the same device bodies as the stand-alone kernels
// (same arithmetic, same summation order), so the marginals are bit-identical to the launch-per-phase
// path in every arithmetic mode.
// ---------------------------------------------------------------------------------------------
struct MfTerm {
    const int32_t *csr_start;
    const void *csr_ent;  // int2 (FMA tables) or int4 (reference tables)
    const int2 *neigh;
    const int32_t *long_rows;
    const int *n_long;
    float *valA, *valB;
    int M, d, long_cap, long_hint, short_rows;
};
struct MfArgs {
    MfTerm term[kMaxPairwise];
    SliceArgs slice;  // term[k].val is set inside the kernel (the blurred buffer of the iteration)
    const float4 *unary4;
    float4 *Q4;
    unsigned Ntot;
    int L, g, n_iter;
    int *counters;  // [2 * n_iter * n_terms] row dispensers of the splats and of their long rows, zeroed before the launch
};

template <int G, bool REF>
__global__ void __launch_bounds__(kThreads) mean_field_persistent_kernel(const MfArgs a) {
    namespace cg = cooperative_groups;
    typedef typename CsrEnt<REF>::type Ent;
    cg::grid_group grid = cg::this_grid();
    const int nblk = gridDim.x, g = a.g, nt = a.slice.n_terms;
    const int n_slice_blocks = (int)(((int64_t)a.Ntot + kWarps * (32 / g) - 1) / (kWarps * (32 / g)));
    SliceArgs sa = a.slice;
    {   // Q0 = softmax(-U)
        SliceArgs s0 = sa;
        s0.n_terms = 0;
        for (int vb = blockIdx.x; vb < n_slice_blocks; vb += nblk) {
            if (REF) slice_softmax_ref_generic_body<G>(s0, a.unary4, a.Q4, a.Ntot, a.L, g, vb);
            else slice_softmax_fast_generic_body<G>(s0, a.unary4, a.Q4, a.Ntot, a.L, g, vb);
        }
    }
    grid.sync();
    int dmax = 0;
    for (int k = 0; k < nt; k++) dmax = max(dmax, a.term[k].d);
    for (int it = 0; it < a.n_iter; it++) {
        // splat of every term (dynamic row queues: no barrier between the terms)
        for (int k = 0; k < nt; k++) {
            const MfTerm &t = a.term[k];
            int *ctr = a.counters + it * nt + k;
            float4 *v4 = reinterpret_cast<float4 *>(t.valA);
            const Ent *ents = reinterpret_cast<const Ent *>(t.csr_ent);
            if (t.short_rows) {   // same kernel choice as the launch-per-phase path (bit-identical results)
                const int nb_rows = (t.M + kWarps * (32 / g) - 1) / (kWarps * (32 / g));
                for (int vb = blockIdx.x; vb < nb_rows; vb += nblk)
                    splat_short_body<G, REF>(t.csr_start, ents, a.Q4, v4, t.M, g, vb);
                continue;
            }
            splat_long_rows_body<G, REF>(t.csr_start, ents, a.Q4, v4, t.long_rows, t.n_long, g, t.long_cap,
                                         a.counters + (a.n_iter + it) * nt + k);
            if (G >= 4 && G <= 8) {
                splat_coop_body<(G >= 4 && G <= 8) ? G : 4, 1, REF>(t.csr_start, ents, a.Q4, v4, t.M, ctr, t.long_cap);
            } else if (t.long_hint) {
                splat_fast_body<G, 8, REF>(t.csr_start, ents, a.Q4, v4, t.M, g, ctr, t.long_cap);
            } else {
                splat_fast_body<G, kSplatBatch, REF>(t.csr_start, ents, a.Q4, v4, t.M, g, ctr, t.long_cap);
            }
        }
        grid.sync();
        // blurs: axis j of every term that has it, ping-pong A <-> B
        for (int j = 0; j <= dmax; j++) {
            for (int k = 0; k < nt; k++) {
                const MfTerm &t = a.term[k];
                if (j > t.d) continue;
                const float *in = (j & 1) ? t.valB : t.valA;
                float *out = (j & 1) ? t.valA : t.valB;
                const int64_t tiles = ((int64_t)t.M * g + kThreads * kBlurUnroll - 1) / (kThreads * kBlurUnroll);
                for (int64_t tile = blockIdx.x; tile < tiles; tile += nblk)
                    blur_tile<G, false, kThreads>(t.neigh + (int64_t)j * t.M, in, out, t.M, g, tile);
            }
            grid.sync();
        }
        for (int k = 0; k < nt; k++) sa.term[k].val = ((a.term[k].d + 1) & 1) ? a.term[k].valB : a.term[k].valA;
        for (int vb = blockIdx.x; vb < n_slice_blocks; vb += nblk) {
            if (REF) slice_softmax_ref_generic_body<G>(sa, a.unary4, a.Q4, a.Ntot, a.L, g, vb);
            else slice_softmax_fast_generic_body<G>(sa, a.unary4, a.Q4, a.Ntot, a.L, g, vb);
        }
        if (it + 1 < a.n_iter) grid.sync();
    }
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
    // persistent grid: a few CTAs per SM, each thread group walks rows v, v + n_groups, ...
    const int nb = (int)std::min<int64_t>(ceil_div(lat.M * g, kThreads), (int64_t)kNumSMs * kSplatBlocksPerSM);
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

void launch_find_long_rows(Lattice &lat, cudaStream_t s) {
    if (lat.long_row_cap <= 0) lat.long_row_cap = kSplatLongRow;
    lat.long_rows.alloc((size_t)(lat.E / lat.long_row_cap + 1), s);
    lat.n_long.alloc(1, s);
    DCRF_CUDA(cudaMemsetAsync(lat.n_long.p, 0, sizeof(int), s));
    if (lat.M == 0) return;
    find_long_rows_kernel<<<ceil_div(lat.M, kThreads), kThreads, 0, s>>>(lat.csr_start.p, lat.M, lat.long_rows.p,
                                                                      lat.n_long.p, lat.long_row_cap);
    DCRF_LAUNCHED();
}

void launch_pack_fast_tables(Lattice &lat, const float *norm_pre, const float *norm_post, cudaStream_t s) {
    if (lat.E == 0) return;
    lat.csr_ent4.release();
    lat.ent.alloc(lat.E, s);
    lat.csr_ent.alloc(lat.E, s);
    lat.row_counter.alloc(2, s);
    static const int env_cap = [] {
        const char *e = getenv("DCRF_SPLAT_LONG_ROW");
        return e ? atoi(e) : 0;
    }();
    // Where rows are long on average (histology: 50-75 entries per bilateral vertex, 30 % of the entries in
    // rows of 256-1600) a lane group per row leaves too few, too uneven work units: cut rows at 48 entries
    // and let whole warps sum the rest (bilateral splat of 16 HistoSegNet 321^2 images: 369 -> 208 us).
    // Short-row lattices keep the cut at 256 (VOC: 450 us at 96-256, 468 at 48, 646 at 24).
    // Small lattices (a single image): a 256-entry row is a chain of 43 dependent trips of one lane group
    // (~40 us) that nothing hides -- cut at 48 there as well (one VOC image, bilateral splat: 48 -> ? us).
    const bool small = lat.E < (int64_t)kSmallLatticeEntries;
    lat.long_row_cap = env_cap > 0 ? env_cap : ((lat.E >= 32 * lat.M || small) ? 48 : kSplatLongRow);
    launch_find_long_rows(lat, s);
    pack_fast_tables_kernel<<<ceil_div(lat.E, kThreads), kThreads, 0, s>>>(
        lat.offset.p, lat.bary.p, lat.csr_pix.p, lat.csr_w.p, norm_pre, norm_post, lat.d + 1, lat.ent.p,
        lat.csr_ent.p, lat.E);
    DCRF_LAUNCHED();
    lat.table_mode = kTablesFma;
}

void launch_pack_ref_tables(Lattice &lat, const float *norm_pre, int long_row_cap, cudaStream_t s) {
    if (lat.E == 0) return;
    lat.csr_ent.release();
    lat.ent.alloc(lat.E, s);
    lat.csr_ent4.alloc(lat.E, s);
    lat.row_counter.alloc(2, s);
    lat.long_row_cap = long_row_cap > 0 ? long_row_cap : kSplatLongRow;
    launch_find_long_rows(lat, s);
    const float alpha = 1.0f / (1.0f + powf(2.0f, (float)-lat.d));
    pack_ref_tables_kernel<<<ceil_div(lat.E, kThreads), kThreads, 0, s>>>(
        lat.offset.p, lat.bary.p, lat.csr_pix.p, lat.csr_w.p, norm_pre, alpha, lat.ent.p, lat.csr_ent4.p, lat.E);
    DCRF_LAUNCHED();
    lat.table_mode = kTablesRef;
}

template <typename K>
static int resident_blocks_per_sm(K kernel) {
    int n = 0;
    DCRF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kThreads, 0));
    return n > 0 ? n : 1;
}

// mean entries per row below which the static short-row kernel runs (DCRF_SPLAT_SHORT_ROWS, default 4)
static bool splat_uses_short_rows(const Lattice &lat) {
    static const int short_rows_below = [] {
        const char *e = getenv("DCRF_SPLAT_SHORT_ROWS");
        return e ? atoi(e) : 4;
    }();
    return lat.E < (int64_t)short_rows_below * lat.M;
}

// Rows a warp claims from the queue per atomic.  Large lattices: 32 (one atomic per 32 rows, ~6 rows per
// lane group).  Small ones (one VOC image: 126 k rows for 5920 resident warps; its Gaussian lattice:
// 24 k) would leave most warps without a chunk and make the launch as long as one warp's ~6 rows per
// lane group in sequence: there a chunk is ~ half a warp's fair share, at least one row per lane group.
static int splat_chunk(int64_t M, int g, int nb) {
    const int64_t warps = (int64_t)nb * kWarps;
    const int64_t fair = M / (2 * std::max<int64_t>(warps, 1));
    return (int)std::max<int64_t>(32 / g, std::min<int64_t>(kSplatChunk, fair));
}

// REF = false: FMA tables (csr_ent); REF = true: reference-association tables (csr_ent4)
template <bool REF>
static void launch_splat_packed(const Lattice &lat, const float *Q, float *val, int Lp, cudaStream_t s) {
    typedef typename CsrEnt<REF>::type Ent;
    const int g = Lp / 4;
    const Ent *ents;
    if (REF) ents = reinterpret_cast<const Ent *>(lat.csr_ent4.p);
    else ents = reinterpret_cast<const Ent *>(lat.csr_ent.p);
    const int cap = lat.long_row_cap;
    const float4 *q4s = reinterpret_cast<const float4 *>(Q);
    float4 *v4s = reinterpret_cast<float4 *>(val);
    if (splat_uses_short_rows(lat)) {
        ProfScope prof(DCRF_K_SPLAT, lat.d, s);
        const int nb = ceil_div(lat.M, rows_per_block(g));
        DCRF_DISPATCH_G(g, { splat_short_kernel<G, REF><<<nb, kThreads, 0, s>>>(lat.csr_start.p, ents, q4s, v4s, (int)lat.M, g); });
        DCRF_LAUNCHED();
        return;
    }
    DCRF_CUDA(cudaMemsetAsync(lat.row_counter.p, 0, 2 * sizeof(int), s));  // [0] rows, [1] long rows
    ProfScope prof(DCRF_K_SPLAT, lat.d, s);
    // entries per trip: long rows (Gaussian lattice, ~23 entries) amortise the loop overhead over 8
    // entries, short skewed rows (bilateral lattice, median 6) waste fewer predicated slots with 4
    const bool long_rows = lat.E >= 16 * lat.M;
    const float4 *q4 = reinterpret_cast<const float4 *>(Q);
    float4 *v4 = reinterpret_cast<float4 *>(val);
    // 4 <= G <= 8 (13..32 labels): cooperative entry loads, G entries per trip (bilateral splat of the
    // VOC batch: 417 us against 463 us for splat_fast_kernel<6, 4>, Gaussian 206 against 218 <6, 8>)
    if (g >= 4 && g <= 8) {
#define DCRF_COOP_LAUNCH(GG, MINB)                                                                          \
    case GG: {                                                                                              \
        constexpr int MB = (REF && kSplatRefMinBlocks) ? kSplatRefMinBlocks : MINB;                         \
        static const int per_sm = resident_blocks_per_sm(splat_coop_kernel<GG, 1, MB, REF>);                \
        const int nb = (int)std::min<int64_t>(ceil_div(lat.M * g, kThreads), (int64_t)kNumSMs * per_sm);   \
        splat_coop_kernel<GG, 1, MB, REF><<<nb + kLongRowBlocks, kThreads, 0, s>>>(                                       \
            lat.csr_start.p, ents, q4, v4, (int)lat.M, lat.row_counter.p, cap, splat_chunk(lat.M, g, nb),  \
            lat.long_rows.p, lat.n_long.p);                                                                \
    } break;
        switch (g) {
            DCRF_COOP_LAUNCH(4, 5)
            DCRF_COOP_LAUNCH(5, 5)
            DCRF_COOP_LAUNCH(6, 5)
            DCRF_COOP_LAUNCH(7, 4)
            DCRF_COOP_LAUNCH(8, 4)
        }
#undef DCRF_COOP_LAUNCH
        DCRF_LAUNCHED();
        return;
    }
    DCRF_DISPATCH_G(g, {
        // persistent grid of exactly one resident wave; rows are claimed dynamically
        if (long_rows) {
            static const int per_sm = resident_blocks_per_sm(splat_fast_kernel<G, 8, REF>);
            const int nb = (int)std::min<int64_t>(ceil_div(lat.M * g, kThreads), (int64_t)kNumSMs * per_sm);
            splat_fast_kernel<G, 8, REF><<<nb + kLongRowBlocks, kThreads, 0, s>>>(lat.csr_start.p, ents, q4, v4, (int)lat.M, g,
                                                                lat.row_counter.p, cap, splat_chunk(lat.M, g, nb),
                                                                lat.long_rows.p, lat.n_long.p);
        } else {
            static const int per_sm = resident_blocks_per_sm(splat_fast_kernel<G, kSplatBatch, REF>);
            const int nb = (int)std::min<int64_t>(ceil_div(lat.M * g, kThreads), (int64_t)kNumSMs * per_sm);
            splat_fast_kernel<G, kSplatBatch, REF><<<nb + kLongRowBlocks, kThreads, 0, s>>>(
                lat.csr_start.p, ents, q4, v4, (int)lat.M, g, lat.row_counter.p, cap, splat_chunk(lat.M, g, nb),
                lat.long_rows.p, lat.n_long.p);
        }
    });
    DCRF_LAUNCHED();
}

void launch_splat_fast(const Lattice &lat, const float *Q, float *val, int Lp, cudaStream_t s) {
    if (lat.M == 0) return;
    DCRF_REQUIRE(lat.table_mode != kTablesNone, DCRF_ESTATE, "packed lattice tables are missing");
    if (lat.table_mode == kTablesRef) launch_splat_packed<true>(lat, Q, val, Lp, s);
    else launch_splat_packed<false>(lat, Q, val, Lp, s);
}

void launch_blur(const Lattice &lat, int axis, const float *in, float *out, int Lp, bool seq,
                 cudaStream_t s) {
    if (lat.M == 0) return;
    const int g = Lp / 4;
    const int nb = ceil_div(lat.M * g, kBlurThreads * kBlurUnroll);
    const int2 *nbr = lat.neigh.p + (int64_t)axis * lat.M;
    ProfScope prof(DCRF_K_BLUR, lat.d, s);
    DCRF_DISPATCH_G(g, {
        if (seq) blur_kernel<G, true><<<nb, kBlurThreads, 0, s>>>(nbr, in, out, lat.M, g);
        else blur_kernel<G, false><<<nb, kBlurThreads, 0, s>>>(nbr, in, out, lat.M, g);
    });
    DCRF_LAUNCHED();
}

void launch_slice_softmax(const SliceArgs &a, const float *unary, float *Q, int64_t Ntot, int L, int Lp,
                          cudaStream_t s) {
    if (Ntot == 0) return;
    const int g = Lp / 4;
    const int nb = ceil_div(Ntot, rows_per_block(g));
    ProfScope prof(DCRF_K_SLICE, a.n_terms, s);
    const bool fast25 = a.n_terms == 2 && a.term[0].d == 2 && a.term[1].d == 5 && !a.seq;
    bool all_potts = !a.seq;
    for (int k = 0; k < a.n_terms; k++)
        all_potts = all_potts && a.term[k].compat_kind == DCRF_COMPAT_POTTS && a.term[k].ent != nullptr;
    // the packed-table kernels index rows with 32 bits: largest row index * g must stay below 2^32
    const bool idx32 = a.max_rows * (int64_t)g < ((int64_t)1 << 32) && Ntot * (int64_t)g < ((int64_t)1 << 31);
    const float4 *u4 = reinterpret_cast<const float4 *>(unary);
    float4 *q4 = reinterpret_cast<float4 *>(Q);
    if (a.fast == kSliceRef && all_potts && idx32) {
        if (fast25) {
            DCRF_DISPATCH_G(g, { slice_softmax_ref_kernel<G, 2, 5><<<nb, kThreads, 0, s>>>(a, u4, q4, (unsigned)Ntot, L, g); });
        } else {
            DCRF_DISPATCH_G(g, { slice_softmax_ref_generic_kernel<G><<<nb, kThreads, 0, s>>>(a, u4, q4, (unsigned)Ntot, L, g); });
        }
        DCRF_LAUNCHED();
        return;
    }
    if (a.fast == kSliceFma && all_potts && idx32) {
        if (a.n_terms == 0) {
            DCRF_DISPATCH_G(g, { softmax_unary_fast_kernel<G><<<nb, kThreads, 0, s>>>(u4, q4, (unsigned)Ntot, L, g); });
        } else if (fast25) {
            DCRF_DISPATCH_G(g, { slice_softmax_fast_kernel<G, 2, 5><<<nb, kThreads, 0, s>>>(a, u4, q4, (unsigned)Ntot, L, g); });
        } else {
            DCRF_DISPATCH_G(g, { slice_softmax_fast_generic_kernel<G><<<nb, kThreads, 0, s>>>(a, u4, q4, (unsigned)Ntot, L, g); });
        }
        DCRF_LAUNCHED();
        return;
    }
    DCRF_DISPATCH_G(g, {
        if (fast25) slice_softmax_kernel<G, true><<<nb, kThreads, 0, s>>>(a, unary, Q, Ntot, L, g);
        else slice_softmax_kernel<G, false><<<nb, kThreads, 0, s>>>(a, unary, Q, Ntot, L, g);
    });
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

void launch_kernel_norm(const Lattice &lat, int64_t N, int ntype, float *norm, cudaStream_t s) {
    if (N == 0 || lat.M == 0) return;
    DevBuf<float> a, b;
    a.alloc(lat.M, s);
    b.alloc(lat.M, s);
    const int nbm = ceil_div(lat.M, kThreads);
    norm_splat_kernel<<<nbm, kThreads, 0, s>>>(lat.csr_start.p, lat.csr_w.p, a.p, lat.M);
    DCRF_LAUNCHED();
    float *cur = a.p, *nxt = b.p;
    for (int j = 0; j <= lat.d; j++) {
        norm_blur_kernel<<<nbm, kThreads, 0, s>>>(lat.neigh.p + (int64_t)j * lat.M, cur, nxt, lat.M);
        DCRF_LAUNCHED();
        std::swap(cur, nxt);
    }
    const float alpha = 1.0f / (1.0f + powf(2.0f, (float)-lat.d));
    norm_slice_kernel<<<ceil_div(N, kThreads), kThreads, 0, s>>>(lat.offset.p, lat.bary.p, cur, lat.d, alpha, norm,
                                                              N, ntype);
    DCRF_LAUNCHED();
}

static int max_image_pixels(const BatchGeom &g) {
    int64_t m = 0;
    for (int b = 0; b < g.B; b++) m = std::max<int64_t>(m, g.pix_start[b + 1] - g.pix_start[b]);
    return (int)m;
}

void launch_ln_to_pm(const float *ln, float *pm, const BatchGeom &g, int L, int Lp, cudaStream_t s) {
    if (g.Ntot == 0) return;
    dim3 grid(ceil_div(max_image_pixels(g), kTP), g.B);
    const size_t smem = sizeof(float) * Lp * (kTP + 1);
    if (smem > 48 * 1024)
        DCRF_CUDA(cudaFuncSetAttribute(ln_to_pm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ln_to_pm_kernel<<<grid, kThreads, smem, s>>>(ln, pm, g.d_pix_start, L, Lp);
    DCRF_LAUNCHED();
}

void launch_pm_to_ln(const float *pm, float *ln, const BatchGeom &g, int L, int Lp, cudaStream_t s) {
    if (g.Ntot == 0) return;
    dim3 grid(ceil_div(max_image_pixels(g), kTP), g.B);
    const size_t smem = sizeof(float) * Lp * (kTP + 1);
    if (smem > 48 * 1024)
        DCRF_CUDA(cudaFuncSetAttribute(pm_to_ln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pm_to_ln_kernel<<<grid, kThreads, smem, s>>>(pm, ln, g.d_pix_start, L, Lp);
    DCRF_LAUNCHED();
}

// ---------------------------------------------------------------------------------------------
// Q -> (pixel, label) rows without padding = the (H, W, C) layout SEC / DSRG's crf_inference returns,
// optionally with the epilogue of their `crf` closure (/root/reference/03a_sec-dsrg/SEC.py:277-279):
//   ret[ret < min_prob] = min_prob;  ret /= np.sum(ret, axis=3, keepdims=True);  ret = np.log(ret)
// The sum follows NumPy's float32 pairwise order for n <= 128 (8 strided partial sums combined as
// ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), remainder added in sequence; plain loop below 8), so clamp,
// sum and quotient are bit-identical to NumPy; only logf differs from NumPy's SIMD log (<= 2 ulp).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float clamp_lo(float v, float lo) { return v < lo ? lo : v; }

__global__ void __launch_bounds__(kThreads) q_to_hwc_kernel(const float *__restrict__ pm, float *__restrict__ out,
                                                            int64_t total, int L, int Lp, float min_prob,
                                                            int renorm, int take_log) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= total) return;
    const int64_t p = i / L;
    const int l = (int)(i - p * L);
    const float *row = pm + p * Lp;
    float v = row[l];
    if (renorm) {
        v = clamp_lo(v, min_prob);
        float sum;
        if (L < 8) {
            sum = 0.f;
            for (int k = 0; k < L; k++) sum = __fadd_rn(sum, clamp_lo(row[k], min_prob));
        } else {
            float r[8];
#pragma unroll
            for (int j = 0; j < 8; j++) r[j] = clamp_lo(row[j], min_prob);
            int k = 8;
            for (; k < L - (L % 8); k += 8)
#pragma unroll
                for (int j = 0; j < 8; j++) r[j] = __fadd_rn(r[j], clamp_lo(row[k + j], min_prob));
            sum = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                            __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
            for (; k < L; k++) sum = __fadd_rn(sum, clamp_lo(row[k], min_prob));
        }
        v = __fdiv_rn(v, sum);
    }
    if (take_log) v = logf(v);
    out[i] = v;
}

void launch_q_to_hwc(const float *pm, float *out, int64_t Ntot, int L, int Lp, float min_prob, int renorm,
                     int take_log, cudaStream_t s) {
    if (Ntot == 0) return;
    const int64_t total = Ntot * L;
    q_to_hwc_kernel<<<ceil_div(total, kThreads), kThreads, 0, s>>>(pm, out, total, L, Lp, min_prob, renorm, take_log);
    DCRF_LAUNCHED();
}

void launch_argmax(const float *pm, int32_t *labels, int64_t Ntot, int L, int Lp, cudaStream_t s) {
    if (Ntot == 0) return;
    argmax_kernel<int32_t><<<ceil_div(Ntot, kThreads), kThreads, 0, s>>>(pm, labels, Ntot, L, Lp);
    DCRF_LAUNCHED();
}
void launch_argmax_u8(const float *pm, uint8_t *labels, int64_t Ntot, int L, int Lp, cudaStream_t s) {
    if (Ntot == 0) return;
    argmax_kernel<uint8_t><<<ceil_div(Ntot, kThreads), kThreads, 0, s>>>(pm, labels, Ntot, L, Lp);
    DCRF_LAUNCHED();
}

// test hook: y[i] = ExpfRef(x[i]) (warp-collective, so whole warps run and only the store is guarded)
__global__ void __launch_bounds__(kThreads) expf_ref_kernel(const float *__restrict__ x, float *__restrict__ y,
                                                            int64_t n) {
    const ExpfRef ex;
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const float r = ex(i < n ? x[i] : 0.f);
    if (i < n) y[i] = r;
}
void launch_expf_ref(const float *x, float *y, int64_t n, cudaStream_t s) {
    if (n == 0) return;
    expf_ref_kernel<<<ceil_div(n, kThreads), kThreads, 0, s>>>(x, y, n);
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

void launch_unary_from_probs(const void *probs, int is_f64, float *pm, const BatchGeom &g, int L, int Lp,
                             double scale, double clip_lo, int has_clip, cudaStream_t s) {
    if (g.Ntot == 0) return;
    dim3 grid(ceil_div(max_image_pixels(g), kTP), g.B);
    const size_t smem = sizeof(float) * Lp * (kTP + 1);
    if (is_f64) {
        if (smem > 48 * 1024)
            DCRF_CUDA(cudaFuncSetAttribute(unary_from_probs_kernel<double>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        unary_from_probs_kernel<double><<<grid, kThreads, smem, s>>>((const double *)probs, pm, g.d_pix_start, L, Lp,
                                                                    scale, clip_lo, has_clip);
    } else {
        if (smem > 48 * 1024)
            DCRF_CUDA(cudaFuncSetAttribute(unary_from_probs_kernel<float>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        unary_from_probs_kernel<float><<<grid, kThreads, smem, s>>>((const float *)probs, pm, g.d_pix_start, L, Lp,
                                                                   scale, clip_lo, has_clip);
    }
    DCRF_LAUNCHED();
}

void launch_unary_from_logits(const float *feat, float *pm, int64_t Ntot, int L, int Lp, int use_log,
                              cudaStream_t s) {
    if (Ntot == 0) return;
    const int g = Lp / 4;
    const int nb = ceil_div(Ntot, rows_per_block(g));
    DCRF_DISPATCH_G(g, { unary_from_logits_kernel<G><<<nb, kThreads, 0, s>>>(feat, pm, Ntot, L, g, use_log); });
    DCRF_LAUNCHED();
}

void launch_unary_from_labels(const int32_t *labels, float *pm, int64_t Ntot, int L, int Lp, float n_energy,
                              float p_energy, float unsure_energy, int zero_unsure, int *bad, cudaStream_t s) {
    if (Ntot == 0) return;
    unary_from_labels_kernel<<<ceil_div(Ntot * Lp, kThreads), kThreads, 0, s>>>(
        labels, pm, Ntot, L, Lp, n_energy, p_energy, unsure_energy, zero_unsure, bad);
    DCRF_LAUNCHED();
}


// persistent path: returns false when the configuration is not covered (caller falls back to the
// launch-per-phase path)
template <int G, bool REF>
static bool launch_persistent_g(const MfArgs &a, cudaStream_t s) {
    static int per_sm = 0;
    if (!per_sm) {
        int n = 0;
        DCRF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, mean_field_persistent_kernel<G, REF>, kThreads, 0));
        const char *e = getenv("DCRF_PERSISTENT_CTAS_PER_SM");
        const int want = e ? atoi(e) : 3;
        per_sm = std::max(1, std::min(n, want > 0 ? want : n));
    }
    void *params[] = {(void *)&a};
    DCRF_CUDA(cudaLaunchCooperativeKernel((const void *)mean_field_persistent_kernel<G, REF>, dim3(kNumSMs * per_sm),
                                          dim3(kThreads), params, 0, s));
    DCRF_LAUNCHED();
    return true;
}

bool launch_mean_field_persistent(const Lattice *const *lats, float *const *valA, float *const *valB,
                                  const SliceArgs &slice, const float *unary, float *Q, int64_t Ntot, int L, int Lp,
                                  int n_iter, int *counters, cudaStream_t s) {
    const int g = Lp / 4;
    if (g < 1 || g > 8 || slice.n_terms < 1) return false;
    const bool ref = slice.fast == kSliceRef;
    if (!ref && slice.fast != kSliceFma) return false;
    MfArgs a;
    memset(&a, 0, sizeof(a));
    a.slice = slice;
    for (int k = 0; k < slice.n_terms; k++) {
        const Lattice &lat = *lats[k];
        if (lat.M <= 0 || lat.table_mode != (ref ? kTablesRef : kTablesFma)) return false;
        MfTerm &t = a.term[k];
        t.csr_start = lat.csr_start.p;
        t.csr_ent = ref ? (const void *)lat.csr_ent4.p : (const void *)lat.csr_ent.p;
        t.neigh = lat.neigh.p;
        t.long_rows = lat.long_rows.p;
        t.n_long = lat.n_long.p;
        t.valA = valA[k];
        t.valB = valB[k];
        t.M = (int)lat.M;
        t.d = lat.d;
        t.long_cap = lat.long_row_cap;
        t.long_hint = lat.E >= 16 * lat.M ? 1 : 0;
        t.short_rows = splat_uses_short_rows(lat) ? 1 : 0;
    }
    a.unary4 = reinterpret_cast<const float4 *>(unary);
    a.Q4 = reinterpret_cast<float4 *>(Q);
    a.Ntot = (unsigned)Ntot;
    a.L = L;
    a.g = g;
    a.n_iter = n_iter;
    a.counters = counters;
    DCRF_CUDA(cudaMemsetAsync(counters, 0, sizeof(int) * std::max(1, 2 * n_iter * slice.n_terms), s));
    switch (g) {
#define DCRF_PERSIST_CASE(GG)                                        \
    case GG:                                                         \
        return ref ? launch_persistent_g<GG, true>(a, s) : launch_persistent_g<GG, false>(a, s);
        DCRF_PERSIST_CASE(1)
        DCRF_PERSIST_CASE(2)
        DCRF_PERSIST_CASE(3)
        DCRF_PERSIST_CASE(4)
        DCRF_PERSIST_CASE(5)
        DCRF_PERSIST_CASE(6)
        DCRF_PERSIST_CASE(7)
        DCRF_PERSIST_CASE(8)
#undef DCRF_PERSIST_CASE
    }
    return false;
}

}  // namespace dcrf
