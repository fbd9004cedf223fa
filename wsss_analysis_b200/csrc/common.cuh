// common.cuh -- shared declarations of the dcrf_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>
#include <vector>

#include "../../include/dcrf_b200.h"

namespace dcrf {

constexpr int kMaxD = 7;        // lattice feature dimension supported (rank nibbles pack into 32 bits)
constexpr int kMaxPairwise = 4; // pairwise kernels fused into one slice launch
constexpr int kNumSMs = 148;    // B200

extern std::atomic<int64_t> g_launches;
void set_error(const std::string &msg);

struct Error {
    int code;
    std::string msg;
};

#define DCRF_CUDA(expr)                                                                        \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            throw ::dcrf::Error{DCRF_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)}; \
    } while (0)

#define DCRF_REQUIRE(cond, code, message)                  \
    do {                                                   \
        if (!(cond)) throw ::dcrf::Error{(code), (message)}; \
    } while (0)

// count + check a kernel launch
#define DCRF_LAUNCHED()                                \
    do {                                               \
        ::dcrf::g_launches.fetch_add(1);               \
        DCRF_CUDA(cudaGetLastError());                 \
    } while (0)

// every copy between caller / host memory and the device goes through these two (byte counters behind
// dcrf_copy_count: tests assert that device-tensor calls move no payload over PCIe)
extern std::atomic<int64_t> g_h2d_bytes, g_d2h_bytes;
inline cudaError_t copy_h2d(void *dst, const void *src, size_t bytes, cudaStream_t s) {
    g_h2d_bytes.fetch_add((int64_t)bytes);
    return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s);
}
inline cudaError_t copy_d2h(void *dst, const void *src, size_t bytes, cudaStream_t s) {
    g_d2h_bytes.fetch_add((int64_t)bytes);
    return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s);
}

// One memory pool PER STREAM (created on first use, destroyed with dcrf_stream_destroy or at exit).
// Every block is then allocated, freed and recycled in the order of ONE stream: no cross-stream
// reuse, hence no hidden inter-stream dependencies and no pool growth after the first batches.
// (A shared pool re-grows whenever a block freed on another stream is still in flight, and growing
// a pool maps device memory, which stalls for 10-600 ms when the GPU is busy -- measured.)
cudaMemPool_t stream_pool(cudaStream_t stream);
void stream_pool_release(cudaStream_t stream);
// DCRF_TRACE=1: report host-side CUDA calls that block for more than 2 ms (diagnostics)
double trace_now();
void trace_slow(const char *what, double t0, size_t bytes);

// Stream-ordered device buffer (cudaMallocFromPoolAsync: allocation is cheap after warm-up).
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    cudaStream_t s = nullptr;
    DevBuf() {}
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n), s(o.s) { o.p = nullptr; o.n = 0; }
    DevBuf &operator=(DevBuf &&o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; s = o.s; o.p = nullptr; o.n = 0; }
        return *this;
    }
    void alloc(size_t count, cudaStream_t stream) {
        release();
        s = stream;
        n = count;
        if (count) {
            const double t0 = trace_now();
            DCRF_CUDA(cudaMallocFromPoolAsync((void **)&p, count * sizeof(T), stream_pool(stream), stream));
            trace_slow("cudaMallocFromPoolAsync", t0, count * sizeof(T));
        }
    }
    void release() {
        if (p) cudaFreeAsync(p, s);
        p = nullptr;
        n = 0;
    }
    ~DevBuf() { release(); }
};

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------
// per-kernel-class device timing with CUDA events on the launching stream (bench.py's roofline)
// ---------------------------------------------------------------------------------------------
struct Profiler {
    struct Rec { int cls, tag; cudaEvent_t a, b; };
    bool on = false;
    std::vector<Rec> recs;
    std::vector<cudaEvent_t> pool;
    cudaEvent_t get() {
        cudaEvent_t e;
        if (!pool.empty()) { e = pool.back(); pool.pop_back(); return e; }
        DCRF_CUDA(cudaEventCreate(&e));
        return e;
    }
    ~Profiler() {
        for (auto &r : recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
        for (auto e : pool) cudaEventDestroy(e);
    }
};
extern thread_local Profiler *t_prof;  // set by the API entry points while a handle is being driven
struct ProfScope {
    Profiler *p;
    cudaStream_t s;
    cudaEvent_t b = nullptr;
    ProfScope(int cls, int tag, cudaStream_t stream) : p(t_prof), s(stream) {
        if (!p || !p->on) { p = nullptr; return; }
        cudaEvent_t a = p->get();
        b = p->get();
        DCRF_CUDA(cudaEventRecord(a, s));
        p->recs.push_back({cls, tag, a, b});
    }
    ~ProfScope() {
        if (p) cudaEventRecord(b, s);
    }
};

// ---------------------------------------------------------------------------------------------
// lattice construction (lattice_build.cu)
// ---------------------------------------------------------------------------------------------
struct FeatureSpec {
    int mode;        // 0 = gaussian (x/sx, y/sy), 1 = bilateral (+ r/sr, g/sg, b/sb), 2 = explicit (d, N)
    int d;
    float s[5];      // sx, sy, sr, sg, sb
    const uint8_t *rgb;    // device, concatenated (N, 3)            (mode 1)
    const float *features; // device, row-major (d, N)               (mode 2)
};

struct BatchGeom {
    int B;
    int64_t Ntot;
    const int *d_w, *d_h;          // device [B]
    const int *d_pix_start;        // device [B+1]
    std::vector<int> w, h;
    std::vector<int64_t> pix_start; // host [B+1]
};

// Result of building one lattice over the whole batch; vertex ids are GLOBAL over the batch and,
// inside each image, follow the reference first-occurrence numbering.
struct Lattice {
    int d = 0;
    int64_t M = 0;                 // total vertices
    int64_t E = 0;                 // Ntot * (d+1) entries
    std::vector<int64_t> vert_start; // host [B+1]
    // host staging of small per-image tables uploaded asynchronously during the build (kept alive here
    // instead of synchronising the stream before they go out of scope)
    std::vector<int64_t> h_tab_start, h_tab2_start;
    std::vector<int> h_tab_mask, h_tab2_mask;
    std::vector<int32_t> h_seg, h_tile, h_rep;
    DevBuf<int32_t> offset;        // [E] vertex id of entry e = p*(d+1)+r
    DevBuf<float> bary;            // [E]
    DevBuf<int2> neigh;            // [(d+1) * M] (n1, n2), -1 = absent
    DevBuf<int16_t> vkeys;         // [M * 8] key of each vertex (first d shorts used)
    DevBuf<int32_t> csr_start;     // [M+1] rows of the transposed incidence (splat as a gather)
    DevBuf<int32_t> csr_pix;       // [E] pixel of each sorted entry (ascending entry order per row)
    DevBuf<float> csr_w;           // [E] barycentric weight of each sorted entry
    // packed tables of the fast path (one 64-bit load per entry instead of two 32-bit loads)
    DevBuf<int2> ent;              // [E] (vertex id, weight * post-norm[pixel] bits) of entry e
    DevBuf<int2> csr_ent;          // [E] (pixel, weight * pre-norm[pixel] bits) of each sorted entry
    // reference-association tables (arithmetic identical to the sequential CPU evaluation):
    // ent = (vertex id, bary * alpha), csr_ent4 = (pixel, weight, pre-norm[pixel] or 1, 0)
    DevBuf<int4> csr_ent4;         // [E]
    int table_mode = 0;            // kTables*: which of the packed tables are valid
    int long_row_cap = 0;          // rows longer than this are cut (tail: splat_tail_warp_kernel); INT_MAX = never
    DevBuf<int> row_counter;       // [1] dynamic row-chunk dispenser of the fast splat
    DevBuf<int32_t> long_rows;     // rows with more than kSplatLongRow entries (tail summed by a whole CTA)
    DevBuf<int> n_long;            // [1] their number (device side)
};

// A build has two halves around its one host synchronisation (the vertex count sizes the arrays of the
// second half).  build_lattice() runs both on one stream; small problems run the first halves of all
// their kernels' lattices concurrently on separate streams (api.cu) -- the temporaries live here.
struct BuildState {
    DevBuf<int4> rec_rem;
    DevBuf<uint32_t> rec_rank;
    DevBuf<int32_t> slot_of, pscan, d_vert_start;
    DevBuf<uint8_t> mask8;
    std::vector<int32_t> h_vs_own;  // destination of the vertex starts when the caller gives no pinned block
    int32_t *h_vs = nullptr;        // [B+1] vertex starts: valid once the stream reached the end of begin()
};
// first half: point, hash, first-occurrence masks, scan, D2H of the per-image vertex starts into
// `pinned_vs` (page-locked, asynchronous) or into st.h_vs_own (pageable: the copy itself blocks)
void build_lattice_begin(const BatchGeom &g, const FeatureSpec &f, Lattice &out, BuildState &st, int32_t *pinned_vs,
                         cudaStream_t stream);
// second half, once the stream has reached the end of the first: ids, neighbours, CSR rows
void build_lattice_finish(const BatchGeom &g, const FeatureSpec &f, Lattice &out, BuildState &st,
                          cudaStream_t stream);
void build_lattice(const BatchGeom &g, const FeatureSpec &f, Lattice &out, cudaStream_t stream);
// `one`: lattice (packed tables included) over the DISTINCT image sizes `og` of the batch `g`; image b of
// the batch is unique image src[b].  `out`: the same lattice for every image of the batch laid back to
// back, pixel / vertex / entry ids shifted per image.
void launch_replicate_lattice(const Lattice &one, const BatchGeom &og, const float *norm_one, const BatchGeom &g,
                              const std::vector<int> &src, Lattice &out, float *norm_out, cudaStream_t s);

// ---------------------------------------------------------------------------------------------
// filtering + mean field (filter.cu)
// ---------------------------------------------------------------------------------------------
// one pairwise term as seen by the fused slice/softmax kernel
struct SliceTerm {
    const int32_t *offset;
    const float *bary;
    const float *val;     // blurred lattice values [M * Lp]
    const float *norm;    // [Ntot] or nullptr (no post-scaling)
    const float *compat;  // device: Potts -> unused, diagonal -> [L], matrix -> [L*L]
    float potts_w;
    float alpha;
    int d;
    int compat_kind;
    const int2 *ent;      // packed (vertex id, weight) per entry (fast path)
};
struct SliceArgs {
    SliceTerm term[kMaxPairwise];
    int n_terms;
    int seq;  // value_size <= 2 association (A.4)
    int fast; // kSlice*: which family of kernels to use
    int64_t max_rows; // largest lattice vertex count among the terms (32-bit row indexing guard)
};
// arithmetic modes of the per-iteration kernels (DCRF_OPT_EXACT_ARITHMETIC values)
constexpr int kArithFma = 0;     // FMA accumulation, normalisation folded into the packed weights
constexpr int kArithRef = 1;     // the specification's association on packed tables (bit-identical to the
                                 // sequential CPU evaluation except for the tails of very long splat rows)
constexpr int kArithStrict = 2;  // kArithRef without the long-row split
constexpr int kTablesNone = 0, kTablesFma = 1, kTablesRef = 2;
// SliceArgs::fast: literal per-entry kernels (value_size <= 2, diagonal / matrix compatibility), FMA
// kernels on packed tables, reference-association kernels on packed tables
constexpr int kSliceLiteral = 0, kSliceFma = 1, kSliceRef = 2;

// values <- splat of (pre ? norm (.) Q : Q)       (A.4 splat, A.5 pre-scaling)
void launch_splat(const Lattice &lat, const float *Q, const float *norm_pre, float *val, int Lp,
                  cudaStream_t s);
// fast path: packed tables; splat weights pre-multiplied by the pre-normalisation, slice weights by
// the post-normalisation (the fast kernels never read the norm vector), FMA accumulation
void launch_pack_fast_tables(Lattice &lat, const float *norm_pre, const float *norm_post, cudaStream_t s);
// reference-association tables; long_row_cap as in Lattice
void launch_pack_ref_tables(Lattice &lat, const float *norm_pre, int long_row_cap, cudaStream_t s);
void launch_splat_fast(const Lattice &lat, const float *Q, float *val, int Lp, cudaStream_t s);
void launch_find_long_rows(Lattice &lat, cudaStream_t s);
// out <- in + 0.5 (in[n1] + in[n2]) along axis j  (A.4 blur)
void launch_blur(const Lattice &lat, int axis, const float *in, float *out, int Lp, bool seq,
                 cudaStream_t s);
// Q <- softmax_L( -U - sum_k compat_k( norm_k (.) slice_k ) )   (A.4 slice, A.5, A.6, A.7)
void launch_slice_softmax(const SliceArgs &a, const float *unary, float *Q, int64_t Ntot, int L,
                          int Lp, cudaStream_t s);
// `inference(n)` of a small problem as ONE cooperative launch (grid barriers between the phases); all
// terms Potts, packed tables of the arithmetic in slice.fast.  counters: device int[2 * n_iter * n_terms].
// Returns false when the configuration is not covered (nothing launched).
bool launch_mean_field_persistent(const Lattice *const *lats, float *const *valA, float *const *valB,
                                  const SliceArgs &slice, const float *unary, float *Q, int64_t Ntot, int L, int Lp,
                                  int n_iter, int *counters, cudaStream_t s);
// plain slice of one lattice into a pixel-major buffer: out[p] = sum_r w v alpha (seq selects the
// value_size<=2 association)
void launch_slice_plain(const Lattice &lat, const float *val, float *out, int64_t Ntot, int Lp,
                        bool seq, cudaStream_t s);
// norm[p] = f( K 1 ) for the N pixels of `lat` (A.5), scalar value_size = 1 path
void launch_kernel_norm(const Lattice &lat, int64_t N, int ntype, float *norm, cudaStream_t s);
// (L, N_b) row-major blocks  <->  (Ntot, Lp) pixel-major
void launch_ln_to_pm(const float *ln, float *pm, const BatchGeom &g, int L, int Lp, cudaStream_t s);
void launch_pm_to_ln(const float *pm, float *ln, const BatchGeom &g, int L, int Lp, cudaStream_t s);
// unary construction on the GPU (wrappers' NumPy glue): straight into the pixel-major unary buffer
void launch_unary_from_probs(const void *probs, int is_f64, float *pm, const BatchGeom &g, int L, int Lp,
                             double scale, double clip_lo, int has_clip, cudaStream_t s);
void launch_unary_from_logits(const float *feat, float *pm, int64_t Ntot, int L, int Lp, int use_log,
                              cudaStream_t s);
void launch_unary_from_labels(const int32_t *labels, float *pm, int64_t Ntot, int L, int Lp, float n_energy,
                              float p_energy, float unsure_energy, int zero_unsure, int *bad, cudaStream_t s);
void launch_argmax(const float *pm, int32_t *labels, int64_t Ntot, int L, int Lp, cudaStream_t s);
void launch_argmax_u8(const float *pm, uint8_t *labels, int64_t Ntot, int L, int Lp, cudaStream_t s);
// out[p * L + l] = Q (pixel-major without padding, the (H, W, C) layout), optionally clamped to
// >= min_prob and renormalised over labels in NumPy's summation order, optionally log
void launch_q_to_hwc(const float *pm, float *out, int64_t Ntot, int L, int Lp, float min_prob, int renorm,
                     int take_log, cudaStream_t s);
// y = expf(x) as evaluated by the reference-association softmax (softmax_ref.cuh); test hook
void launch_expf_ref(const float *x, float *y, int64_t n, cudaStream_t s);
// deterministic double-precision KL terms
void launch_kl(const float *Q, const float *unary, const float *const *pair_out, int n_pair,
               int64_t Ntot, int L, int Lp, double *out, cudaStream_t s);
// pairwise_out <- compat( norm (.) slice ) for one term (used by klDivergence)
void launch_slice_pairwise_only(const SliceTerm &t, float *out, int64_t Ntot, int L, int Lp,
                                cudaStream_t s);

// ---------------------------------------------------------------------------------------------
// device-wide primitives (primitives.cu)
// ---------------------------------------------------------------------------------------------
// out[i] = exclusive prefix sum of in[0..i), out[n] = total.  in/out may alias.  int32.
void exclusive_scan_i32(const int32_t *in, int32_t *out, int64_t n, cudaStream_t s);
// Stable LSD radix sort of interleaved (key, value) pairs inside consecutive segments, by
// (key - key_base[segment]) on `local_bits` bits.  Returns 0 when the sorted pairs ended in pairs_a, 1
// when they ended in pairs_b.  h_seg / h_tile: caller-owned staging of two small uploads.
// one stable pass on the high bits + a counting sort per bucket that writes the CSR arrays; false = not
// applicable (vertex ids of an image need more than 23 bits): use segmented_radix_sort_pairs
bool bucket_sort_to_csr(uint2 *pairs_a, uint2 *pairs_b, const std::vector<int64_t> &seg_start,
                        const int32_t *d_key_base, int local_bits, const float *bary, int d1, int64_t E, int64_t M,
                        int32_t *csr_start, int32_t *csr_pix, float *csr_w, int prof_tag, cudaStream_t s,
                        std::vector<int32_t> &h_seg, std::vector<int32_t> &h_tile);
int segmented_radix_sort_pairs(uint2 *pairs_a, uint2 *pairs_b, const std::vector<int64_t> &seg_start,
                               const int32_t *d_key_base, int local_bits, cudaStream_t s,
                               std::vector<int32_t> &h_seg, std::vector<int32_t> &h_tile);

}  // namespace dcrf
