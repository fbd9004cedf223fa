// primitives.cu -- hand-written device-wide exclusive scan and stable LSD radix sort (sm_100a).
// Used only by lattice construction (vertex numbering and the splat's transposed incidence rows).
// Everything here is integer work: results are bit-deterministic.
#include <algorithm>

#include "common.cuh"

namespace dcrf {

namespace {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;  // 4096

__device__ __forceinline__ int warp_inclusive_scan(int v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// exclusive scan of one value per thread over the block; returns exclusive prefix, *total = block sum
__device__ __forceinline__ int block_exclusive_scan(int v, int *total) {
    __shared__ int warp_sums[kScanThreads / 32];
    __shared__ int block_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = warp_inclusive_scan(v);
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int s = lane < kScanThreads / 32 ? warp_sums[lane] : 0;
        int sinc = warp_inclusive_scan(s);
        if (lane < kScanThreads / 32) warp_sums[lane] = sinc - s;
        if (lane == kScanThreads / 32 - 1) block_total = sinc;
    }
    __syncthreads();
    int r = inc - v + warp_sums[warp];
    *total = block_total;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const int32_t *__restrict__ in,
                                                                   int32_t *__restrict__ sums,
                                                                   int64_t n) {
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    int s = 0;
    if (base + kScanItems <= n) {
        const int4 *p = reinterpret_cast<const int4 *>(in + base);
#pragma unroll
        for (int i = 0; i < kScanItems / 4; i++) {
            int4 v = p[i];
            s += v.x + v.y + v.z + v.w;
        }
    } else {
        for (int i = 0; i < kScanItems; i++)
            if (base + i < n) s += in[base + i];
    }
    int total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// carry == nullptr: single-tile scan (n <= kScanTile, one block)
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const int32_t *__restrict__ in,
                                                                  int32_t *__restrict__ out,
                                                                  const int32_t *__restrict__ carry,
                                                                  int64_t n, int write_total) {
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    int v[kScanItems];
    if (base + kScanItems <= n) {
        const int4 *p = reinterpret_cast<const int4 *>(in + base);
#pragma unroll
        for (int i = 0; i < kScanItems / 4; i++) {
            int4 t = p[i];
            v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < kScanItems; i++) v[i] = (base + i < n) ? in[base + i] : 0;
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) s += v[i];
    int total;
    int run = block_exclusive_scan(s, &total) + (carry ? carry[blockIdx.x] : 0);
    // in and out may alias: every thread has read all of its inputs before any write of its range
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        int t = v[i];
        v[i] = run;
        run += t;
    }
    if (base + kScanItems <= n) {
        int4 *q = reinterpret_cast<int4 *>(out + base);
#pragma unroll
        for (int i = 0; i < kScanItems / 4; i++)
            q[i] = make_int4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else {
#pragma unroll
        for (int i = 0; i < kScanItems; i++)
            if (base + i < n) out[base + i] = v[i];
    }
    if (write_total && blockIdx.x == gridDim.x - 1 && threadIdx.x == kScanThreads - 1) out[n] = run;
}

void scan_rec(const int32_t *in, int32_t *out, int64_t n, bool write_total, cudaStream_t s) {
    const int nb = ceil_div(n, kScanTile);
    if (nb <= 1) {
        scan_apply_kernel<<<1, kScanThreads, 0, s>>>(in, out, nullptr, n, write_total ? 1 : 0);
        DCRF_LAUNCHED();
        return;
    }
    DevBuf<int32_t> sums;
    sums.alloc((size_t)nb + 1, s);
    scan_reduce_kernel<<<nb, kScanThreads, 0, s>>>(in, sums.p, n);
    DCRF_LAUNCHED();
    scan_rec(sums.p, sums.p, nb, false, s);
    scan_apply_kernel<<<nb, kScanThreads, 0, s>>>(in, out, sums.p, n, write_total ? 1 : 0);
    DCRF_LAUNCHED();
}

// ---------------- radix sort ----------------
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortRounds = 16;                                   // 32-key rounds per warp
constexpr int kSortTile = kSortThreads * kSortRounds;             // 4096 keys per block
constexpr int kSortWarpChunk = 32 * kSortRounds;                  // 512 contiguous keys per warp


// ---------------- segmented radix sort ----------------
// Sorts (key, value) pairs inside S consecutive segments (the images of a batch) by
// (key - key_base[segment]), stable, LSD.  Keys of a segment are vertex ids local to one image, so
// they need ~17-20 bits instead of the ~22-24 bits of batch-global ids: two passes of <= 10 bits
// instead of three or four of 8.  Every tile of kSortTile pairs belongs to exactly one segment; the
// histogram is laid out [segment][digit][tile of the segment], so that ONE exclusive scan over it
// yields the global output position of every (segment, digit, tile) bucket.
constexpr int kSegMaxDigitBits = 10;
constexpr int kSegMaxRadix = 1 << kSegMaxDigitBits;

struct SegInfo {
    const int32_t *seg_start;   // [S+1] first pair of each segment
    const int32_t *key_base;    // [S]   subtracted from the keys of the segment
    const int32_t *tile_base;   // [S+1] first tile of each segment
    int S;
};

__device__ __forceinline__ int seg_of_tile(const SegInfo &si, int tile) {
    int lo = 0, hi = si.S - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (si.tile_base[mid] <= tile) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}

__global__ void __launch_bounds__(kSortThreads) seg_radix_hist_kernel(const uint2 *__restrict__ pairs,
                                                                      int32_t *__restrict__ hist, SegInfo si,
                                                                      int shift, int digit_bits) {
    __shared__ int h[kSegMaxRadix];
    const int radix = 1 << digit_bits;
    for (int i = threadIdx.x; i < radix; i += kSortThreads) h[i] = 0;
    __syncthreads();
    const int seg = seg_of_tile(si, blockIdx.x);
    const int t = blockIdx.x - si.tile_base[seg];
    const int tiles = si.tile_base[seg + 1] - si.tile_base[seg];
    const int64_t base = (int64_t)si.seg_start[seg] + (int64_t)t * kSortTile;
    const int64_t end = si.seg_start[seg + 1];
    const uint32_t kb = (uint32_t)si.key_base[seg];
#pragma unroll
    for (int i = 0; i < kSortRounds; i++) {
        const int64_t idx = base + (int64_t)i * kSortThreads + threadIdx.x;
        if (idx < end) atomicAdd(&h[((pairs[idx].x - kb) >> shift) & (radix - 1)], 1);
    }
    __syncthreads();
    int32_t *hs = hist + (int64_t)radix * si.tile_base[seg];
    for (int i = threadIdx.x; i < radix; i += kSortThreads) hs[(int64_t)i * tiles + t] = h[i];
}

__global__ void __launch_bounds__(kSortThreads, 4) seg_radix_scatter_kernel(
    const uint2 *__restrict__ pairs_in, uint2 *__restrict__ pairs_out, const int32_t *__restrict__ hist_scanned,
    SegInfo si, int shift, int digit_bits) {
    __shared__ int cnt[kSortWarps][kSegMaxRadix];
    const int radix = 1 << digit_bits;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = lane; i < radix; i += 32) cnt[warp][i] = 0;
    __syncwarp();
    const int seg = seg_of_tile(si, blockIdx.x);
    const int t = blockIdx.x - si.tile_base[seg];
    const int tiles = si.tile_base[seg + 1] - si.tile_base[seg];
    const int64_t end = si.seg_start[seg + 1];
    const uint32_t kb = (uint32_t)si.key_base[seg];
    const int64_t wbase = (int64_t)si.seg_start[seg] + (int64_t)t * kSortTile + (int64_t)warp * kSortWarpChunk;
    // pass A keeps only the digits (one byte-pair each); the pairs are read again in pass B from L2 -- holding
    // all 16 pairs in registers cost 102 registers per thread and a quarter of the SM's warp slots
    uint32_t dig[kSortRounds / 2];
#pragma unroll
    for (int r = 0; r < kSortRounds; r++) {
        const int64_t idx = wbase + r * 32 + lane;
        const bool valid = idx < end;
        const uint32_t key = valid ? pairs_in[idx].x : 0u;
        const int digit = valid ? (int)(((key - kb) >> shift) & (radix - 1)) : (kSegMaxRadix + lane);
        if (r & 1) dig[r / 2] |= (uint32_t)digit << 16;
        else dig[r / 2] = (uint32_t)digit;
        const unsigned peers = __match_any_sync(0xffffffffu, digit);
        if (valid && (peers & ((1u << lane) - 1)) == 0) cnt[warp][digit] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    const int32_t *hs = hist_scanned + (int64_t)radix * si.tile_base[seg];
    for (int digit = threadIdx.x; digit < radix; digit += kSortThreads) {
        int run = hs[(int64_t)digit * tiles + t];
#pragma unroll
        for (int w = 0; w < kSortWarps; w++) {
            const int c = cnt[w][digit];
            cnt[w][digit] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kSortRounds; r++) {
        const int64_t idx = wbase + r * 32 + lane;
        const bool valid = idx < end;
        const uint2 kv = valid ? pairs_in[idx] : make_uint2(0u, 0u);
        const int digit = (int)((dig[r / 2] >> ((r & 1) * 16)) & 0xffffu);  // (kSegMaxRadix + lane for invalid lanes)
        const unsigned peers = __match_any_sync(0xffffffffu, digit);
        const int rank = __popc(peers & ((1u << lane) - 1));
        if (valid) pairs_out[cnt[warp][digit] + rank] = kv;  // one 8-byte scatter per pair
        __syncwarp();
        if (valid && rank == 0) cnt[warp][digit] += __popc(peers);
        __syncwarp();
    }
}

}  // namespace

int segmented_radix_sort_pairs(uint2 *pairs_a, uint2 *pairs_b, const std::vector<int64_t> &seg_start,
                               const int32_t *d_key_base, int local_bits, cudaStream_t s,
                               std::vector<int32_t> &h_seg, std::vector<int32_t> &h_tile) {
    const int S = (int)seg_start.size() - 1;
    if (S <= 0 || seg_start[S] <= 0) return 0;
    int passes = (local_bits + kSegMaxDigitBits - 1) / kSegMaxDigitBits;
    if (passes < 1) passes = 1;
    const int digit_bits = std::max(1, (local_bits + passes - 1) / passes);
    const int radix = 1 << digit_bits;
    h_seg.assign(S + 1, 0);   // staging of asynchronous uploads: owned by the caller, outlives them
    h_tile.assign(S + 1, 0);
    for (int i = 0; i <= S; i++) h_seg[i] = (int32_t)seg_start[i];
    for (int i = 0; i < S; i++) h_tile[i + 1] = h_tile[i] + ceil_div(seg_start[i + 1] - seg_start[i], kSortTile);
    const int total_tiles = h_tile[S];
    DevBuf<int32_t> d_seg, d_tile, hist;
    d_seg.alloc(S + 1, s);
    d_tile.alloc(S + 1, s);
    DCRF_CUDA(copy_h2d(d_seg.p, h_seg.data(), sizeof(int32_t) * (S + 1), s));
    DCRF_CUDA(copy_h2d(d_tile.p, h_tile.data(), sizeof(int32_t) * (S + 1), s));
    hist.alloc((size_t)radix * total_tiles + 1, s);
    SegInfo si{d_seg.p, d_key_base, d_tile.p, S};
    uint2 *ki = pairs_a, *ko = pairs_b;
    for (int p = 0; p < passes; p++) {
        const int shift = p * digit_bits;
        seg_radix_hist_kernel<<<total_tiles, kSortThreads, 0, s>>>(ki, hist.p, si, shift, digit_bits);
        DCRF_LAUNCHED();
        scan_rec(hist.p, hist.p, (int64_t)radix * total_tiles, false, s);
        seg_radix_scatter_kernel<<<total_tiles, kSortThreads, 0, s>>>(ki, ko, hist.p, si, shift, digit_bits);
        DCRF_LAUNCHED();
        uint2 *t = ki;
        ki = ko;
        ko = t;
    }
    return passes & 1;
}

namespace {

// ---------------- bucket sort straight into the CSR rows ----------------
// The splat's rows are the entries sorted by (image, vertex id), stable in the entry index.  Instead of
// two or three LSD passes plus a finalising pass, ONE stable radix pass on the HIGH bits of the local
// vertex id groups the pairs into buckets of 2^lo_bits consecutive vertices (a few thousand entries);
// a CTA per bucket then counting-sorts its bucket on the low bits and writes the CSR arrays directly:
// row starts from the scanned counts, (pixel, weight) of every entry at its final position.  The second
// pass reads a contiguous bucket and writes inside the bucket's own output range -- no global scatter,
// no second histogram / scan, no separate finalising kernel.
// Stability of the in-bucket placement: warp w of the CTA owns the w-th contiguous slice of the bucket
// and a private set of per-vertex cursors; the cursors of (vertex, warp) start where the entries of
// that vertex in the slices before w end, so every warp places its slice in order on its own, rank
// among equal vertices inside a 32-entry round from __match_any_sync in lane order.  No block barrier
// inside the placement loop.  Shared memory: warps * (vertices of the CTA) cursors, at most 32 KB.
constexpr int kBucketMaxLoBits = 13;  // up to 2^13 vertices per bucket (vertex ids of an image: up to 23 bits)
constexpr int kBucketCtaBits = 10;    // up to 2^10 vertices per CTA: 8 warps x 1024 cursors = 32 KB
constexpr int kBucketMaxThreads = 256;

__global__ void __launch_bounds__(kBucketMaxThreads) bucket_csr_kernel(
    const uint2 *__restrict__ pairs, const int32_t *__restrict__ hist_scanned, SegInfo si, int hi_bits, int lo_bits,
    int vbits, const float *__restrict__ bary, int d1, int32_t E, int32_t M, int32_t *__restrict__ csr_start,
    int32_t *__restrict__ csr_pix, float *__restrict__ csr_w) {
    extern __shared__ int cursor[];  // [warps][1 << (lo_bits - vbits)]: counts, then running output positions
    __shared__ int warp_sums[kBucketMaxThreads / 32];
    constexpr unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nthreads = blockDim.x, nwarps = nthreads >> 5;
    if (blockIdx.x == 0 && tid == 0) csr_start[M] = E;
    const int radix = 1 << hi_bits;
    // Large images (2^lo_bits vertices per bucket would not fit the cursors): 2^vbits CTAs share a bucket,
    // each owns a contiguous 2^(lo_bits - vbits) sub-range of its vertices, walks ALL entries of the bucket
    // (L2 hits) and places the ones of its sub-range.
    const int sub = blockIdx.x & ((1 << vbits) - 1);
    const int bucket = blockIdx.x >> vbits;
    const int seg = bucket >> hi_bits, digit = bucket & (radix - 1);
    const int kb = si.key_base[seg];
    const int Mb = si.key_base[seg + 1] - kb;
    const int sbits = lo_bits - vbits;
    const int v0 = (digit << lo_bits) + (sub << sbits);  // first local vertex of this CTA
    if (v0 >= Mb) return;
    const int nv = min(1 << sbits, Mb - v0);
    int start, end;
    if (hi_bits > 0) {
        const int tiles = si.tile_base[seg + 1] - si.tile_base[seg];
        const int32_t *hs = hist_scanned + (int64_t)radix * si.tile_base[seg];
        start = hs[(int64_t)digit * tiles];
        end = digit + 1 < radix ? hs[(int64_t)(digit + 1) * tiles] : si.seg_start[seg + 1];
    } else {
        start = si.seg_start[seg];
        end = si.seg_start[seg + 1];
    }
    const uint32_t vbase = (uint32_t)(kb + v0);  // keys of this CTA: [vbase, vbase + nv); below: earlier sub-ranges
    const int stride = 1 << sbits;
    for (int j = tid; j < nwarps * stride; j += nthreads) cursor[j] = 0;
    __syncthreads();
    // slice of this warp: whole 32-entry rounds, the last warp takes the remainder
    const int n = end - start;
    const int rounds_per_warp = (n / 32 + nwarps - 1) / nwarps;
    const int s0 = min(n, warp * rounds_per_warp * 32);
    const int s1 = (warp == nwarps - 1) ? n : min(n, s0 + rounds_per_warp * 32);
    int *mine = cursor + warp * stride;
    // counts of the slice (4 loads in flight per lane); entries of earlier sub-ranges only move my rows up
    int below = 0;
    {
        int i = start + s0 + lane;
        const int iend = start + s1;
        const uint32_t unv = (uint32_t)nv;
        for (; i + 96 < iend; i += 128) {
            const uint32_t k0 = pairs[i].x, k1 = pairs[i + 32].x, k2 = pairs[i + 64].x, k3 = pairs[i + 96].x;
            if (k0 - vbase < unv) atomicAdd(&mine[k0 - vbase], 1);
            if (k1 - vbase < unv) atomicAdd(&mine[k1 - vbase], 1);
            if (k2 - vbase < unv) atomicAdd(&mine[k2 - vbase], 1);
            if (k3 - vbase < unv) atomicAdd(&mine[k3 - vbase], 1);
            below += (k0 < vbase) + (k1 < vbase) + (k2 < vbase) + (k3 < vbase);
        }
        for (; i < iend; i += 32) {
            const uint32_t k = pairs[i].x;
            if (k - vbase < unv) atomicAdd(&mine[k - vbase], 1);
            below += k < vbase;
        }
    }
    if (vbits > 0) {  // block total of `below`
        __shared__ int below_total;
        if (tid == 0) below_total = 0;
        __syncthreads();
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) below += __shfl_xor_sync(FULL, below, o);
        if (lane == 0 && below) atomicAdd(&below_total, below);
        __syncthreads();
        below = below_total;
    }
    __syncthreads();
    // row starts: scan over the vertices of (sum over the warps); cursors: (vertex, warp) in that order
    const int per = (nv + nthreads - 1) / nthreads;
    const int j0 = min(tid * per, nv), j1 = min(j0 + per, nv);
    int sum = 0;
    for (int j = j0; j < j1; j++)
        for (int w = 0; w < nwarps; w++) sum += cursor[w * stride + j];
    const int inc = warp_inclusive_scan(sum);
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const int ws = lane < nwarps ? warp_sums[lane] : 0;
        const int winc = warp_inclusive_scan(ws);
        if (lane < nwarps) warp_sums[lane] = winc - ws;
    }
    __syncthreads();
    int run = start + below + warp_sums[warp] + inc - sum;
    for (int j = j0; j < j1; j++) {
        csr_start[vbase + j] = run;
        for (int w = 0; w < nwarps; w++) {
            const int c = cursor[w * stride + j];
            cursor[w * stride + j] = run;
            run += c;
        }
    }
    __syncthreads();
    // stable placement of the slice in 32-entry rounds, R rounds per trip.  The placement itself is a short
    // chain through shared memory; what has to be hidden is the memory latency of the pair loads and of
    // the dependent weight gathers: pairs are requested two trips ahead, weights one trip ahead.
    constexpr int R = 4;
    const int ibeg = start + s0, iend = start + s1;
    uint2 eA[R], eB[R], eC[R];
    float wA[R], wB[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int ia = ibeg + r * 32 + lane, ib = ia + R * 32;
        eA[r] = ia < iend ? pairs[ia] : make_uint2(0u, 0u);
        eB[r] = ib < iend ? pairs[ib] : make_uint2(0u, 0u);
    }
#pragma unroll
    for (int r = 0; r < R; r++) wA[r] = (ibeg + r * 32 + lane < iend) ? bary[eA[r].y] : 0.f;
    for (int base = ibeg; base < iend; base += R * 32) {  // warp-uniform loop
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int ib = base + (R + r) * 32 + lane, ic = ib + R * 32;
            wB[r] = ib < iend ? bary[eB[r].y] : 0.f;
            eC[r] = ic < iend ? pairs[ic] : make_uint2(0u, 0u);
        }
#pragma unroll
        for (int r = 0; r < R; r++) {
            const bool valid = base + r * 32 + lane < iend && eA[r].x - vbase < (uint32_t)nv;
            if (vbits > 0 && !__any_sync(FULL, valid)) continue;  // a round without entries of my sub-range
            const int key = valid ? (int)(eA[r].x - vbase) : -1 - lane;
            const unsigned peers = __match_any_sync(FULL, key);
            const int rank = __popc(peers & ((1u << lane) - 1));
            int pos = 0;
            if (valid) pos = mine[key] + rank;
            __syncwarp();
            if (valid && rank == 0) mine[key] += __popc(peers);
            __syncwarp();
            if (valid) {
                csr_pix[pos] = (int32_t)(eA[r].y / (uint32_t)d1);
                csr_w[pos] = wA[r];
            }
        }
#pragma unroll
        for (int r = 0; r < R; r++) {
            eA[r] = eB[r];
            wA[r] = wB[r];
            eB[r] = eC[r];
        }
    }
}

}  // namespace

bool bucket_sort_to_csr(uint2 *pairs_a, uint2 *pairs_b, const std::vector<int64_t> &seg_start,
                        const int32_t *d_key_base, int local_bits, const float *bary, int d1, int64_t E, int64_t M,
                        int32_t *csr_start, int32_t *csr_pix, float *csr_w, int prof_tag, cudaStream_t s,
                        std::vector<int32_t> &h_seg, std::vector<int32_t> &h_tile) {
    const int S = (int)seg_start.size() - 1;
    if (S <= 0 || E <= 0) return false;
    // DCRF_SORT_LSD=1: take the multi-pass LSD path that otherwise only images with more than 2^23 vertices
    // reach (kept under test by tests/test_gpu_parity.py::test_lsd_sort_path_gives_the_same_rows)
    static const bool force_lsd = [] { const char *e = getenv("DCRF_SORT_LSD"); return e && atoi(e) != 0; }();
    if (force_lsd) return false;
    static const int env_lo = [] { const char *e = getenv("DCRF_BUCKET_LO"); return e ? atoi(e) : 0; }();
    const int want_lo = env_lo > 0 ? env_lo : 8;
    int hi_bits = std::min(kSegMaxDigitBits, std::max(0, local_bits - want_lo));
    const int lo_bits = std::max(1, local_bits - hi_bits);
    if (lo_bits > kBucketMaxLoBits || ((int64_t)S << (hi_bits + 3)) > (int64_t)2000000000) return false;  // LSD passes instead
    h_seg.assign(S + 1, 0);  // staging of asynchronous uploads: owned by the caller, outlives them
    h_tile.assign(S + 1, 0);
    for (int i = 0; i <= S; i++) h_seg[i] = (int32_t)seg_start[i];
    for (int i = 0; i < S; i++) h_tile[i + 1] = h_tile[i] + ceil_div(seg_start[i + 1] - seg_start[i], kSortTile);
    const int total_tiles = h_tile[S];
    DevBuf<int32_t> d_seg, d_tile, hist;
    d_seg.alloc(S + 1, s);
    d_tile.alloc(S + 1, s);
    DCRF_CUDA(copy_h2d(d_seg.p, h_seg.data(), sizeof(int32_t) * (S + 1), s));
    DCRF_CUDA(copy_h2d(d_tile.p, h_tile.data(), sizeof(int32_t) * (S + 1), s));
    SegInfo si{d_seg.p, d_key_base, d_tile.p, S};
    const uint2 *grouped = pairs_a;
    {
        ProfScope prof(DCRF_K_BUILD_SORT, prof_tag, s);
        if (hi_bits > 0) {
            const int radix = 1 << hi_bits;
            hist.alloc((size_t)radix * total_tiles + 1, s);
            seg_radix_hist_kernel<<<total_tiles, kSortThreads, 0, s>>>(pairs_a, hist.p, si, lo_bits, hi_bits);
            DCRF_LAUNCHED();
            scan_rec(hist.p, hist.p, (int64_t)radix * total_tiles, false, s);
            seg_radix_scatter_kernel<<<total_tiles, kSortThreads, 0, s>>>(pairs_a, pairs_b, hist.p, si, lo_bits, hi_bits);
            DCRF_LAUNCHED();
            grouped = pairs_b;
        }
    }
    ProfScope prof(DCRF_K_BUILD_CSR, prof_tag, s);
    static const int env_warps = [] { const char *e = getenv("DCRF_BUCKET_WARPS"); return e ? atoi(e) : 0; }();
    const int vbits = std::max(0, lo_bits - kBucketCtaBits);
    const int warps = env_warps > 0 ? std::min(env_warps, kBucketMaxThreads / 32) : kBucketMaxThreads / 32;
    const size_t smem = sizeof(int) * ((size_t)warps << (lo_bits - vbits));
    bucket_csr_kernel<<<(unsigned)((int64_t)S << (hi_bits + vbits)), warps * 32, smem, s>>>(
        grouped, hist.p, si, hi_bits, lo_bits, vbits, bary, d1, (int32_t)E, (int32_t)M, csr_start, csr_pix, csr_w);
    DCRF_LAUNCHED();
    return true;
}

void exclusive_scan_i32(const int32_t *in, int32_t *out, int64_t n, cudaStream_t s) {
    scan_rec(in, out, n, true, s);
}

}  // namespace dcrf
