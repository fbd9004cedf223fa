// lattice_build.cu -- permutohedral lattice construction on the GPU (sm_100a).
//
// Replaces the set-up half of pydensecrf's addPairwiseGaussian / addPairwiseBilateral /
// addPairwiseEnergy (call sites: /root/reference/03c_hsn/utilities.py:435,439-440), i.e. the
// sequential `Permutohedral::init` [EXT] specified in SURVEY.md Appendix A.3:
//   K1  point kernel : features -> elevate -> nearest remainder-0 point -> rank -> barycentric
//   K2  hash insert  : open-addressing table of lattice keys, slot value = min entry index
//   K3  numbering    : per-pixel first-occurrence masks + exclusive scan of their counts
//                      => the reference's vertex ids
//   K4  assign       : offsets, vertex keys; compact (L2-resident) table of vertex ids
//   (K2-K4 skip "duplicate" pixels -- same simplex as their left neighbour -- see run_leader)
//   K5  neighbours   : +-1 neighbours along each of the d+1 axes by table lookup
//   K6  CSR          : stable radix sort of entries by vertex id => rows of the transposed
//                      incidence matrix in ascending entry order (deterministic splat)
//
// Bit-exactness: the coordinate math uses explicit round-to-nearest intrinsics (__fmul_rn, ...)
// so nvcc can never contract it into FMAs, true IEEE division for the features, and the same
// float/double/int mixing as the specification.  The vertex numbering is made independent of the
// (racy) insertion order: every table slot ends up holding the SMALLEST entry index e = p*(d+1)+r
// that carries its key (atomicMin), and ids are the scan of "I am the first occurrence" flags --
// exactly "id = rank of the key's first occurrence in pixel-major, remainder-minor scan order".
#include <math.h>

#include <algorithm>
#include <memory>

#include "common.cuh"

namespace dcrf {

namespace {

constexpr int kThreads = 256;

struct GeomDev {
    const int *w, *h, *pix_start;  // [B], [B], [B+1]
    int B;
};

__device__ __forceinline__ int find_image(const int *__restrict__ start, int B, int64_t x) {
    // largest b with start[b] <= x
    int lo = 0, hi = B - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if ((int64_t)start[mid] <= x) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}

template <int D>
struct Scale {
    float s[D];
};

// canonical simplex entry (A.3 step 7): canon[r][k] = r for k <= D - r, else r - (D+1)
template <int D>
__device__ __forceinline__ int canon(int r, int k) {
    return (k <= D - r) ? r : r - (D + 1);
}

// key of entry (pixel record, remainder r): key[i] = rem0[i] + canon[r][rank[i]], i < D
template <int D>
__device__ __forceinline__ void entry_key(const int4 rem, uint32_t rankpack, int r, short key[8]) {
    const short *rm = reinterpret_cast<const short *>(&rem);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        if (i < D) {
            int rk = (rankpack >> (4 * i)) & 15;
            key[i] = (short)(rm[i] + canon<D>(r, rk));
        } else {
            key[i] = 0;
        }
    }
}

__device__ __forceinline__ int4 pack_key(const short key[8]) {
    int4 v;
    v.x = (int)(((uint32_t)(uint16_t)key[0]) | ((uint32_t)(uint16_t)key[1] << 16));
    v.y = (int)(((uint32_t)(uint16_t)key[2]) | ((uint32_t)(uint16_t)key[3] << 16));
    v.z = (int)(((uint32_t)(uint16_t)key[4]) | ((uint32_t)(uint16_t)key[5] << 16));
    v.w = (int)(((uint32_t)(uint16_t)key[6]) | ((uint32_t)(uint16_t)key[7] << 16));
    return v;
}

__device__ __forceinline__ bool key_eq(const int4 a, const int4 b) {
    return a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w;
}

__device__ __forceinline__ uint32_t key_hash(const int4 k) {
    uint32_t h = 0x811C9DC5u;
    h = (h ^ (uint32_t)k.x) * 0x01000193u;
    h ^= h >> 15;
    h = (h ^ (uint32_t)k.y) * 0x85EBCA6Bu;
    h ^= h >> 13;
    h = (h ^ (uint32_t)k.z) * 0xC2B2AE35u;
    h ^= h >> 16;
    h = (h ^ (uint32_t)k.w) * 0x27D4EB2Fu;
    h ^= h >> 15;
    return h;
}

// Pixel record = first D coordinates of the remainder-0 point (int16 each) + the first D ranks (a nibble
// each).  For D <= 6 both fit ONE int4 (coordinates in x, y, z, ranks in w): every look-up of the owner of
// an occupied hash slot is a single 16-byte sector instead of a 16-byte and a 4-byte gather.  D = 7
// keeps the ranks in a second array.
struct PixRec {
    int4 rem;       // w = 0 for D <= 6 (entry_key ignores coordinates >= D)
    uint32_t rank;
};
template <int D>
__device__ __forceinline__ PixRec load_rec(const int4 *__restrict__ rec_rem, const uint32_t *__restrict__ rec_rank,
                                           int64_t p) {
    PixRec r;
    r.rem = rec_rem[p];
    if (D <= 6) {
        r.rank = (uint32_t)r.rem.w;
        r.rem.w = 0;
    } else {
        r.rank = rec_rank[p];
    }
    return r;
}

// ---------------------------------------------------------------------------------------------
// K1: one thread per pixel (Appendix A.2 + A.3 steps 2-6)
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(kThreads) lattice_point_kernel(
    GeomDev g, int64_t Ntot, int mode, float sx, float sy, float sr, float sg, float sb,
    const uint8_t *__restrict__ rgb, const float *__restrict__ features, Scale<D> scale,
    int4 *__restrict__ rec_rem, uint32_t *__restrict__ rec_rank, float *__restrict__ bary_out) {
    const int64_t gp = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (gp >= Ntot) return;
    float f[D];
    if (mode == 2) {
#pragma unroll
        for (int j = 0; j < D; j++) f[j] = features[(int64_t)j * Ntot + gp];
    } else {
        const int b = find_image(g.pix_start, g.B, gp);
        const int lp = (int)(gp - g.pix_start[b]);
        const int W = g.w[b];
        const int y = lp / W, x = lp - y * W;
        // A.2: float32 true division of an integer by a float32 parameter
        if (D >= 1) f[0] = __fdiv_rn((float)x, sx);
        if (D >= 2) f[1] = __fdiv_rn((float)y, sy);
        if (mode == 1 && D >= 5) {
            f[2] = __fdiv_rn((float)rgb[gp * 3 + 0], sr);
            f[3] = __fdiv_rn((float)rgb[gp * 3 + 1], sg);
            f[4] = __fdiv_rn((float)rgb[gp * 3 + 2], sb);
        }
    }
    // step 2: elevate (no FMA)
    float e[D + 1];
    float sm = 0.f;
#pragma unroll
    for (int j = D; j > 0; j--) {
        float cf = __fmul_rn(f[j - 1], scale.s[j - 1]);
        e[j] = __fsub_rn(sm, __fmul_rn((float)j, cf));
        sm = __fadd_rn(sm, cf);
    }
    e[0] = sm;
    // step 3: nearest remainder-0 point; sum is an int accumulator of float terms
    const float down_factor = 1.0f / (float)(D + 1);
    const float up_factor = (float)(D + 1);
    float rem0[D + 1];
    int sum = 0;
#pragma unroll
    for (int i = 0; i <= D; i++) {
        float v = __fmul_rn(down_factor, e[i]);
        float up = __fmul_rn(ceilf(v), up_factor);
        float down = __fmul_rn(floorf(v), up_factor);
        int rd2;
        if (__fsub_rn(up, e[i]) < __fsub_rn(e[i], down)) rd2 = (short)(int)up;
        else rd2 = (short)(int)down;
        rem0[i] = (float)rd2;
        sum = (int)__fadd_rn((float)sum, __fmul_rn((float)rd2, down_factor));
    }
    // step 4: rank
    int rank[D + 1];
#pragma unroll
    for (int i = 0; i <= D; i++) rank[i] = 0;
#pragma unroll
    for (int i = 0; i < D; i++) {
        float di = __fsub_rn(e[i], rem0[i]);
#pragma unroll
        for (int j = i + 1; j <= D; j++) {
            if (di < __fsub_rn(e[j], rem0[j])) rank[i]++;
            else rank[j]++;
        }
    }
    // step 5: re-project
#pragma unroll
    for (int i = 0; i <= D; i++) {
        rank[i] += sum;
        if (rank[i] < 0) {
            rank[i] += D + 1;
            rem0[i] = __fadd_rn(rem0[i], (float)(D + 1));
        } else if (rank[i] > D) {
            rank[i] -= D + 1;
            rem0[i] = __fsub_rn(rem0[i], (float)(D + 1));
        }
    }
    // step 6: barycentric (updates applied in i order, like the sequential code)
    float bc[D + 2];
#pragma unroll
    for (int i = 0; i <= D + 1; i++) bc[i] = 0.f;
#pragma unroll
    for (int i = 0; i <= D; i++) {
        float v = __fmul_rn(__fsub_rn(e[i], rem0[i]), down_factor);
        int idx = D - rank[i];
        // (selects, not branches: an `if (k == idx)` form is turned into a dynamically indexed local array)
#pragma unroll
        for (int k = 0; k <= D + 1; k++) {
            const float up = __fadd_rn(bc[k], v), dn = __fsub_rn(bc[k], v);
            bc[k] = (k == idx) ? up : ((k == idx + 1) ? dn : bc[k]);
        }
    }
    bc[0] = (float)((double)bc[0] + (1.0 + (double)bc[D + 1]));
    if ((D + 1) % 2 == 0) {  // 8-byte aligned run of d+1 floats: half the store instructions
        float2 *bo = reinterpret_cast<float2 *>(bary_out + gp * (D + 1));
#pragma unroll
        for (int r = 0; r + 1 <= D; r += 2) bo[r / 2] = make_float2(bc[r], bc[r + 1]);
    } else {
#pragma unroll
        for (int r = 0; r <= D; r++) bary_out[gp * (D + 1) + r] = bc[r];
    }
    // pixel record: first D coordinates of rem0 as int16, first D ranks as nibbles
    short rm[8];
    uint32_t rp = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        if (i < D) {
            rm[i] = (short)(int)rem0[i];
            rp |= (uint32_t)(rank[i] & 15) << (4 * i);
        } else {
            rm[i] = 0;
        }
    }
    int4 rec = pack_key(rm);
    if (D <= 6) rec.w = (int)rp;  // one-sector record (see load_rec)
    else rec_rank[gp] = rp;
    rec_rem[gp] = rec;
}

// The d+1 per-entry words of a pixel are contiguous (entry e = p * (d+1) + r): move them with the widest
// aligned access instead of d+1 strided 4-byte ones (ncu: K3 / K4 sat at 80-86 % of the L1TEX pipe).
template <int N>
__device__ __forceinline__ void load_run(const int32_t *__restrict__ base, int32_t (&v)[N]) {
    if (N % 4 == 0) {
#pragma unroll
        for (int i = 0; i < N / 4; i++) {
            const int4 t = reinterpret_cast<const int4 *>(base)[i];
            v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
        }
    } else if (N % 2 == 0) {
#pragma unroll
        for (int i = 0; i < N / 2; i++) {
            const int2 t = reinterpret_cast<const int2 *>(base)[i];
            v[2 * i] = t.x; v[2 * i + 1] = t.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < N; i++) v[i] = base[i];
    }
}
template <int N>
__device__ __forceinline__ void store_run(int32_t *__restrict__ base, const int32_t (&v)[N]) {
    if (N % 4 == 0) {
#pragma unroll
        for (int i = 0; i < N / 4; i++)
            reinterpret_cast<int4 *>(base)[i] = make_int4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else if (N % 2 == 0) {
#pragma unroll
        for (int i = 0; i < N / 2; i++) reinterpret_cast<int2 *>(base)[i] = make_int2(v[2 * i], v[2 * i + 1]);
    } else {
#pragma unroll
        for (int i = 0; i < N; i++) base[i] = v[i];
    }
}

// ---------------------------------------------------------------------------------------------
// Run leaders.  Neighbouring pixels of a natural image very often fall into the SAME simplex with the
// same remainder-0 point: their pixel records (rem0, rank) are equal, hence all d+1 lattice keys are.
// Such a pixel ("duplicate") can never hold the first occurrence of a key -- its left neighbour has the
// same keys at smaller entry indices -- so K2-K4 let only the first pixel of a run inside a warp (the
// "leader") touch the hash table and the numbering arrays; the duplicates copy the leader's vertex ids
// with a shuffle.  K2, K3 and K4 use the same pixel <-> lane mapping, so each recomputes the same flags.
// ---------------------------------------------------------------------------------------------
struct RunLeader {
    bool leader;   // this lane does the look-ups itself
    int src;       // lane to copy from (== own lane for a leader)
};
__device__ __forceinline__ RunLeader run_leader(bool valid, bool first_of_image, const int4 rem,
                                                const uint32_t rp) {
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int px = __shfl_up_sync(FULL, rem.x, 1), py = __shfl_up_sync(FULL, rem.y, 1);
    const int pz = __shfl_up_sync(FULL, rem.z, 1), pw = __shfl_up_sync(FULL, rem.w, 1);
    const uint32_t pr = __shfl_up_sync(FULL, rp, 1);
    const bool dup = valid && lane > 0 && !first_of_image && px == rem.x && py == rem.y && pz == rem.z &&
                     pw == rem.w && pr == rp;
    const unsigned leaders = __ballot_sync(FULL, !dup);
    RunLeader r;
    r.leader = valid && !dup;
    r.src = 31 - __clz(leaders & (FULL >> (31 - lane)));  // lane 0 is always a leader
    return r;
}

// ---------------------------------------------------------------------------------------------
// K2: hash insert.  One thread per pixel, d+1 inserts by the run leaders.  table[] holds entry indices
// (-1 = empty); a slot's key never changes once claimed, only its representative shrinks (atomicMin).
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(kThreads) hash_insert_kernel(
    GeomDev g, int64_t Ntot, const int64_t *__restrict__ tab_start, const int *__restrict__ tab_mask,
    const int4 *__restrict__ rec_rem, const uint32_t *__restrict__ rec_rank, int32_t *table,
    int32_t *__restrict__ slot_of) {
    const int64_t gp0 = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const bool valid = gp0 < Ntot;
    const int64_t gp = valid ? gp0 : Ntot - 1;
    const int b = find_image(g.pix_start, g.B, gp);
    int32_t *tab = table + tab_start[b];
    const uint32_t mask = (uint32_t)tab_mask[b];
    const PixRec me = load_rec<D>(rec_rem, rec_rank, gp);
    const int4 rem = me.rem;
    const uint32_t rp = me.rank;
    const RunLeader rl = run_leader(valid, gp == (int64_t)g.pix_start[b], rem, rp);
    if (!rl.leader) return;  // no warp-level operation below
    int32_t slots[D + 1];
#pragma unroll
    for (int r = 0; r <= D; r++) {
        short key[8];
        entry_key<D>(rem, rp, r, key);
        const int4 pk = pack_key(key);
        const int32_t e = (int32_t)(gp * (D + 1) + r);
        uint32_t h = key_hash(pk) & mask;
        for (;;) {
            int32_t cur = tab[h];
            if (cur < 0) {
                cur = atomicCAS(&tab[h], -1, e);
                if (cur < 0) break;  // claimed
            }
            // occupied by entry `cur`: same key?
            const int64_t cp = cur / (D + 1);
            const int cr = cur - (int32_t)cp * (D + 1);
            short ck[8];
            const PixRec owner = load_rec<D>(rec_rem, rec_rank, cp);
            entry_key<D>(owner.rem, owner.rank, cr, ck);
            if (key_eq(pack_key(ck), pk)) {
                if (e < cur) atomicMin(&tab[h], e);
                break;
            }
            h = (h + 1) & mask;
        }
        slots[r] = (int32_t)h;
    }
    store_run<D + 1>(slot_of + gp * (D + 1), slots);  // leaders' entries only; the duplicates' slots are never read
}

// K3: per pixel, which of its d+1 entries are the first occurrence of their key (bit r of mask8) and how
// many (cnt, scanned into the vertex ids); the slot of every leader entry is replaced by the table's
// representative entry, so that K4 does not probe the (DRAM-sized) table again.
template <int D>
__global__ void __launch_bounds__(kThreads) first_mask_kernel(
    GeomDev g, int64_t Ntot, const int64_t *__restrict__ tab_start, const int32_t *__restrict__ table,
    const int4 *__restrict__ rec_rem, const uint32_t *__restrict__ rec_rank, int32_t *__restrict__ slot_rep,
    int32_t *__restrict__ cnt, uint8_t *__restrict__ mask8) {
    const int64_t gp0 = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const bool valid = gp0 < Ntot;
    const int64_t gp = valid ? gp0 : Ntot - 1;
    const int b = find_image(g.pix_start, g.B, gp);
    const PixRec me = load_rec<D>(rec_rem, rec_rank, gp);
    const RunLeader rl = run_leader(valid, gp == (int64_t)g.pix_start[b], me.rem, me.rank);
    if (!valid) return;
    unsigned m = 0;
    if (rl.leader) {
        const int32_t *tab = table + tab_start[b];
        int32_t slot[D + 1], rep[D + 1];
        load_run<D + 1>(slot_rep + gp * (D + 1), slot);
#pragma unroll
        for (int r = 0; r <= D; r++) rep[r] = tab[slot[r]];
#pragma unroll
        for (int r = 0; r <= D; r++)
            if (rep[r] == (int32_t)(gp * (D + 1) + r)) m |= 1u << r;
        store_run<D + 1>(slot_rep + gp * (D + 1), rep);
    }
    cnt[gp] = __popc(m);
    mask8[gp] = (uint8_t)m;
}

// K4: vertex id of every entry = (first occurrences in the pixels before the representative's pixel)
// + (first occurrences among the lower remainders of that pixel); offsets, (vertex, entry) sort pairs,
// and the keys of the vertices (published by their first occurrences).
template <int D>
__global__ void __launch_bounds__(kThreads) assign_kernel(
    GeomDev g, int64_t Ntot, const int32_t *__restrict__ slot_rep, const int32_t *__restrict__ pscan,
    const uint8_t *__restrict__ mask8, const int4 *__restrict__ rec_rem, const uint32_t *__restrict__ rec_rank,
    int32_t *__restrict__ offset, int4 *__restrict__ vkeys, uint2 *__restrict__ sort_pairs) {
    constexpr unsigned FULL = 0xffffffffu;
    const int64_t gp0 = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const bool valid = gp0 < Ntot;
    const int64_t gp = valid ? gp0 : Ntot - 1;
    const int b = find_image(g.pix_start, g.B, gp);
    const PixRec me = load_rec<D>(rec_rem, rec_rank, gp);
    const int4 rem = me.rem;
    const uint32_t rp = me.rank;
    const RunLeader rl = run_leader(valid, gp == (int64_t)g.pix_start[b], rem, rp);
    const unsigned mine = rl.leader ? mask8[gp] : 0u;
    int32_t rep[D + 1];
#pragma unroll
    for (int r = 0; r <= D; r++) rep[r] = 0;
    if (rl.leader) load_run<D + 1>(slot_rep + gp * (D + 1), rep);
    int32_t id[D + 1];
#pragma unroll
    for (int r = 0; r <= D; r++) {
        const int32_t pp = rep[r] / (D + 1);
        const int rr = rep[r] - pp * (D + 1);
        id[r] = rl.leader ? pscan[pp] + __popc((unsigned)mask8[pp] & ((1u << rr) - 1u)) : 0;
    }
#pragma unroll
    for (int r = 0; r <= D; r++) id[r] = __shfl_sync(FULL, id[r], rl.src);
    if (!valid) return;
    const int64_t e0 = gp * (D + 1);
    store_run<D + 1>(offset + e0, id);
    // (vertex, entry) pairs for the CSR sort (K6): 16-byte stores of two pairs where d+1 is even
    if ((D + 1) % 2 == 0) {
        uint4 *sp = reinterpret_cast<uint4 *>(sort_pairs + e0);
#pragma unroll
        for (int r = 0; r + 1 <= D; r += 2)
            sp[r / 2] = make_uint4((uint32_t)id[r], (uint32_t)(e0 + r), (uint32_t)id[r + 1], (uint32_t)(e0 + r + 1));
    } else {
#pragma unroll
        for (int r = 0; r <= D; r++) sort_pairs[e0 + r] = make_uint2((uint32_t)id[r], (uint32_t)(e0 + r));
    }
#pragma unroll
    for (int r = 0; r <= D; r++) {
        if ((mine >> r) & 1u) {
            short key[8];
            entry_key<D>(rem, rp, r, key);
            vkeys[id[r]] = pack_key(key);
        }
    }
}

// K4b: compact table of vertex ids, one insert per vertex (keys are unique: plain CAS claim)
__global__ void __launch_bounds__(kThreads) compact_insert_kernel(
    int64_t M, int B, const int32_t *__restrict__ vert_start, const int64_t *__restrict__ tab_start,
    const int *__restrict__ tab_mask, const int4 *__restrict__ vkeys, int32_t *table) {
    const int64_t v = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (v >= M) return;
    const int b = find_image(vert_start, B, v);
    int32_t *tab = table + tab_start[b];
    const uint32_t mask = (uint32_t)tab_mask[b];
    uint32_t h = key_hash(vkeys[v]) & mask;
    while (atomicCAS(&tab[h], -1, (int32_t)v) != -1) h = (h + 1) & mask;
}

// Wide-slot form of the compact table (d <= 6: the key fits three ints): a slot is (key.x, key.y,
// key.z, vertex id), so a probe is ONE 16-byte read and never touches vkeys[] -- the narrow table
// costs a 4-byte slot read plus a dependent 16-byte vkeys gather per probed slot.  Keys are unique, so
// an insert claims the id word with a CAS and then stores its key words; the look-ups run in a later
// kernel.  (neighbour_kernel<5>, VOC batch of 32: 485 us with narrow slots.)
__global__ void __launch_bounds__(kThreads) compact_insert_wide_kernel(
    int64_t M, int B, const int32_t *__restrict__ vert_start, const int64_t *__restrict__ tab_start,
    const int *__restrict__ tab_mask, const int4 *__restrict__ vkeys, int4 *table) {
    const int64_t v = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (v >= M) return;
    const int b = find_image(vert_start, B, v);
    int4 *tab = table + tab_start[b];
    const uint32_t mask = (uint32_t)tab_mask[b];
    const int4 k = vkeys[v];
    uint32_t h = key_hash(k) & mask;
    while (atomicCAS(&tab[h].w, -1, (int32_t)v) != -1) h = (h + 1) & mask;
    int *slot = reinterpret_cast<int *>(&tab[h]);
    slot[0] = k.x;
    slot[1] = k.y;
    slot[2] = k.z;
}

template <int D>
__global__ void __launch_bounds__(kThreads) neighbour_wide_kernel(
    int64_t M, int B, const int32_t *__restrict__ vert_start, const int64_t *__restrict__ tab_start,
    const int *__restrict__ tab_mask, const int4 *__restrict__ table, const int4 *__restrict__ vkeys,
    int2 *__restrict__ neigh) {
    // D+1 consecutive CTAs look up the D+1 axes of the same 256 vertices: an image's (few MB) table is
    // probed for all axes while it is L2 resident (an axis-major grid streams every table D+1 times)
    const int j = (int)(blockIdx.x % (D + 1));
    const int64_t v = (int64_t)(blockIdx.x / (D + 1)) * kThreads + threadIdx.x;
    if (v >= M) return;
    const int b = find_image(vert_start, B, v);
    const int4 *tab = table + tab_start[b];
    const uint32_t mask = (uint32_t)tab_mask[b];
    const int4 kv = vkeys[v];
    const short *key = reinterpret_cast<const short *>(&kv);
    short nk[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        if (k < D) nk[k] = (short)(key[k] + ((k == j) ? D : -1));
        else nk[k] = 0;
    }
    const int4 pk = pack_key(nk);
    uint32_t h = key_hash(pk) & mask;
    int found = -1;
    for (;;) {
        const int4 cur = __ldg(tab + h);
        if (cur.w < 0) break;
        if (cur.x == pk.x && cur.y == pk.y && cur.z == pk.z) { found = cur.w; break; }
        h = (h + 1) & mask;
    }
    int *nflat = reinterpret_cast<int *>(neigh + (int64_t)j * M);
    nflat[2 * v] = found;
    if (found >= 0) nflat[2 * (int64_t)found + 1] = (int)v;
}

// vert_start[b] = number of first occurrences before image b's first pixel
__global__ void vert_start_kernel(const int *__restrict__ pix_start, int B,
                                  const int32_t *__restrict__ pscan, int32_t *__restrict__ vert_start) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b <= B) vert_start[b] = pscan[pix_start[b]];
}

// K5: neighbours.  One thread per (axis j, vertex v); table now holds vertex ids.  Only the first
// neighbour n1 = key + (-1,..,+d at j,..,-1) is looked up: the relation is mutual (n2 of u is v iff
// n1 of v is u), so the thread also stores itself as the second neighbour of the vertex it found.
// neigh[] is pre-filled with -1; every int is written by at most one thread.
template <int D>
__global__ void __launch_bounds__(kThreads) neighbour_kernel(
    int64_t M, int B, const int32_t *__restrict__ vert_start, const int64_t *__restrict__ tab_start,
    const int *__restrict__ tab_mask, const int32_t *__restrict__ table,
    const int4 *__restrict__ vkeys, int2 *__restrict__ neigh) {
    const int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (t >= M * (D + 1)) return;
    const int j = (int)(t / M);
    const int64_t v = t - (int64_t)j * M;
    const int b = find_image(vert_start, B, v);
    const int32_t *tab = table + tab_start[b];
    const uint32_t mask = (uint32_t)tab_mask[b];
    const int4 kv = vkeys[v];
    const short *key = reinterpret_cast<const short *>(&kv);
    short nk[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        if (k < D) nk[k] = (short)(key[k] + ((k == j) ? D : -1));
        else nk[k] = 0;
    }
    const int4 pk = pack_key(nk);
    uint32_t h = key_hash(pk) & mask;
    int found = -1;
    for (;;) {
        const int32_t cur = tab[h];
        if (cur < 0) break;
        if (key_eq(vkeys[cur], pk)) { found = cur; break; }
        h = (h + 1) & mask;
    }
    int *nflat = reinterpret_cast<int *>(neigh + (int64_t)j * M);
    nflat[2 * v] = found;
    if (found >= 0) nflat[2 * (int64_t)found + 1] = (int)v;
}

// K6 epilogue: sorted (vertex, entry) pairs -> CSR rows
__global__ void __launch_bounds__(kThreads) csr_finalize_kernel(
    const uint2 *__restrict__ sorted, const float *__restrict__ bary, int d1, int64_t E, int64_t M,
    int32_t *__restrict__ csr_start, int32_t *__restrict__ csr_pix, float *__restrict__ csr_w) {
    const int64_t s = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (s >= E) return;
    const uint2 ve = sorted[s];
    const uint32_t v = ve.x, e = ve.y;
    csr_pix[s] = (int32_t)(e / (uint32_t)d1);
    csr_w[s] = bary[e];
    if (s == 0 || sorted[s - 1].x != v) csr_start[v] = (int32_t)s;
    if (s == E - 1) csr_start[M] = (int32_t)E;
}


template <int D>
void build_begin_impl(const BatchGeom &g, const FeatureSpec &f, Lattice &out, BuildState &st, int32_t *pinned_vs,
                      cudaStream_t s) {
    const int B = g.B;
    const int64_t Ntot = g.Ntot;
    const int d1 = D + 1;
    const int64_t E = Ntot * d1;
    DCRF_REQUIRE(E < (int64_t)2147483000, DCRF_EINVAL, "batch too large: N*(d+1) must stay below 2^31");
    out.d = D;
    out.E = E;
    GeomDev gd{g.d_w, g.d_h, g.d_pix_start, B};

    // A.3 step 1 on the host: double math stored as float (identical to the specification)
    Scale<D> scale;
    const float inv_std_dev = (float)(sqrt(2.0 / 3.0) * (double)d1);
    for (int i = 0; i < D; i++)
        scale.s[i] = (float)(1.0 / sqrt((double)((i + 2) * (i + 1))) * (double)inv_std_dev);

    DevBuf<int4> &rec_rem = st.rec_rem;
    DevBuf<uint32_t> &rec_rank = st.rec_rank;
    rec_rem.alloc(Ntot, s);
    if (D > 6) rec_rank.alloc(Ntot, s);
    out.bary.alloc(E, s);
    const int nbp = ceil_div(Ntot, kThreads);
    std::unique_ptr<ProfScope> prof(new ProfScope(DCRF_K_BUILD_POINT, D, s));
    lattice_point_kernel<D><<<nbp, kThreads, 0, s>>>(gd, Ntot, f.mode, f.s[0], f.s[1], f.s[2], f.s[3],
                                                    f.s[4], f.rgb, f.features, scale, rec_rem.p,
                                                    rec_rank.p, out.bary.p);
    DCRF_LAUNCHED();

    prof.reset();
    prof.reset(new ProfScope(DCRF_K_BUILD_HASH, D, s));
    // per-image table regions: capacity = pow2 >= 1.25 * N_b * (d+1) entries, i.e. a load factor of at
    // most 0.8 in the worst case of all-distinct keys (natural images: M ~ 0.1 E, load < 0.1)
    std::vector<int64_t> &tab_start = out.h_tab_start;
    std::vector<int> &tab_mask = out.h_tab_mask;
    tab_start.assign(B + 1, 0);
    tab_mask.assign(B, 0);
    for (int b = 0; b < B; b++) {
        int64_t need = (5 * (g.pix_start[b + 1] - g.pix_start[b]) * d1 + 3) / 4;
        int64_t cap = 64;
        while (cap < need) cap <<= 1;
        tab_mask[b] = (int)(cap - 1);
        tab_start[b + 1] = tab_start[b] + cap;
    }
    DevBuf<int64_t> d_tab_start;
    DevBuf<int> d_tab_mask;
    d_tab_start.alloc(B + 1, s);
    d_tab_mask.alloc(B, s);
    DCRF_CUDA(copy_h2d(d_tab_start.p, tab_start.data(), sizeof(int64_t) * (B + 1), s));
    DCRF_CUDA(copy_h2d(d_tab_mask.p, tab_mask.data(), sizeof(int) * B, s));
    DevBuf<int32_t> table;
    table.alloc(tab_start[B], s);
    DCRF_CUDA(cudaMemsetAsync(table.p, 0xFF, sizeof(int32_t) * tab_start[B], s));

    DevBuf<int32_t> &slot_of = st.slot_of;
    slot_of.alloc(E, s);
    hash_insert_kernel<D><<<nbp, kThreads, 0, s>>>(gd, Ntot, d_tab_start.p, d_tab_mask.p, rec_rem.p,
                                                  rec_rank.p, table.p, slot_of.p);
    DCRF_LAUNCHED();

    prof.reset();
    prof.reset(new ProfScope(DCRF_K_BUILD_NUMBER, D, s));
    DevBuf<int32_t> &pscan = st.pscan;   // per pixel: count of first occurrences, scanned in place
    DevBuf<uint8_t> &mask8 = st.mask8;   // per pixel: which remainders are first occurrences
    pscan.alloc(Ntot + 1, s);
    mask8.alloc(Ntot, s);
    first_mask_kernel<D><<<nbp, kThreads, 0, s>>>(gd, Ntot, d_tab_start.p, table.p, rec_rem.p, rec_rank.p,
                                                 slot_of.p, pscan.p, mask8.p);
    DCRF_LAUNCHED();
    exclusive_scan_i32(pscan.p, pscan.p, Ntot, s);

    // vertex counts: total + per image (host needs them to size everything else)
    DevBuf<int32_t> &d_vert_start = st.d_vert_start;
    d_vert_start.alloc(B + 1, s);
    vert_start_kernel<<<ceil_div(B + 1, 128), 128, 0, s>>>(g.d_pix_start, B, pscan.p, d_vert_start.p);
    DCRF_LAUNCHED();
    if (pinned_vs) {
        st.h_vs = pinned_vs;
    } else {
        st.h_vs_own.assign(B + 1, 0);
        st.h_vs = st.h_vs_own.data();
    }
    DCRF_CUDA(copy_d2h(st.h_vs, d_vert_start.p, sizeof(int32_t) * (B + 1), s));
}

template <int D>
void build_finish_impl(const BatchGeom &g, const FeatureSpec &f, Lattice &out, BuildState &st, cudaStream_t s) {
    (void)f;
    const int B = g.B;
    const int64_t Ntot = g.Ntot;
    const int d1 = D + 1;
    const int64_t E = Ntot * d1;
    GeomDev gd{g.d_w, g.d_h, g.d_pix_start, B};
    const int nbp = ceil_div(Ntot, kThreads);
    const int nbe = ceil_div(E, kThreads);
    DevBuf<int4> &rec_rem = st.rec_rem;
    DevBuf<uint32_t> &rec_rank = st.rec_rank;
    DevBuf<int32_t> &slot_of = st.slot_of, &pscan = st.pscan, &d_vert_start = st.d_vert_start;
    DevBuf<uint8_t> &mask8 = st.mask8;
    const int32_t *h_vs = st.h_vs;
    out.vert_start.assign(h_vs, h_vs + B + 1);
    const int64_t M = h_vs[B];
    out.M = M;
    std::unique_ptr<ProfScope> prof(new ProfScope(DCRF_K_BUILD_NUMBER, D, s));

    out.offset.alloc(E, s);
    out.vkeys.alloc((size_t)M * 8, s);
    int4 *vkeys4 = reinterpret_cast<int4 *>(out.vkeys.p);
    DevBuf<uint2> pa, pb;  // interleaved (vertex, entry) pairs: one 8-byte scatter per pair and pass
    pa.alloc(E, s);
    assign_kernel<D><<<nbp, kThreads, 0, s>>>(gd, Ntot, slot_of.p, pscan.p, mask8.p, rec_rem.p, rec_rank.p,
                                             out.offset.p, vkeys4, pa.p);
    DCRF_LAUNCHED();
    prof.reset();
    prof.reset(new ProfScope(DCRF_K_BUILD_NEIGH, D, s));
    // Neighbour look-ups go through a second, COMPACT table of vertex ids (capacity = pow2 >= 2 M_b per
    // image, ~1 MB per VOC image: the batch's tables stay L2 resident), instead of the insertion table
    // that is sized for the worst case M = E (16 MB per image, every probe a DRAM access).
    std::vector<int64_t> &tab2_start = out.h_tab2_start;
    std::vector<int> &tab2_mask = out.h_tab2_mask;
    tab2_start.assign(B + 1, 0);
    tab2_mask.assign(B, 0);
    for (int b = 0; b < B; b++) {
        const int64_t need = 2 * (int64_t)(h_vs[b + 1] - h_vs[b]);
        int64_t cap = 64;
        while (cap < need) cap <<= 1;
        tab2_mask[b] = (int)(cap - 1);
        tab2_start[b + 1] = tab2_start[b] + cap;
    }
    DevBuf<int64_t> d_tab2_start;
    DevBuf<int> d_tab2_mask;
    d_tab2_start.alloc(B + 1, s);
    d_tab2_mask.alloc(B, s);
    DCRF_CUDA(copy_h2d(d_tab2_start.p, tab2_start.data(), sizeof(int64_t) * (B + 1), s));
    DCRF_CUDA(copy_h2d(d_tab2_mask.p, tab2_mask.data(), sizeof(int) * B, s));
    out.neigh.alloc((size_t)M * d1, s);
    DCRF_CUDA(cudaMemsetAsync(out.neigh.p, 0xFF, sizeof(int2) * M * d1, s));
    if (D <= 6) {  // wide slots: (key.x, key.y, key.z, id); 0xFF fill = id -1 = empty
        DevBuf<int4> table2;
        table2.alloc(tab2_start[B], s);
        DCRF_CUDA(cudaMemsetAsync(table2.p, 0xFF, sizeof(int4) * tab2_start[B], s));
        compact_insert_wide_kernel<<<ceil_div(M, kThreads), kThreads, 0, s>>>(M, B, d_vert_start.p, d_tab2_start.p,
                                                                             d_tab2_mask.p, vkeys4, table2.p);
        DCRF_LAUNCHED();
        neighbour_wide_kernel<D><<<ceil_div(M, kThreads) * d1, kThreads, 0, s>>>(
            M, B, d_vert_start.p, d_tab2_start.p, d_tab2_mask.p, table2.p, vkeys4, out.neigh.p);
        DCRF_LAUNCHED();
    } else {
        DevBuf<int32_t> table2;
        table2.alloc(tab2_start[B], s);
        DCRF_CUDA(cudaMemsetAsync(table2.p, 0xFF, sizeof(int32_t) * tab2_start[B], s));
        compact_insert_kernel<<<ceil_div(M, kThreads), kThreads, 0, s>>>(M, B, d_vert_start.p, d_tab2_start.p,
                                                                        d_tab2_mask.p, vkeys4, table2.p);
        DCRF_LAUNCHED();
        neighbour_kernel<D><<<ceil_div(M * d1, kThreads), kThreads, 0, s>>>(
            M, B, d_vert_start.p, d_tab2_start.p, d_tab2_mask.p, table2.p, vkeys4, out.neigh.p);
        DCRF_LAUNCHED();
    }

    // transposed incidence rows: stable sort of entries by vertex id
    prof.reset();
    prof.reset(new ProfScope(DCRF_K_BUILD_SORT, D, s));
    pb.alloc(E, s);
    // per-image segments: keys local to an image need fewer radix passes than batch-global ids
    int64_t max_mb = 1;
    std::vector<int64_t> ent_start(B + 1);
    for (int b = 0; b <= B; b++) ent_start[b] = g.pix_start[b] * d1;
    for (int b = 0; b < B; b++) max_mb = std::max<int64_t>(max_mb, out.vert_start[b + 1] - out.vert_start[b]);
    int bits = 1;
    while (((int64_t)1 << bits) < max_mb) bits++;
    out.csr_start.alloc(M + 1, s);
    out.csr_pix.alloc(E, s);
    out.csr_w.alloc(E, s);
    prof.reset();
    if (bucket_sort_to_csr(pa.p, pb.p, ent_start, d_vert_start.p, bits, out.bary.p, d1, E, M, out.csr_start.p,
                           out.csr_pix.p, out.csr_w.p, D, s, out.h_seg, out.h_tile))
        return;
    prof.reset(new ProfScope(DCRF_K_BUILD_SORT, D, s));
    const int in_b = segmented_radix_sort_pairs(pa.p, pb.p, ent_start, d_vert_start.p, bits, s, out.h_seg, out.h_tile);
    const uint2 *sorted = in_b ? pb.p : pa.p;
    prof.reset();
    prof.reset(new ProfScope(DCRF_K_BUILD_CSR, D, s));
    csr_finalize_kernel<<<nbe, kThreads, 0, s>>>(sorted, out.bary.p, d1, E, M, out.csr_start.p, out.csr_pix.p,
                                                out.csr_w.p);
    DCRF_LAUNCHED();
}

}  // namespace

#define DCRF_DISPATCH_D(d, ...)                                                  \
    switch (d) {                                                                 \
        case 1: { constexpr int D = 1; __VA_ARGS__; } break;                     \
        case 2: { constexpr int D = 2; __VA_ARGS__; } break;                     \
        case 3: { constexpr int D = 3; __VA_ARGS__; } break;                     \
        case 4: { constexpr int D = 4; __VA_ARGS__; } break;                     \
        case 5: { constexpr int D = 5; __VA_ARGS__; } break;                     \
        case 6: { constexpr int D = 6; __VA_ARGS__; } break;                     \
        case 7: { constexpr int D = 7; __VA_ARGS__; } break;                     \
        default:                                                                 \
            throw Error{DCRF_EINVAL, "feature dimension d must be in [1, 7]"};   \
    }

void build_lattice_begin(const BatchGeom &g, const FeatureSpec &f, Lattice &out, BuildState &st, int32_t *pinned_vs,
                         cudaStream_t s) {
    DCRF_DISPATCH_D(f.d, build_begin_impl<D>(g, f, out, st, pinned_vs, s));
}

void build_lattice_finish(const BatchGeom &g, const FeatureSpec &f, Lattice &out, BuildState &st, cudaStream_t s) {
    DCRF_DISPATCH_D(f.d, build_finish_impl<D>(g, f, out, st, s));
}

void build_lattice(const BatchGeom &g, const FeatureSpec &f, Lattice &out, cudaStream_t s) {
    BuildState st;
    build_lattice_begin(g, f, out, st, nullptr, s);
    DCRF_CUDA(cudaStreamSynchronize(s));
    build_lattice_finish(g, f, out, st, s);
}


// ---------------------------------------------------------------------------------------------
// replication of a geometry-only lattice: `one` was built over the DISTINCT image sizes of a batch,
// image b of the batch is a copy of unique image src[b] with its pixel / vertex / entry ids shifted
// ---------------------------------------------------------------------------------------------
namespace {
// where image b's copy comes from and goes to (pixels, vertices; entries = pixels * (d+1))
struct RepImg {
    int src_pix, dst_pix, n_pix;
    int src_vert, dst_vert, n_vert;
};

// what to add to an id stored in the array: nothing, the vertex shift, the pixel shift or the entry shift
enum { kShiftNone = 0, kShiftVert = 1, kShiftPix = 2, kShiftEnt = 3 };
// which index space the array lives in
enum { kOverPix = 0, kOverVert = 1, kOverEnt = 2 };

__device__ __forceinline__ int rep_shift(const RepImg &r, int kind, int d1) {
    if (kind == kShiftVert) return r.dst_vert - r.src_vert;
    if (kind == kShiftPix) return r.dst_pix - r.src_pix;
    if (kind == kShiftEnt) return (r.dst_pix - r.src_pix) * d1;
    return 0;
}
__device__ __forceinline__ void rep_range(const RepImg &r, int over, int d1, int64_t &src, int64_t &dst, int64_t &n) {
    if (over == kOverVert) {
        src = r.src_vert; dst = r.dst_vert; n = r.n_vert;
    } else if (over == kOverEnt) {
        src = (int64_t)r.src_pix * d1; dst = (int64_t)r.dst_pix * d1; n = (int64_t)r.n_pix * d1;
    } else {
        src = r.src_pix; dst = r.dst_pix; n = r.n_pix;
    }
}

// grid (ceil(max count / 256), B).  T = int32 / float (W = 1), int2 (W = 2: x shifted), int4 (W = 4: x
// shifted); ids < 0 (absent neighbour) are never shifted.
template <typename T>
__global__ void __launch_bounds__(kThreads) rep_kernel(const T *__restrict__ in, T *__restrict__ out,
                                                       const RepImg *__restrict__ tab, int over, int shift_kind,
                                                       int d1) {
    const RepImg r = tab[blockIdx.y];
    int64_t src, dst, n;
    rep_range(r, over, d1, src, dst, n);
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    T v = in[src + i];
    if (shift_kind != kShiftNone) {
        int *x = reinterpret_cast<int *>(&v);  // first word: the id
        if (*x >= 0) *x += rep_shift(r, shift_kind, d1);
    }
    out[dst + i] = v;
}
// neighbour table [axis][vertex] of (n1, n2): both ids shifted
__global__ void __launch_bounds__(kThreads) rep_neigh_kernel(const int2 *__restrict__ in, int2 *__restrict__ out,
                                                             const RepImg *__restrict__ tab, int64_t M_in,
                                                             int64_t M_out, int d1) {
    const RepImg r = tab[blockIdx.y];
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= (int64_t)r.n_vert * d1) return;
    const int j = (int)(i / r.n_vert);
    const int64_t v = i - (int64_t)j * r.n_vert;
    int2 nb = in[(int64_t)j * M_in + r.src_vert + v];
    const int sh = r.dst_vert - r.src_vert;
    if (nb.x >= 0) nb.x += sh;
    if (nb.y >= 0) nb.y += sh;
    out[(int64_t)j * M_out + r.dst_vert + v] = nb;
}
__global__ void set_i32_kernel(int32_t *p, int32_t v) { *p = v; }
}  // namespace

void launch_replicate_lattice(const Lattice &one, const BatchGeom &og, const float *norm_one, const BatchGeom &g,
                              const std::vector<int> &src, Lattice &out, float *norm_out, cudaStream_t s) {
    const int d1 = one.d + 1;
    const int B = g.B;
    ProfScope prof(DCRF_K_BUILD_REPL, one.d, s);
    DCRF_REQUIRE(g.Ntot * d1 < (int64_t)2147483000, DCRF_EINVAL, "batch too large: N*(d+1) must stay below 2^31");
    out.d = one.d;
    out.E = g.Ntot * d1;
    out.vert_start.assign(B + 1, 0);
    static_assert(sizeof(RepImg) == 6 * sizeof(int32_t), "RepImg is six ints");
    out.h_rep.assign((size_t)B * 6, 0);  // staging of an asynchronous upload: lives with the lattice
    RepImg *tab = reinterpret_cast<RepImg *>(out.h_rep.data());
    int64_t max_pix = 1, max_vert = 1;
    for (int b = 0; b < B; b++) {
        const int u = src[b];
        const int64_t mv = one.vert_start[u + 1] - one.vert_start[u];
        out.vert_start[b + 1] = out.vert_start[b] + mv;
        tab[b] = RepImg{(int)og.pix_start[u], (int)g.pix_start[b], (int)(og.pix_start[u + 1] - og.pix_start[u]),
                        (int)one.vert_start[u], (int)out.vert_start[b], (int)mv};
        max_pix = std::max<int64_t>(max_pix, tab[b].n_pix);
        max_vert = std::max<int64_t>(max_vert, mv);
    }
    out.M = out.vert_start[B];
    DevBuf<RepImg> d_tab;
    d_tab.alloc(B, s);
    DCRF_CUDA(copy_h2d(d_tab.p, tab, sizeof(RepImg) * B, s));
    out.offset.alloc(out.E, s);
    out.bary.alloc(out.E, s);
    out.neigh.alloc((size_t)out.M * d1, s);
    out.vkeys.alloc((size_t)out.M * 8, s);
    out.csr_start.alloc(out.M + 1, s);
    out.csr_pix.alloc(out.E, s);
    out.csr_w.alloc(out.E, s);
    out.ent.alloc(out.E, s);
    out.table_mode = one.table_mode;
    out.long_row_cap = one.long_row_cap;
    auto grid = [&](int64_t n) { return dim3(ceil_div(n, kThreads), B); };
    const dim3 ge = grid(max_pix * d1), gv = grid(max_vert), gp = grid(max_pix);
    rep_kernel<int32_t><<<ge, kThreads, 0, s>>>(one.offset.p, out.offset.p, d_tab.p, kOverEnt, kShiftVert, d1);
    DCRF_LAUNCHED();
    rep_kernel<float><<<ge, kThreads, 0, s>>>(one.bary.p, out.bary.p, d_tab.p, kOverEnt, kShiftNone, d1);
    DCRF_LAUNCHED();
    rep_neigh_kernel<<<grid(max_vert * d1), kThreads, 0, s>>>(one.neigh.p, out.neigh.p, d_tab.p, one.M, out.M, d1);
    DCRF_LAUNCHED();
    rep_kernel<int4><<<gv, kThreads, 0, s>>>(reinterpret_cast<const int4 *>(one.vkeys.p),
                                            reinterpret_cast<int4 *>(out.vkeys.p), d_tab.p, kOverVert, kShiftNone, d1);
    DCRF_LAUNCHED();
    // row starts: the rows of an image are the entries of that image, so they move by its entry shift
    rep_kernel<int32_t><<<gv, kThreads, 0, s>>>(one.csr_start.p, out.csr_start.p, d_tab.p, kOverVert, kShiftEnt, d1);
    DCRF_LAUNCHED();
    set_i32_kernel<<<1, 1, 0, s>>>(out.csr_start.p + out.M, (int32_t)out.E);
    DCRF_LAUNCHED();
    rep_kernel<int32_t><<<ge, kThreads, 0, s>>>(one.csr_pix.p, out.csr_pix.p, d_tab.p, kOverEnt, kShiftPix, d1);
    DCRF_LAUNCHED();
    rep_kernel<float><<<ge, kThreads, 0, s>>>(one.csr_w.p, out.csr_w.p, d_tab.p, kOverEnt, kShiftNone, d1);
    DCRF_LAUNCHED();
    rep_kernel<int2><<<ge, kThreads, 0, s>>>(one.ent.p, out.ent.p, d_tab.p, kOverEnt, kShiftVert, d1);
    DCRF_LAUNCHED();
    if (one.table_mode == kTablesRef) {
        out.csr_ent4.alloc(out.E, s);
        rep_kernel<int4><<<ge, kThreads, 0, s>>>(one.csr_ent4.p, out.csr_ent4.p, d_tab.p, kOverEnt, kShiftPix, d1);
    } else {
        out.csr_ent.alloc(out.E, s);
        rep_kernel<int2><<<ge, kThreads, 0, s>>>(one.csr_ent.p, out.csr_ent.p, d_tab.p, kOverEnt, kShiftPix, d1);
    }
    DCRF_LAUNCHED();
    if (one.row_counter.p) out.row_counter.alloc(2, s);
    launch_find_long_rows(out, s);
    if (norm_one && norm_out) {
        rep_kernel<float><<<gp, kThreads, 0, s>>>(norm_one, norm_out, d_tab.p, kOverPix, kShiftNone, d1);
        DCRF_LAUNCHED();
    }
}

}  // namespace dcrf
