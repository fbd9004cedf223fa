// api.cu -- the extern "C" boundary (include/dcrf_b200.h) and the host-side orchestration of the
// mean-field loop.  One handle = a batch of independent images sharing L; all kernels run over the
// concatenated pixel / vertex arrays of the batch on the handle's stream.
#include <math.h>
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <memory>
#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"

namespace dcrf {

std::atomic<int64_t> g_launches{0};
std::atomic<int64_t> g_h2d_bytes{0}, g_d2h_bytes{0};
static thread_local std::string t_error;
thread_local Profiler *t_prof = nullptr;
void set_error(const std::string &msg) { t_error = msg; }

namespace {
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) DCRF_CUDA(cudaSetDevice(dev));
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// The streams that accompany one primary stream: side streams for the concurrent pairwise filters
// and an upload stream, plus the primary itself when the library created it.  Reference counted --
// every handle that runs on the primary holds a reference, so a handle may outlive the thread that
// created it (a Python object collected or closed from another thread, a ThreadPool worker that
// exits): the streams and their memory pools go away with the LAST holder, never under a live handle.
struct StreamSet {
    int dev = 0;
    cudaStream_t primary = nullptr;
    bool owns_primary = false;
    std::mutex mu;
    cudaStream_t side[kMaxPairwise - 1] = {};
    cudaStream_t upload = nullptr;
    cudaStream_t get_side(int k) {
        std::lock_guard<std::mutex> lock(mu);
        if (!side[k]) {
            DeviceGuard g(dev);
            DCRF_CUDA(cudaStreamCreateWithFlags(&side[k], cudaStreamNonBlocking));
        }
        return side[k];
    }
    cudaStream_t get_upload() {
        std::lock_guard<std::mutex> lock(mu);
        if (!upload) {
            DeviceGuard g(dev);
            DCRF_CUDA(cudaStreamCreateWithFlags(&upload, cudaStreamNonBlocking));
        }
        return upload;
    }
    ~StreamSet() {
        try {
            DeviceGuard g(dev);
            for (auto st : side)
                if (st) {
                    cudaStreamSynchronize(st);
                    stream_pool_release(st);
                    cudaStreamDestroy(st);
                }
            if (upload) {
                cudaStreamSynchronize(upload);
                stream_pool_release(upload);
                cudaStreamDestroy(upload);
            }
            if (owns_primary && primary) {
                cudaStreamSynchronize(primary);
                stream_pool_release(primary);
                cudaStreamDestroy(primary);
            }
        } catch (...) {
        }
    }
};

std::shared_ptr<StreamSet> new_owned_set(int dev) {
    auto set = std::make_shared<StreamSet>();
    set->dev = dev;
    set->owns_primary = true;
    DCRF_CUDA(cudaStreamCreateWithFlags(&set->primary, cudaStreamNonBlocking));
    return set;
}

// the calling thread's persistent stream set per device: a handle created without a caller stream
// runs on it, so consecutive handles of a thread reuse the thread's pool blocks in plain stream order
// and no stream is created / destroyed per image
thread_local std::shared_ptr<StreamSet> t_sets[64];
std::shared_ptr<StreamSet> thread_set(int dev) {
    DCRF_REQUIRE(dev >= 0 && dev < 64, DCRF_EINVAL, "device index out of range");
    if (!t_sets[dev]) t_sets[dev] = new_owned_set(dev);
    return t_sets[dev];
}

// sets of caller-provided primaries (torch streams, dcrf_stream_create): kept in a registry so that
// handle after handle on the same stream reuses the same side / upload streams
// (The registries of this file are never destroyed: static destructors run in reverse order of definition
// at process exit, and a StreamSet dying there would release its pools into a map that is already gone
// -- glibc "double free or corruption" once side streams own pools.  The process is ending; the driver
// reclaims streams and pools.)
std::mutex &g_sets_mu = *new std::mutex();
auto &g_sets = *new std::map<std::pair<int, cudaStream_t>, std::shared_ptr<StreamSet>>();
std::shared_ptr<StreamSet> caller_set(int dev, cudaStream_t primary) {
    std::lock_guard<std::mutex> lock(g_sets_mu);
    auto &slot = g_sets[std::make_pair(dev, primary)];
    if (!slot) {
        slot = std::make_shared<StreamSet>();
        slot->dev = dev;
        slot->primary = primary;
    }
    return slot;
}
}  // namespace

double trace_now() {
    static const bool on = getenv("DCRF_TRACE") != nullptr;
    if (!on) return 0.0;
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
void trace_slow(const char *what, double t0, size_t bytes) {
    if (t0 == 0.0) return;
    const double dt = trace_now() - t0;
    if (dt > 2.0) fprintf(stderr, "[dcrf trace] %s blocked %.1f ms (%zu bytes)\n", what, dt, bytes);
}

static std::mutex &g_pool_mu = *new std::mutex();
static auto &g_pools = *new std::map<std::pair<int, cudaStream_t>, cudaMemPool_t>();  // key: (device, stream); never destroyed

cudaMemPool_t stream_pool(cudaStream_t stream) {
    int dev = 0;
    DCRF_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_pool_mu);
    auto it = g_pools.find(std::make_pair(dev, stream));
    if (it != g_pools.end()) return it->second;
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    cudaMemPool_t pool;
    DCRF_CUDA(cudaMemPoolCreate(&pool, &props));
    uint64_t thr = UINT64_MAX;  // keep freed blocks cached: every image needs fresh lattice buffers
    DCRF_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    g_pools[std::make_pair(dev, stream)] = pool;
    return pool;
}

// Pre-grow a stream's pool to `want` bytes in ONE contiguous block (allocate + free: the release
// threshold keeps it).  Without slack the pool settles at a reserved size just above a handle's peak
// and, its free space being fragmented, cudaMallocFromPoolAsync then re-maps physical memory behind
// the larger requests -- measured (DCRF_TRACE=1, 8 ADP 1088^2 images per handle): stalls of 10-900 ms
// in one step out of three, with the reserved size constant.  One early growth to the estimated
// footprint of the handle replaces them.  Best effort: a failed reservation is not an error.
static void pool_reserve(cudaStream_t stream, size_t want) {
    static const double factor = [] {
        const char *e = getenv("DCRF_POOL_RESERVE_FACTOR");
        return e ? atof(e) : 1.0;
    }();
    want = (size_t)((double)want * factor);
    if (want == 0) return;
    cudaMemPool_t pool = stream_pool(stream);
    uint64_t reserved = 0;
    if (cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved) != cudaSuccess) return;
    if (reserved >= want) return;
    want += want / 4;  // head-room: batches of slightly varying size (a sweep over mixed image sizes) grow once
    void *p = nullptr;
    const double t0 = trace_now();
    if (cudaMallocFromPoolAsync(&p, want, pool, stream) != cudaSuccess) {
        cudaGetLastError();  // not enough memory for the slack: run without it
        return;
    }
    cudaFreeAsync(p, stream);
    trace_slow("pool_reserve (one-time growth of the stream's pool)", t0, want);
}

static void trim_all_pools() {
    std::lock_guard<std::mutex> lock(g_pool_mu);
    for (auto &kv : g_pools) cudaMemPoolTrimTo(kv.second, 0);
}

// (current device = the stream's device)
void stream_pool_release(cudaStream_t stream) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    std::lock_guard<std::mutex> lock(g_pool_mu);
    auto it = g_pools.find(std::make_pair(dev, stream));
    if (it == g_pools.end()) return;
    cudaMemPoolDestroy(it->second);
    g_pools.erase(it);
}

// A lattice whose first build half has been enqueued (on `bs`) and whose second half still has to be:
// everything the second half needs, kept alive until then.
struct PendingBuild {
    BuildState st;
    FeatureSpec fs;
    BatchGeom ug;                      // the distinct image sizes (position-only features, see add_pairwise)
    DevBuf<int> ug_w, ug_h, ug_ps;
    std::vector<int> ug_ps32, src;
    bool shared = false;
    Lattice single;
    DevBuf<float> norm_single;
    cudaStream_t bs = nullptr;         // the stream the build runs on
    cudaEvent_t ev = nullptr;          // end of the first half on bs
    int32_t *pinned = nullptr;         // page-locked [kPinnedInts] block for the vertex starts
    DevBuf<uint8_t> rgb_stage;         // staged copies of caller host memory the first half reads
    DevBuf<float> feat_stage;
    ~PendingBuild();                   // an abandoned build: waits for its stream, returns the pinned block
};

struct Pairwise {
    std::unique_ptr<PendingBuild> pending;
    cudaStream_t built_on = nullptr;   // side stream that owns the lattice's memory (nullptr: the handle's stream)
    Lattice lat;
    DevBuf<float> norm;    // [Ntot]; empty for NO_NORMALIZATION
    DevBuf<float> compat;  // diagonal: [Lp]; matrix: [Lp*Lp] zero padded
    std::vector<float> compat_host;
    DevBuf<float> valA, valB;  // lattice value ping-pong, [M * Lp]
    int ntype = DCRF_NORMALIZE_SYMMETRIC, ktype = DCRF_DIAG_KERNEL;
    int compat_kind = DCRF_COMPAT_POTTS;
    float potts_w = 0.f;
    bool stiff = false;  // narrow appearance kernel (see auto_arith)
};

}  // namespace dcrf

using namespace dcrf;

struct dcrf_handle {
    int device = 0;
    cudaStream_t stream = nullptr;          // = streams->primary
    std::shared_ptr<StreamSet> streams;     // keeps the primary's side / upload streams (and pools) alive
    int L = 0, Lp = 0;
    bool has_geom = false;  // 2-D image geometry available (Gaussian / bilateral features)
    BatchGeom geom;
    std::vector<int> ps32;
    DevBuf<int> d_w, d_h, d_pix_start;
    DevBuf<float> unary, Q;
    DevBuf<int> counters;  // row dispensers of the persistent mean-field kernel
    bool unary_set = false, q_valid = false;
    int arith = kArithFma;    // resolved arithmetic: kArithFma / kArithRef / kArithStrict
    int persistent = -1;      // DCRF_OPT_PERSISTENT: -1 = by problem size, 0 = never, 1 = whenever the model allows
    bool arith_auto = true;   // DCRF_ARITH_AUTO: `arith` follows the conditioning of the pairwise terms
    bool async_host = false;  // DCRF_OPT_ASYNC_HOST
    std::vector<std::unique_ptr<Pairwise>> pw;
    Profiler prof;
    // side streams: the filters of all pairwise terms but the last run concurrently with the last
    // one (the small, latency-bound Gaussian lattice hides behind the bandwidth-bound bilateral one)
    cudaEvent_t ev_fork = nullptr, ev_join[kMaxPairwise - 1] = {};
    // async-host mode: the unary upload (H2D + layout change) runs on the thread's upload stream so
    // that the lattice builds enqueued next on `stream` overlap it; joined before the unary is read
    DevBuf<float> upload_stage;
    cudaEvent_t ev_upload_begin = nullptr, ev_upload_end = nullptr;
    bool upload_pending = false;
};

namespace {

struct ProfGuard {
    explicit ProfGuard(dcrf_handle *h) { t_prof = &h->prof; }
    ~ProfGuard() { t_prof = nullptr; }
};

template <typename F>
int guarded(F &&f) {
    try {
        f();
        return DCRF_OK;
    } catch (const Error &e) {
        set_error(e.msg);
        return e.code;
    } catch (const std::bad_alloc &) {
        set_error("out of host memory");
        return DCRF_ENOMEM;
    } catch (const std::exception &e) {
        set_error(e.what());
        return DCRF_EINVAL;
    }
}

// end of a call that read or wrote caller HOST memory: block unless the handle is in async-host mode
void host_sync(dcrf_handle *h) {
    if (!h->async_host) DCRF_CUDA(cudaStreamSynchronize(h->stream));
}

// make `stream` wait for a unary upload still running on the upload stream, then recycle its staging
void join_upload(dcrf_handle *h) {
    if (!h->upload_pending) return;
    DCRF_CUDA(cudaStreamWaitEvent(h->stream, h->ev_upload_end, 0));
    h->upload_stage.release();
    h->upload_pending = false;
}

int64_t total_ln(const dcrf_handle *h) { return h->geom.Ntot * (int64_t)h->L; }

// default arithmetic of new handles: DCRF_ARITH_AUTO unless DCRF_ARITHMETIC says otherwise
int default_arith() {
    const char *e = getenv("DCRF_ARITHMETIC");
    if (!e || !*e || !strcmp(e, "auto") || !strcmp(e, "3")) return DCRF_ARITH_AUTO;
    if (!strcmp(e, "fma") || !strcmp(e, "0")) return kArithFma;
    if (!strcmp(e, "strict") || !strcmp(e, "2")) return kArithStrict;
    DCRF_REQUIRE(!strcmp(e, "reference") || !strcmp(e, "1"), DCRF_EINVAL,
                 "DCRF_ARITHMETIC must be auto, fma, reference or strict");
    return kArithRef;
}
void set_arith(dcrf_handle *h, int mode) {
    h->arith_auto = mode == DCRF_ARITH_AUTO;
    if (!h->arith_auto) h->arith = mode;
}
// DCRF_ARITH_AUTO: a narrow appearance kernel (colour bandwidth below 8 grey levels: the IRN label CRF,
// srgb = 5, and SEC's ADP-func test setting, srgb = 4) concentrates a pixel's message on a handful of
// neighbours; with a Potts weight of 10..25 the mean-field update is then expansive at bistable
// pixels and float rounding differences grow ~2.5x per iteration.  Such models
// run the reference arithmetic, whose marginals are bit-identical to the sequential evaluation; all
// others keep the faster FMA kernels, which stay within 1e-5 of it.
void auto_arith(dcrf_handle *h, const FeatureSpec &fs) {
    if (!h->arith_auto || h->arith != kArithFma) return;
    if (fs.mode == 1 && std::min(fs.s[2], std::min(fs.s[3], fs.s[4])) < 8.0f) h->arith = kArithRef;
}

void create_common(int B, const int *w, const int *hgt, bool has_geom, int L, int device, void *stream,
                   dcrf_t **out) {
    DCRF_REQUIRE(out != nullptr, DCRF_EINVAL, "out handle pointer is NULL");
    DCRF_REQUIRE(B >= 1, DCRF_EINVAL, "n_images must be >= 1");
    DCRF_REQUIRE(L >= 0, DCRF_EINVAL, "n_labels must be >= 0");
    DCRF_REQUIRE(L <= 128, DCRF_EINVAL, "n_labels > 128 is not supported");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        throw Error{DCRF_ECUDA, std::string("no CUDA device: dcrf_b200 has no CPU fallback (") +
                                    cudaGetErrorString(ce) + ")"};
    if (device < 0) DCRF_CUDA(cudaGetDevice(&device));
    DCRF_REQUIRE(device < ndev, DCRF_EINVAL, "device index out of range");
    std::unique_ptr<dcrf_handle> h(new dcrf_handle());
    h->device = device;
    DeviceGuard guard(device);
    if (stream == DCRF_STREAM_DEDICATED) h->streams = new_owned_set(device);  // lives and dies with this handle
    else if (stream) h->streams = caller_set(device, (cudaStream_t)stream);
    else h->streams = thread_set(device);  // persistent; shared with the calling thread's other handles
    h->stream = h->streams->primary;
    h->L = L;
    h->Lp = ((L + 3) / 4) * 4;
    h->has_geom = has_geom;
    set_arith(h.get(), default_arith());
    BatchGeom &g = h->geom;
    g.B = B;
    g.w.resize(B);
    g.h.resize(B);
    g.pix_start.assign(B + 1, 0);
    std::vector<int> &ps32 = h->ps32;  // staging of an asynchronous upload: lives as long as the handle
    ps32.assign(B + 1, 0);
    for (int b = 0; b < B; b++) {
        DCRF_REQUIRE(w[b] >= 1 && hgt[b] >= 1, DCRF_EINVAL, "image width/height must be >= 1");
        g.w[b] = w[b];
        g.h[b] = hgt[b];
        g.pix_start[b + 1] = g.pix_start[b] + (int64_t)w[b] * hgt[b];
        DCRF_REQUIRE(g.pix_start[b + 1] < (int64_t)1 << 30, DCRF_EINVAL, "batch has too many pixels");
        ps32[b + 1] = (int)g.pix_start[b + 1];
    }
    g.Ntot = g.pix_start[B];
    h->d_w.alloc(B, h->stream);
    h->d_h.alloc(B, h->stream);
    h->d_pix_start.alloc(B + 1, h->stream);
    DCRF_CUDA(copy_h2d(h->d_w.p, g.w.data(), sizeof(int) * B, h->stream));
    DCRF_CUDA(copy_h2d(h->d_h.p, g.h.data(), sizeof(int) * B, h->stream));
    DCRF_CUDA(copy_h2d(h->d_pix_start.p, ps32.data(), sizeof(int) * (B + 1), h->stream));
    g.d_w = h->d_w.p;
    g.d_h = h->d_h.p;
    g.d_pix_start = h->d_pix_start.p;
    // estimated device footprint of a Gaussian + bilateral model on this batch (lattices of up to ~4
    // vertices per pixel in total; wsss.py sizes its batches with the same figure)
    pool_reserve(h->stream, (size_t)g.Ntot * (size_t)(40 * std::max(h->Lp, 4) + 830));
    if (h->Lp > 0) {
        h->unary.alloc((size_t)g.Ntot * h->Lp, h->stream);
        h->Q.alloc((size_t)g.Ntot * h->Lp, h->stream);
        DCRF_CUDA(cudaMemsetAsync(h->unary.p, 0, sizeof(float) * g.Ntot * h->Lp, h->stream));
    }
    *out = h.release();
}

// bring `count` elements to the device if they are on the host; returns the device pointer
template <typename T>
const T *to_device(dcrf_handle *h, const T *src, size_t count, int on_device, DevBuf<T> &stage) {
    if (on_device) return src;
    stage.alloc(count, h->stream);
    const double t0 = trace_now();
    DCRF_CUDA(copy_h2d(stage.p, src, sizeof(T) * count, h->stream));
    trace_slow("cudaMemcpyAsync H2D (enqueue)", t0, sizeof(T) * count);
    return stage.p;
}

// splat + (d+1) blurs of pairwise k applied to `in` (pixel-major Lp); returns the blurred buffer
const float *filter_to_lattice(dcrf_handle *h, Pairwise &p, const float *in, int Lp, bool pre_norm,
                               bool seq, float *bufA, float *bufB, bool fast = false,
                               cudaStream_t st = nullptr) {
    if (!st) st = h->stream;
    if (fast) launch_splat_fast(p.lat, in, bufA, Lp, st);  // packed tables (pre-norm inside them)
    else launch_splat(p.lat, in, pre_norm ? p.norm.p : nullptr, bufA, Lp, st);
    float *cur = bufA, *nxt = bufB;
    for (int j = 0; j <= p.lat.d; j++) {
        launch_blur(p.lat, j, cur, nxt, Lp, seq, st);
        std::swap(cur, nxt);
    }
    return cur;
}

bool pre_norm(int ntype) { return ntype == DCRF_NORMALIZE_SYMMETRIC || ntype == DCRF_NORMALIZE_BEFORE; }
bool post_norm(int ntype) { return ntype == DCRF_NORMALIZE_SYMMETRIC || ntype == DCRF_NORMALIZE_AFTER; }

// packed entry tables of the handle's arithmetic mode (norm: the kernel's [Ntot] vector or nullptr)
void pack_tables(dcrf_handle *h, Lattice &lat, int ntype, const float *norm, cudaStream_t s) {
    if (h->arith == kArithFma)
        launch_pack_fast_tables(lat, pre_norm(ntype) ? norm : nullptr, post_norm(ntype) ? norm : nullptr, s);
    else
        launch_pack_ref_tables(lat, pre_norm(ntype) ? norm : nullptr, h->arith == kArithStrict ? INT_MAX : 0, s);
}
int wanted_tables(const dcrf_handle *h) { return h->arith == kArithFma ? kTablesFma : kTablesRef; }

// ---- concurrent builds of small problems ----
// One VOC image (or a batch of 41x41 SEC maps) builds each lattice in ~0.3 ms of ~30 dependent small
// launches around one host synchronisation: latency, not throughput.  For such handles the FIRST half of
// a build is only enqueued, on the side stream of its kernel, and add_pairwise returns; the second
// halves run when the lattices are first needed (inference, export, ...).  The Gaussian and the bilateral
// lattice of a model are then built side by side, and the host waits once per lattice for a vertex count
// that is usually already there.  Larger problems fill the GPU with every kernel and keep the plain
// single-stream build (and its single memory pool).
constexpr int64_t kConcurrentBuildMaxPixels = 1000000;
bool concurrent_builds_enabled() {
    static const bool on = [] {
        const char *e = getenv("DCRF_CONCURRENT_BUILDS");
        return !e || atoi(e) != 0;
    }();
    return on;
}
constexpr int kPinnedInts = 4096;
std::mutex &g_pinned_mu = *new std::mutex();
auto &g_pinned_free = *new std::vector<int32_t *>();  // page-locked blocks are expensive to create: recycled for the life of the process
int32_t *pinned_get() {
    {
        std::lock_guard<std::mutex> lock(g_pinned_mu);
        if (!g_pinned_free.empty()) {
            int32_t *b = g_pinned_free.back();
            g_pinned_free.pop_back();
            return b;
        }
    }
    int32_t *b = nullptr;
    DCRF_CUDA(cudaMallocHost((void **)&b, sizeof(int32_t) * kPinnedInts));
    return b;
}
void pinned_put(int32_t *b) {
    if (!b) return;
    std::lock_guard<std::mutex> lock(g_pinned_mu);
    g_pinned_free.push_back(b);
}
}  // namespace
namespace dcrf {
PendingBuild::~PendingBuild() {
    if (ev || pinned) {  // never finished (error path, handle destroyed before its first use)
        if (bs) cudaStreamSynchronize(bs);  // its kernels read rgb_stage / write `pinned`
        if (ev) cudaEventDestroy(ev);
        pinned_put(pinned);
    }
}
}  // namespace dcrf
namespace {

// second half of a build + norm + packed tables + replication + value buffers, on the build's stream
void finish_pairwise(dcrf_handle *h, Pairwise &p) {
    if (!p.pending) return;
    PendingBuild &pb = *p.pending;
    cudaStream_t s = pb.bs;
    const int Lp = h->Lp;
    const int64_t Ntot = h->geom.Ntot;
    if (pb.ev) {
        DCRF_CUDA(cudaEventSynchronize(pb.ev));
    } else {
        DCRF_CUDA(cudaStreamSynchronize(s));
    }
    const BatchGeom &bg = pb.shared ? pb.ug : h->geom;
    Lattice &lat = pb.shared ? pb.single : p.lat;
    DevBuf<float> &norm = pb.shared ? pb.norm_single : p.norm;
    build_lattice_finish(bg, pb.fs, lat, pb.st, s);
    // A.5: norm = filter(ones) through the value_size = 1 path
    if (p.ntype != DCRF_NO_NORMALIZATION) {
        ProfScope prof(DCRF_K_BUILD_NORM, lat.d, s);
        norm.alloc(bg.Ntot, s);
        launch_kernel_norm(lat, bg.Ntot, p.ntype, norm.p, s);
    }
    {
        ProfScope prof(DCRF_K_BUILD_CSR, lat.d, s);
        pack_tables(h, lat, p.ntype, norm.p, s);
    }
    if (pb.shared) {
        if (norm.p) p.norm.alloc(Ntot, s);
        launch_replicate_lattice(pb.single, pb.ug, norm.p, h->geom, pb.src, p.lat, p.norm.p, s);
    }
    p.valA.alloc((size_t)p.lat.M * Lp, s);
    p.valB.alloc((size_t)p.lat.M * Lp, s);
    if (s != h->stream) {  // the handle's stream continues after the build
        if (!pb.ev) DCRF_CUDA(cudaEventCreateWithFlags(&pb.ev, cudaEventDisableTiming));
        DCRF_CUDA(cudaEventRecord(pb.ev, s));
        DCRF_CUDA(cudaStreamWaitEvent(h->stream, pb.ev, 0));
        p.built_on = s;
    }
    if (pb.ev) cudaEventDestroy(pb.ev);  // released when the recorded work completes
    pb.ev = nullptr;
    pinned_put(pb.pinned);
    pb.pinned = nullptr;
    p.pending.reset();  // temporaries and staged inputs: freed in the order of the streams they were allocated on
}

void finish_builds(dcrf_handle *h) {
    for (auto &p : h->pw) finish_pairwise(h, *p);
}

void add_pairwise(dcrf_handle *h, const FeatureSpec &fs, int compat_kind, const float *compat, int ktype,
                  int ntype, DevBuf<uint8_t> *rgb_stage = nullptr, DevBuf<float> *feat_stage = nullptr,
                  bool caller_device_input = false) {
    DCRF_REQUIRE(h->L >= 1, DCRF_ESTATE, "model has no labels");
    DCRF_REQUIRE((int)h->pw.size() < kMaxPairwise, DCRF_EINVAL, "too many pairwise terms (max 4)");
    DCRF_REQUIRE(ktype >= DCRF_CONST_KERNEL && ktype <= DCRF_FULL_KERNEL, DCRF_EINVAL, "bad kernel type");
    DCRF_REQUIRE(ntype >= DCRF_NO_NORMALIZATION && ntype <= DCRF_NORMALIZE_SYMMETRIC, DCRF_EINVAL,
                 "bad normalization type");
    DCRF_REQUIRE(compat_kind >= DCRF_COMPAT_POTTS && compat_kind <= DCRF_COMPAT_MATRIX, DCRF_EINVAL,
                 "bad compatibility kind");
    DCRF_REQUIRE(compat != nullptr, DCRF_EINVAL, "compat is NULL");
    cudaStream_t s = h->stream;
    const int L = h->L, Lp = h->Lp;
    const int64_t Ntot = h->geom.Ntot;
    std::unique_ptr<Pairwise> p(new Pairwise());
    p->stiff = fs.mode == 1 && std::min(fs.s[2], std::min(fs.s[3], fs.s[4])) < 8.0f;
    auto_arith(h, fs);
    p->ktype = ktype;  // CONST / DIAG / FULL are all the identity feature map at default parameters
    p->ntype = ntype;
    p->compat_kind = compat_kind;
    if (compat_kind == DCRF_COMPAT_POTTS) {
        p->potts_w = compat[0];
    } else if (compat_kind == DCRF_COMPAT_DIAGONAL) {
        std::vector<float> &c = p->compat_host;  // staging of an asynchronous upload
        c.assign(Lp, 0.f);
        for (int l = 0; l < L; l++) c[l] = compat[l];
        p->compat.alloc(Lp, s);
        DCRF_CUDA(copy_h2d(p->compat.p, c.data(), sizeof(float) * Lp, s));
        DCRF_CUDA(cudaStreamSynchronize(s));
    } else {
        // [EXT] MatrixCompatibility stores 0.5 * (m + m^T)
        std::vector<float> &c = p->compat_host;
        c.assign((size_t)Lp * Lp, 0.f);
        for (int a = 0; a < L; a++)
            for (int b = 0; b < L; b++) c[(size_t)a * Lp + b] = 0.5f * (compat[a * L + b] + compat[b * L + a]);
        p->compat.alloc((size_t)Lp * Lp, s);
        DCRF_CUDA(copy_h2d(p->compat.p, c.data(), sizeof(float) * Lp * Lp, s));
        DCRF_CUDA(cudaStreamSynchronize(s));
    }
    // A lattice whose features depend on the pixel position only (the Gaussian kernel) is the same
    // for every image of a given size: it is built (and its norm filtered) ONCE per distinct size of
    // the batch and replicated with per-image id offsets (one size: bench.py's batches; a handful of
    // sizes: a batch of PASCAL VOC val images).
    p->pending.reset(new PendingBuild());
    PendingBuild &pb = *p->pending;
    pb.fs = fs;
    pb.src.assign(h->geom.B, 0);
    BatchGeom &ug = pb.ug;   // the distinct sizes, in order of first appearance
    if (fs.mode == 0 && h->geom.B > 1) {
        std::map<std::pair<int, int>, int> seen;
        for (int b = 0; b < h->geom.B; b++) {
            auto key = std::make_pair(h->geom.w[b], h->geom.h[b]);
            auto it = seen.find(key);
            if (it == seen.end()) {
                it = seen.emplace(key, (int)ug.w.size()).first;
                ug.w.push_back(key.first);
                ug.h.push_back(key.second);
            }
            pb.src[b] = it->second;
        }
        pb.shared = (int)ug.w.size() < h->geom.B;
    }
    // small problems: first half on the kernel's side stream, second half when the lattice is needed
    const int k = (int)h->pw.size();
    const bool concurrent = Ntot <= kConcurrentBuildMaxPixels && !h->prof.on && k < kMaxPairwise - 1 &&
                            h->geom.B + 1 <= kPinnedInts && concurrent_builds_enabled();
    cudaStream_t bs = concurrent ? h->streams->get_side(k) : s;
    pb.bs = bs;
    if (concurrent) {
        // everything enqueued on the handle's stream so far (geometry uploads, staged inputs) comes first
        DCRF_CUDA(cudaEventCreateWithFlags(&pb.ev, cudaEventDisableTiming));
        DCRF_CUDA(cudaEventRecord(pb.ev, s));
        DCRF_CUDA(cudaStreamWaitEvent(bs, pb.ev, 0));
        pb.pinned = pinned_get();
        if (rgb_stage) pb.rgb_stage = std::move(*rgb_stage);
        if (feat_stage) pb.feat_stage = std::move(*feat_stage);
    }
    if (pb.shared) {
        ug.B = (int)ug.w.size();
        ug.pix_start.assign(ug.B + 1, 0);
        pb.ug_ps32.assign(ug.B + 1, 0);
        for (int u = 0; u < ug.B; u++) {
            ug.pix_start[u + 1] = ug.pix_start[u] + (int64_t)ug.w[u] * ug.h[u];
            pb.ug_ps32[u + 1] = (int)ug.pix_start[u + 1];
        }
        ug.Ntot = ug.pix_start[ug.B];
        pb.ug_w.alloc(ug.B, bs);
        pb.ug_h.alloc(ug.B, bs);
        pb.ug_ps.alloc(ug.B + 1, bs);
        DCRF_CUDA(copy_h2d(pb.ug_w.p, ug.w.data(), sizeof(int) * ug.B, bs));
        DCRF_CUDA(copy_h2d(pb.ug_h.p, ug.h.data(), sizeof(int) * ug.B, bs));
        DCRF_CUDA(copy_h2d(pb.ug_ps.p, pb.ug_ps32.data(), sizeof(int) * (ug.B + 1), bs));
        ug.d_w = pb.ug_w.p;
        ug.d_h = pb.ug_h.p;
        ug.d_pix_start = pb.ug_ps.p;
    }
    const BatchGeom &bg = pb.shared ? ug : h->geom;
    Lattice &lat = pb.shared ? pb.single : p->lat;
    build_lattice_begin(bg, fs, lat, pb.st, pb.pinned, bs);
    if (concurrent) {
        DCRF_CUDA(cudaEventRecord(pb.ev, bs));
        // memory of the CALLER on the device (a torch tensor) is only read by the first kernels: it must
        // not be released to its allocator before they ran
        if (caller_device_input) DCRF_CUDA(cudaEventSynchronize(pb.ev));
        h->pw.push_back(std::move(p));
        return;
    }
    Pairwise &pr = *p;
    h->pw.push_back(std::move(p));
    finish_pairwise(h, pr);
}


SliceTerm make_term(Pairwise &p, const float *blurred) {
    SliceTerm t;
    t.offset = p.lat.offset.p;
    t.bary = p.lat.bary.p;
    t.val = blurred;
    t.norm = post_norm(p.ntype) ? p.norm.p : nullptr;
    t.compat = p.compat.p;
    t.potts_w = p.potts_w;
    t.alpha = 1.0f / (1.0f + powf(2.0f, (float)-p.lat.d));
    t.d = p.lat.d;
    t.compat_kind = p.compat_kind;
    t.ent = p.lat.ent.p;
    return t;
}

void start_inference(dcrf_handle *h) {
    DCRF_REQUIRE(h->L >= 1, DCRF_ESTATE, "model has no labels");
    join_upload(h);
    SliceArgs a;
    memset(&a, 0, sizeof(a));
    a.n_terms = 0;
    a.seq = h->L <= 2 ? 1 : 0;
    a.fast = h->L <= 2 ? kSliceLiteral : (h->arith == kArithFma ? kSliceFma : kSliceRef);
    launch_slice_softmax(a, h->unary.p, h->Q.p, h->geom.Ntot, h->L, h->Lp, h->stream);
    h->q_valid = true;
}

void step_inference(dcrf_handle *h) {
    DCRF_REQUIRE(h->q_valid, DCRF_ESTATE, "stepInference before startInference");
    finish_builds(h);
    const bool seq = h->L <= 2;
    // packed-table kernels index rows with 32 bits: (largest row index) * (float4 per row) < 2^32
    int64_t max_rows = h->geom.Ntot;
    for (auto &p : h->pw) max_rows = std::max<int64_t>(max_rows, p->lat.M);
    const bool fast = !seq && max_rows * (int64_t)(h->Lp / 4) < ((int64_t)1 << 32) &&
                      h->geom.Ntot * (int64_t)(h->Lp / 4) < ((int64_t)1 << 31);
    SliceArgs a;
    memset(&a, 0, sizeof(a));
    a.seq = seq ? 1 : 0;
    a.fast = !fast ? kSliceLiteral : (h->arith == kArithFma ? kSliceFma : kSliceRef);
    a.max_rows = max_rows;
    const int n = (int)h->pw.size();
    if (fast)  // the arithmetic option was changed after the kernel was added: repack its tables
        for (auto &p : h->pw)
            if (p->lat.table_mode != wanted_tables(h)) pack_tables(h, p->lat, p->ntype, p->norm.p, h->stream);
    // per-kernel profiling wants serialised kernels; otherwise fork the first n-1 terms
    const bool overlap = n >= 2 && !h->prof.on;
    if (overlap) {
        if (!h->ev_fork) {
            const double t0 = trace_now();
            DCRF_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
            for (int k = 0; k < kMaxPairwise - 1; k++)
                DCRF_CUDA(cudaEventCreateWithFlags(&h->ev_join[k], cudaEventDisableTiming));
            trace_slow("side stream/event creation", t0, 0);
        }
        DCRF_CUDA(cudaEventRecord(h->ev_fork, h->stream));
    }
    for (int k = 0; k < n; k++) {
        Pairwise &p = *h->pw[k];
        cudaStream_t st = (overlap && k < n - 1) ? h->streams->get_side(k) : h->stream;
        if (st != h->stream) DCRF_CUDA(cudaStreamWaitEvent(st, h->ev_fork, 0));
        const float *blurred =
            filter_to_lattice(h, p, h->Q.p, h->Lp, pre_norm(p.ntype), seq, p.valA.p, p.valB.p, fast, st);
        if (st != h->stream) DCRF_CUDA(cudaEventRecord(h->ev_join[k], st));
        a.term[a.n_terms++] = make_term(p, blurred);
    }
    if (overlap)
        for (int k = 0; k < n - 1; k++) DCRF_CUDA(cudaStreamWaitEvent(h->stream, h->ev_join[k], 0));
    launch_slice_softmax(a, h->unary.p, h->Q.p, h->geom.Ntot, h->L, h->Lp, h->stream);
}

void emit_q(dcrf_handle *h, float *Q_out, int on_device) {
    DCRF_REQUIRE(Q_out != nullptr, DCRF_EINVAL, "Q_out is NULL");
    const int64_t n = total_ln(h);
    if (on_device) {
        launch_pm_to_ln(h->Q.p, Q_out, h->geom, h->L, h->Lp, h->stream);
    } else {
        DevBuf<float> stage;
        stage.alloc(n, h->stream);
        launch_pm_to_ln(h->Q.p, stage.p, h->geom, h->L, h->Lp, h->stream);
        DCRF_CUDA(copy_d2h(Q_out, stage.p, sizeof(float) * n, h->stream));
        host_sync(h);
    }
}

// argmax of the running Q as int32 or uint8 labels, to device or host memory
template <typename T>
void emit_labels(dcrf_handle *h, T *labels_out, int on_device) {
    const int64_t N = h->geom.Ntot;
    auto launch = [&](T *dst) {
        if (sizeof(T) == 1) launch_argmax_u8(h->Q.p, (uint8_t *)dst, N, h->L, h->Lp, h->stream);
        else launch_argmax(h->Q.p, (int32_t *)dst, N, h->L, h->Lp, h->stream);
    };
    if (on_device) {
        launch(labels_out);
    } else {
        DevBuf<T> stage;
        stage.alloc(N, h->stream);
        launch(stage.p);
        DCRF_CUDA(copy_d2h(labels_out, stage.p, sizeof(T) * N, h->stream));
        host_sync(h);
    }
}

// Opt-in: the whole of inference(n) as one cooperative launch (filter.cu,
// mean_field_persistent_kernel).  Measured on B200 it LOSES against the launch-per-phase path it was
// meant to beat (one VOC image: 2.54 vs 1.90 ms per step, 32 SEC maps: 1.41 vs 1.21 ms): ~9 grid
// barriers per iteration cost as much as the launches they replace, and the fused kernel's 64-105
// registers leave half the resident warps of the stand-alone latency-bound kernels.  Kept behind
// DCRF_OPT_PERSISTENT = 1 (or DCRF_PERSISTENT_MAX_PIXELS > 0) with its bit-identity tests.
bool try_persistent(dcrf_handle *h, int n_iter) {
    static const int64_t max_pix = [] {
        const char *e = getenv("DCRF_PERSISTENT_MAX_PIXELS");
        return e ? (int64_t)atoll(e) : (int64_t)0;  // default: off (measured slower on B200)
    }();
    const int n = (int)h->pw.size();
    if (h->persistent == 0 || (h->persistent < 0 && (max_pix <= 0 || h->geom.Ntot > max_pix))) return false;
    if (h->prof.on || h->L <= 2 || n == 0 || n_iter < 1 || h->Lp > 32) return false;
    int64_t max_rows = h->geom.Ntot;
    for (auto &p : h->pw) {
        if (p->compat_kind != DCRF_COMPAT_POTTS || p->lat.M == 0) return false;
        max_rows = std::max<int64_t>(max_rows, p->lat.M);
    }
    if (max_rows * (int64_t)(h->Lp / 4) >= ((int64_t)1 << 31)) return false;
    SliceArgs a;
    memset(&a, 0, sizeof(a));
    a.fast = h->arith == kArithFma ? kSliceFma : kSliceRef;
    a.max_rows = max_rows;
    const Lattice *lats[kMaxPairwise];
    float *va[kMaxPairwise], *vb[kMaxPairwise];
    for (int k = 0; k < n; k++) {
        Pairwise &p = *h->pw[k];
        if (p.lat.table_mode != wanted_tables(h)) pack_tables(h, p.lat, p.ntype, p.norm.p, h->stream);
        a.term[a.n_terms++] = make_term(p, nullptr);
        lats[k] = &p.lat;
        va[k] = p.valA.p;
        vb[k] = p.valB.p;
    }
    h->counters.alloc((size_t)2 * n_iter * n, h->stream);
    return launch_mean_field_persistent(lats, va, vb, a, h->unary.p, h->Q.p, h->geom.Ntot, h->L, h->Lp, n_iter,
                                        h->counters.p, h->stream);
}

void run_inference(dcrf_handle *h, int n_iter) {
    DCRF_REQUIRE(n_iter >= 0, DCRF_EINVAL, "n_iter must be >= 0");
    DCRF_REQUIRE(h->L >= 1, DCRF_ESTATE, "model has no labels");
    join_upload(h);
    finish_builds(h);
    if (try_persistent(h, n_iter)) {
        h->q_valid = true;
        return;
    }
    start_inference(h);
    // (One iteration captured as a CUDA graph and replayed n times was measured for the small
    // configurations: one VOC image 1.77 vs 1.76 ms per step, 32 SEC maps 1.18 vs 1.05 -- instantiating
    // ~20 nodes per handle costs what the shorter launch gaps save; profiles/r2_small_problem_graphs.txt.)
    for (int it = 0; it < n_iter; it++) step_inference(h);
}

Pairwise &get_pw(dcrf_handle *h, int k) {
    DCRF_REQUIRE(k >= 0 && k < (int)h->pw.size(), DCRF_EINVAL, "pairwise index out of range");
    finish_pairwise(h, *h->pw[k]);
    return *h->pw[k];
}

}  // namespace

extern "C" {

const char *dcrf_last_error(void) { return t_error.c_str(); }
const char *dcrf_version(void) { return "dcrf_b200 0.1 sm_100a"; }
int64_t dcrf_launch_count(void) { return g_launches.load(); }
void dcrf_copy_count(int64_t *h2d_bytes, int64_t *d2h_bytes) {
    if (h2d_bytes) *h2d_bytes = g_h2d_bytes.load();
    if (d2h_bytes) *d2h_bytes = g_d2h_bytes.load();
}

int dcrf_create(int w, int h, int n_labels, int device, void *stream, dcrf_t **out) {
    return guarded([&] { create_common(1, &w, &h, true, n_labels, device, stream, out); });
}

int dcrf_create_nd(int n_vars, int n_labels, int device, void *stream, dcrf_t **out) {
    return guarded([&] {
        int one = 1;
        create_common(1, &n_vars, &one, false, n_labels, device, stream, out);
    });
}

int dcrf_create_batch(int n_images, const int *w, const int *h, int n_labels, int device, void *stream,
                      dcrf_t **out) {
    return guarded([&] {
        DCRF_REQUIRE(w && h, DCRF_EINVAL, "w/h arrays are NULL");
        create_common(n_images, w, h, true, n_labels, device, stream, out);
    });
}

void dcrf_destroy(dcrf_t *h) {
    if (!h) return;
    try {
        DeviceGuard guard(h->device);
        const double t0 = trace_now();
        join_upload(h);
        if (h->ev_upload_begin) {
            cudaEventDestroy(h->ev_upload_begin);
            cudaEventDestroy(h->ev_upload_end);
        }
        // lattices built on a side stream free their memory in THAT stream's order: it must come after
        // everything the handle's stream still does with them
        {
            cudaEvent_t ev = nullptr;
            for (auto &p : h->pw) {
                if (!p->built_on) continue;
                if (!ev) {
                    if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) break;
                    cudaEventRecord(ev, h->stream);
                }
                cudaStreamWaitEvent(p->built_on, ev, 0);
            }
            if (ev) cudaEventDestroy(ev);
        }
        h->pw.clear();
        h->unary.release();
        h->Q.release();
        h->counters.release();
        h->d_w.release();
        h->d_h.release();
        h->d_pix_start.release();
        if (h->ev_fork) {
            cudaStreamSynchronize(h->stream);
            cudaEventDestroy(h->ev_fork);
            for (int k = 0; k < kMaxPairwise - 1; k++) cudaEventDestroy(h->ev_join[k]);
        }
        if (t0 != 0.0) {  // DCRF_TRACE: state of the handle's pool after everything was freed
            uint64_t reserved = 0, used = 0, high = 0;
            cudaMemPool_t pool = stream_pool(h->stream);
            cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved);
            cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used);
            cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemHigh, &high);
            fprintf(stderr, "[dcrf trace] pool of stream %p: reserved %.1f MB, in use %.1f MB, high-water %.1f MB\n",
                    (void *)h->stream, reserved / 1e6, used / 1e6, high / 1e6);
        }
        h->streams.reset();  // the last holder destroys the set's streams and their pools
        trace_slow("dcrf_destroy", t0, 0);
    } catch (...) {
    }
    delete h;
}

int dcrf_mem_info(int device, int64_t *free_bytes, int64_t *total_bytes) {
    return guarded([&] {
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
            throw Error{DCRF_ECUDA, "no CUDA device: dcrf_b200 has no CPU fallback"};
        if (device < 0) DCRF_CUDA(cudaGetDevice(&device));
        DCRF_REQUIRE(device < ndev, DCRF_EINVAL, "device index out of range");
        DeviceGuard guard(device);
        size_t f = 0, t = 0;
        DCRF_CUDA(cudaMemGetInfo(&f, &t));
        // memory cached by the library's own pools is reusable by the next handle: count it as free
        {
            std::lock_guard<std::mutex> lock(g_pool_mu);
            for (auto &kv : g_pools) {
                if (kv.first.first != device) continue;
                uint64_t reserved = 0, used = 0;
                cudaMemPoolGetAttribute(kv.second, cudaMemPoolAttrReservedMemCurrent, &reserved);
                cudaMemPoolGetAttribute(kv.second, cudaMemPoolAttrUsedMemCurrent, &used);
                if (reserved > used) f += (size_t)(reserved - used);
            }
        }
        if (free_bytes) *free_bytes = (int64_t)f;
        if (total_bytes) *total_bytes = (int64_t)t;
    });
}

int dcrf_trim_memory(void) {
    return guarded([&] {
        DCRF_CUDA(cudaDeviceSynchronize());
        trim_all_pools();
    });
}

int dcrf_stream_create(int device, void **stream_out) {
    return guarded([&] {
        DCRF_REQUIRE(stream_out, DCRF_EINVAL, "NULL argument");
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
            throw Error{DCRF_ECUDA, "no CUDA device: dcrf_b200 has no CPU fallback"};
        if (device < 0) DCRF_CUDA(cudaGetDevice(&device));
        DCRF_REQUIRE(device < ndev, DCRF_EINVAL, "device index out of range");
        DeviceGuard guard(device);
        auto set = new_owned_set(device);
        {
            std::lock_guard<std::mutex> lock(g_sets_mu);
            g_sets[std::make_pair(device, set->primary)] = set;
        }
        *stream_out = (void *)set->primary;
    });
}

int dcrf_stream_destroy(void *stream) {
    return guarded([&] {
        DCRF_REQUIRE(stream, DCRF_EINVAL, "NULL stream");
        // drop the registry's reference: the stream, its side streams and their pools are destroyed now,
        // or with the last handle still running on it
        std::shared_ptr<StreamSet> set;
        {
            std::lock_guard<std::mutex> lock(g_sets_mu);
            for (auto it = g_sets.begin(); it != g_sets.end(); ++it)
                if (it->first.second == (cudaStream_t)stream && it->second->owns_primary) {
                    set = it->second;
                    g_sets.erase(it);
                    break;
                }
        }
        DCRF_REQUIRE(set != nullptr, DCRF_EINVAL, "not a stream created by dcrf_stream_create");
    });
}

int dcrf_set_option(dcrf_t *h, int option, int value) {
    return guarded([&] {
        DCRF_REQUIRE(h, DCRF_EINVAL, "NULL handle");
        DCRF_REQUIRE(option == DCRF_OPT_EXACT_ARITHMETIC || option == DCRF_OPT_ASYNC_HOST ||
                         option == DCRF_OPT_PERSISTENT, DCRF_EINVAL, "unknown option");
        if (option == DCRF_OPT_PERSISTENT) {
            DCRF_REQUIRE(value >= -1 && value <= 1, DCRF_EINVAL, "DCRF_OPT_PERSISTENT takes -1, 0 or 1");
            h->persistent = value;
            return;
        }
        if (option == DCRF_OPT_EXACT_ARITHMETIC) {
            DCRF_REQUIRE(value >= 0 && value <= 3, DCRF_EINVAL, "arithmetic mode must be 0, 1, 2 or 3");
            if (value == DCRF_ARITH_AUTO) {  // re-derive from the terms added so far
                h->arith_auto = true;
                h->arith = kArithFma;
                for (auto &p : h->pw)
                    if (p->stiff) h->arith = kArithRef;
            } else {
                set_arith(h, value);
            }
        } else {
            h->async_host = value != 0;
        }
    });
}

int dcrf_get_arithmetic(dcrf_t *h, int *mode_out) {
    return guarded([&] {
        DCRF_REQUIRE(h && mode_out, DCRF_EINVAL, "NULL argument");
        *mode_out = h->arith;
    });
}

int dcrf_synchronize(dcrf_t *h) {
    return guarded([&] {
        DCRF_REQUIRE(h, DCRF_EINVAL, "NULL handle");
        join_upload(h);
        DCRF_CUDA(cudaStreamSynchronize(h->stream));
    });
}

int dcrf_set_unary(dcrf_t *h, const float *U, int on_device) {
    return guarded([&] {
        DCRF_REQUIRE(h && U, DCRF_EINVAL, "NULL argument");
        DCRF_REQUIRE(h->L >= 1, DCRF_ESTATE, "model has no labels");
        DeviceGuard guard(h->device);
        join_upload(h);
        if (!on_device && h->async_host) {
            // H2D + layout change on the upload stream; `stream` goes on (lattice builds) meanwhile
            const cudaStream_t up = h->streams->get_upload();
            if (!h->ev_upload_begin) {
                DCRF_CUDA(cudaEventCreateWithFlags(&h->ev_upload_begin, cudaEventDisableTiming));
                DCRF_CUDA(cudaEventCreateWithFlags(&h->ev_upload_end, cudaEventDisableTiming));
            }
            const size_t n = (size_t)total_ln(h);
            h->upload_stage.alloc(n, h->stream);
            DCRF_CUDA(cudaEventRecord(h->ev_upload_begin, h->stream));
            DCRF_CUDA(cudaStreamWaitEvent(up, h->ev_upload_begin, 0));
            DCRF_CUDA(copy_h2d(h->upload_stage.p, U, sizeof(float) * n, up));
            launch_ln_to_pm(h->upload_stage.p, h->unary.p, h->geom, h->L, h->Lp, up);
            DCRF_CUDA(cudaEventRecord(h->ev_upload_end, up));
            h->upload_pending = true;
        } else {
            DevBuf<float> stage;
            const float *src = to_device(h, U, (size_t)total_ln(h), on_device, stage);
            launch_ln_to_pm(src, h->unary.p, h->geom, h->L, h->Lp, h->stream);
            if (!on_device) host_sync(h);  // caller may reuse U
        }
        h->unary_set = true;
        h->q_valid = false;
    });
}

int dcrf_set_unary_from_probs(dcrf_t *h, const void *probs, int is_f64, double scale, double clip,
                              int has_clip, int on_device) {
    return guarded([&] {
        DCRF_REQUIRE(h && probs, DCRF_EINVAL, "NULL argument");
        DCRF_REQUIRE(h->L >= 1, DCRF_ESTATE, "model has no labels");
        DCRF_REQUIRE(scale > 0.0 && scale <= 1.0, DCRF_EINVAL, "`scale` needs to be in (0,1]");
        DeviceGuard guard(h->device);
        join_upload(h);
        const size_t n = (size_t)total_ln(h);
        const void *src = probs;
        DevBuf<double> stage64;
        DevBuf<float> stage32;
        if (!on_device) {
            if (is_f64) src = to_device(h, (const double *)probs, n, 0, stage64);
            else src = to_device(h, (const float *)probs, n, 0, stage32);
        }
        launch_unary_from_probs(src, is_f64, h->unary.p, h->geom, h->L, h->Lp, scale, clip, has_clip, h->stream);
        if (!on_device) host_sync(h);
        h->unary_set = true;
        h->q_valid = false;
    });
}

int dcrf_set_unary_from_logits(dcrf_t *h, const float *feat, int use_log, int on_device) {
    return guarded([&] {
        DCRF_REQUIRE(h && feat, DCRF_EINVAL, "NULL argument");
        DCRF_REQUIRE(h->L >= 1, DCRF_ESTATE, "model has no labels");
        // lib/crf.py is missing from the reference tree and every call site uses use_log = True
        // (SEC.py:275, DSRG.py:328, model.py:689,693): the other branch is not guessed
        DCRF_REQUIRE(use_log != 0, DCRF_EINVAL, "use_log = 0 is not exercised by the reference and not implemented");
        DeviceGuard guard(h->device);
        join_upload(h);
        DevBuf<float> stage;
        const float *src = to_device(h, feat, (size_t)total_ln(h), on_device, stage);
        launch_unary_from_logits(src, h->unary.p, h->geom.Ntot, h->L, h->Lp, use_log, h->stream);
        if (!on_device) host_sync(h);
        h->unary_set = true;
        h->q_valid = false;
    });
}

int dcrf_set_unary_from_labels(dcrf_t *h, const int32_t *labels, float gt_prob, int zero_unsure, int on_device) {
    return guarded([&] {
        DCRF_REQUIRE(h && labels, DCRF_EINVAL, "NULL argument");
        DCRF_REQUIRE(h->L >= 2, DCRF_ESTATE, "unary_from_labels needs at least 2 labels");
        DCRF_REQUIRE(gt_prob > 0.f && gt_prob < 1.f, DCRF_EINVAL, "`gt_prob must be in (0,1).");
        DeviceGuard guard(h->device);
        join_upload(h);
        DevBuf<int32_t> stage;
        DevBuf<int> bad;
        const int32_t *src = to_device(h, labels, (size_t)h->geom.Ntot, on_device, stage);
        bad.alloc(1, h->stream);
        DCRF_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), h->stream));
        // energies in double like NumPy, stored as float32
        const float n_energy = (float)(-log((1.0 - (double)gt_prob) / (double)(h->L - 1)));
        const float p_energy = (float)(-log((double)gt_prob));
        const float unsure = (float)(-log(1.0 / (double)h->L));
        launch_unary_from_labels(src, h->unary.p, h->geom.Ntot, h->L, h->Lp, n_energy, p_energy, unsure,
                                 zero_unsure, bad.p, h->stream);
        int h_bad = 0;
        DCRF_CUDA(copy_d2h(&h_bad, bad.p, sizeof(int), h->stream));
        DCRF_CUDA(cudaStreamSynchronize(h->stream));
        DCRF_REQUIRE(h_bad == 0, DCRF_EINVAL, "label out of range in unary_from_labels");
        h->unary_set = true;
        h->q_valid = false;
    });
}

int dcrf_add_pairwise_gaussian(dcrf_t *h, float sx, float sy, int compat_kind, const float *compat,
                               int kernel_type, int normalization_type) {
    return guarded([&] {
        DCRF_REQUIRE(h, DCRF_EINVAL, "NULL handle");
        DCRF_REQUIRE(h->has_geom, DCRF_ESTATE, "addPairwiseGaussian needs a 2-D model");
        DeviceGuard guard(h->device);
        ProfGuard pguard(h);
        FeatureSpec fs;
        memset(&fs, 0, sizeof(fs));
        fs.mode = 0;
        fs.d = 2;
        fs.s[0] = sx;
        fs.s[1] = sy;
        add_pairwise(h, fs, compat_kind, compat, kernel_type, normalization_type);
    });
}

int dcrf_add_pairwise_bilateral(dcrf_t *h, float sx, float sy, float sr, float sg, float sb,
                                const uint8_t *rgb, int on_device, int compat_kind, const float *compat,
                                int kernel_type, int normalization_type) {
    return guarded([&] {
        DCRF_REQUIRE(h && rgb, DCRF_EINVAL, "NULL argument");
        DCRF_REQUIRE(h->has_geom, DCRF_ESTATE, "addPairwiseBilateral needs a 2-D model");
        DeviceGuard guard(h->device);
        ProfGuard pguard(h);
        DevBuf<uint8_t> stage;
        FeatureSpec fs;
        memset(&fs, 0, sizeof(fs));
        fs.mode = 1;
        fs.d = 5;
        fs.s[0] = sx; fs.s[1] = sy; fs.s[2] = sr; fs.s[3] = sg; fs.s[4] = sb;
        fs.rgb = to_device(h, rgb, (size_t)h->geom.Ntot * 3, on_device, stage);
        add_pairwise(h, fs, compat_kind, compat, kernel_type, normalization_type, &stage, nullptr, on_device != 0);
        if (!on_device) host_sync(h);
    });
}

int dcrf_add_pairwise_energy(dcrf_t *h, const float *features, int d, int on_device, int compat_kind,
                             const float *compat, int kernel_type, int normalization_type) {
    return guarded([&] {
        DCRF_REQUIRE(h && features, DCRF_EINVAL, "NULL argument");
        DCRF_REQUIRE(h->geom.B == 1, DCRF_ESTATE, "addPairwiseEnergy is single-image only");
        DCRF_REQUIRE(d >= 1 && d <= kMaxD, DCRF_EINVAL, "feature dimension d must be in [1, 7]");
        DeviceGuard guard(h->device);
        ProfGuard pguard(h);
        DevBuf<float> stage;
        FeatureSpec fs;
        memset(&fs, 0, sizeof(fs));
        fs.mode = 2;
        fs.d = d;
        fs.features = to_device(h, features, (size_t)h->geom.Ntot * d, on_device, stage);
        add_pairwise(h, fs, compat_kind, compat, kernel_type, normalization_type, nullptr, &stage, on_device != 0);
        if (!on_device) host_sync(h);
    });
}

int dcrf_inference(dcrf_t *h, int n_iter, float *Q_out, int on_device) {
    return guarded([&] {
        DCRF_REQUIRE(h, DCRF_EINVAL, "NULL handle");
        DeviceGuard guard(h->device);
        ProfGuard pguard(h);
        run_inference(h, n_iter);
        emit_q(h, Q_out, on_device);
    });
}

int dcrf_map(dcrf_t *h, int n_iter, int32_t *labels_out, int on_device) {
    return guarded([&] {
        DCRF_REQUIRE(h && labels_out, DCRF_EINVAL, "NULL argument");
        DeviceGuard guard(h->device);
        ProfGuard pguard(h);
        run_inference(h, n_iter);
        emit_labels(h, labels_out, on_device);
    });
}

int dcrf_map_u8(dcrf_t *h, int n_iter, uint8_t *labels_out, int on_device) {
    return guarded([&] {
        DCRF_REQUIRE(h && labels_out, DCRF_EINVAL, "NULL argument");
        DCRF_REQUIRE(h->L <= 256, DCRF_EINVAL, "uint8 labels need n_labels <= 256");
        DeviceGuard guard(h->device);
        ProfGuard pguard(h);
        run_inference(h, n_iter);
        emit_labels(h, labels_out, on_device);
    });
}

int dcrf_run(dcrf_t *h, int n_iter) {
    return guarded([&] {
        DCRF_REQUIRE(h, DCRF_EINVAL, "NULL handle");
        DeviceGuard guard(h->device);
        ProfGuard pguard(h);
        run_inference(h, n_iter);
        host_sync(h);
    });
}

int dcrf_get_labels(dcrf_t *h, int32_t *labels_out, int on_device) {
    return guarded([&] {
        DCRF_REQUIRE(h && labels_out, DCRF_EINVAL, "NULL argument");
        DCRF_REQUIRE(h->q_valid, DCRF_ESTATE, "no running Q: call dcrf_run / startInference first");
        DeviceGuard guard(h->device);
        emit_labels(h, labels_out, on_device);
    });
}

int dcrf_get_labels_u8(dcrf_t *h, uint8_t *labels_out, int on_device) {
    return guarded([&] {
        DCRF_REQUIRE(h && labels_out, DCRF_EINVAL, "NULL argument");
        DCRF_REQUIRE(h->q_valid, DCRF_ESTATE, "no running Q: call dcrf_run / startInference first");
        DCRF_REQUIRE(h->L <= 256, DCRF_EINVAL, "uint8 labels need n_labels <= 256");
        DeviceGuard guard(h->device);
        emit_labels(h, labels_out, on_device);
    });
}

int dcrf_start_inference(dcrf_t *h) {
    return guarded([&] {
        DCRF_REQUIRE(h, DCRF_EINVAL, "NULL handle");
        DeviceGuard guard(h->device);
        start_inference(h);
    });
}

int dcrf_step_inference(dcrf_t *h) {
    return guarded([&] {
        DCRF_REQUIRE(h, DCRF_EINVAL, "NULL handle");
        DeviceGuard guard(h->device);
        ProfGuard pguard(h);
        step_inference(h);
    });
}

int dcrf_get_q(dcrf_t *h, float *Q_out, int on_device) {
    return guarded([&] {
        DCRF_REQUIRE(h, DCRF_EINVAL, "NULL handle");
        DCRF_REQUIRE(h->q_valid, DCRF_ESTATE, "no running Q: call startInference first");
        DeviceGuard guard(h->device);
        emit_q(h, Q_out, on_device);
    });
}

int dcrf_get_q_hwc(dcrf_t *h, float min_prob, int take_log, float *out, int on_device) {
    return guarded([&] {
        DCRF_REQUIRE(h && out, DCRF_EINVAL, "NULL argument");
        DCRF_REQUIRE(h->q_valid, DCRF_ESTATE, "no running Q: call startInference first");
        DCRF_REQUIRE(!(min_prob > 0.f) || min_prob < 1.f, DCRF_EINVAL, "min_prob must be below 1");
        DeviceGuard guard(h->device);
        const int64_t n = total_ln(h);
        const int renorm = min_prob > 0.f ? 1 : 0;
        if (on_device) {
            launch_q_to_hwc(h->Q.p, out, h->geom.Ntot, h->L, h->Lp, min_prob, renorm, take_log, h->stream);
        } else {
            DevBuf<float> stage;
            stage.alloc(n, h->stream);
            launch_q_to_hwc(h->Q.p, stage.p, h->geom.Ntot, h->L, h->Lp, min_prob, renorm, take_log, h->stream);
            DCRF_CUDA(copy_d2h(out, stage.p, sizeof(float) * n, h->stream));
            host_sync(h);
        }
    });
}

int dcrf_set_q(dcrf_t *h, const float *Q_in, int on_device) {
    return guarded([&] {
        DCRF_REQUIRE(h && Q_in, DCRF_EINVAL, "NULL argument");
        DCRF_REQUIRE(h->L >= 1, DCRF_ESTATE, "model has no labels");
        DeviceGuard guard(h->device);
        DevBuf<float> stage;
        const float *src = to_device(h, Q_in, (size_t)total_ln(h), on_device, stage);
        launch_ln_to_pm(src, h->Q.p, h->geom, h->L, h->Lp, h->stream);
        if (!on_device) host_sync(h);
        h->q_valid = true;
    });
}

int dcrf_kl_divergence(dcrf_t *h, double *kl_out) {
    return guarded([&] {
        DCRF_REQUIRE(h && kl_out, DCRF_EINVAL, "NULL argument");
        DCRF_REQUIRE(h->q_valid, DCRF_ESTATE, "no running Q: call startInference first");
        DeviceGuard guard(h->device);
        join_upload(h);
        finish_builds(h);
        cudaStream_t s = h->stream;
        const int64_t Ntot = h->geom.Ntot;
        const bool seq = h->L <= 2;
        std::vector<DevBuf<float>> outs(h->pw.size());
        const float *ptrs[kMaxPairwise] = {nullptr, nullptr, nullptr, nullptr};
        for (size_t k = 0; k < h->pw.size(); k++) {
            Pairwise &p = *h->pw[k];
            const float *blurred =
                filter_to_lattice(h, p, h->Q.p, h->Lp, pre_norm(p.ntype), seq, p.valA.p, p.valB.p);
            outs[k].alloc((size_t)Ntot * h->Lp, s);
            launch_slice_pairwise_only(make_term(p, blurred), outs[k].p, Ntot, h->L, h->Lp, s);
            ptrs[k] = outs[k].p;
        }
        DevBuf<double> d_out;
        d_out.alloc(1, s);
        launch_kl(h->Q.p, h->unary.p, ptrs, (int)h->pw.size(), Ntot, h->L, h->Lp, d_out.p, s);
        DCRF_CUDA(copy_d2h(kl_out, d_out.p, sizeof(double), s));
        DCRF_CUDA(cudaStreamSynchronize(s));
    });
}

int dcrf_profile_enable(dcrf_t *h, int enable) {
    return guarded([&] {
        DCRF_REQUIRE(h, DCRF_EINVAL, "NULL handle");
        h->prof.on = enable != 0;
    });
}

int dcrf_profile_read(dcrf_t *h, int kernel_class, int tag, double *total_ms, int64_t *launches, int reset) {
    return guarded([&] {
        DCRF_REQUIRE(h, DCRF_EINVAL, "NULL handle");
        DeviceGuard guard(h->device);
        DCRF_CUDA(cudaStreamSynchronize(h->stream));
        double ms = 0;
        int64_t n = 0;
        for (auto &r : h->prof.recs) {
            if (r.cls != kernel_class || (tag >= 0 && r.tag != tag)) continue;
            float t = 0;
            DCRF_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
            ms += t;
            n++;
        }
        if (total_ms) *total_ms = ms;
        if (launches) *launches = n;
        if (reset) {
            for (auto &r : h->prof.recs) { h->prof.pool.push_back(r.a); h->prof.pool.push_back(r.b); }
            h->prof.recs.clear();
        }
    });
}

int dcrf_num_pairwise(dcrf_t *h, int *n_out) {
    return guarded([&] {
        DCRF_REQUIRE(h && n_out, DCRF_EINVAL, "NULL argument");
        *n_out = (int)h->pw.size();
    });
}

int dcrf_lattice_info(dcrf_t *h, int kernel, int *d_out, int64_t *M_out, int64_t *M_per_image) {
    return guarded([&] {
        DCRF_REQUIRE(h, DCRF_EINVAL, "NULL handle");
        Pairwise &p = get_pw(h, kernel);
        if (d_out) *d_out = p.lat.d;
        if (M_out) *M_out = p.lat.M;
        if (M_per_image)
            for (int b = 0; b < h->geom.B; b++) M_per_image[b] = p.lat.vert_start[b + 1] - p.lat.vert_start[b];
    });
}

int dcrf_lattice_export(dcrf_t *h, int kernel, int image, int16_t *keys, int32_t *offsets, float *bary,
                        int32_t *neighbours, float *norm) {
    return guarded([&] {
        DCRF_REQUIRE(h, DCRF_EINVAL, "NULL handle");
        DCRF_REQUIRE(image >= 0 && image < h->geom.B, DCRF_EINVAL, "image index out of range");
        DeviceGuard guard(h->device);
        Pairwise &p = get_pw(h, kernel);
        cudaStream_t s = h->stream;
        const int d = p.lat.d, d1 = d + 1;
        const int64_t p0 = h->geom.pix_start[image], Nb = h->geom.pix_start[image + 1] - p0;
        const int64_t v0 = p.lat.vert_start[image], Mb = p.lat.vert_start[image + 1] - v0;
        std::vector<int16_t> k8;
        std::vector<int2> nb;
        if (keys) {
            k8.resize((size_t)Mb * 8);
            DCRF_CUDA(copy_d2h(k8.data(), p.lat.vkeys.p + v0 * 8, sizeof(int16_t) * Mb * 8, s));
        }
        if (offsets)
            DCRF_CUDA(copy_d2h(offsets, p.lat.offset.p + p0 * d1, sizeof(int32_t) * Nb * d1, s));
        if (bary)
            DCRF_CUDA(copy_d2h(bary, p.lat.bary.p + p0 * d1, sizeof(float) * Nb * d1, s));
        if (neighbours) {
            nb.resize((size_t)Mb * d1);
            for (int j = 0; j < d1; j++)
                DCRF_CUDA(copy_d2h(nb.data() + (size_t)j * Mb, p.lat.neigh.p + (int64_t)j * p.lat.M + v0,
                                          sizeof(int2) * Mb, s));
        }
        if (norm) {
            DCRF_REQUIRE(p.norm.p != nullptr, DCRF_ESTATE, "kernel has no norm (NO_NORMALIZATION)");
            DCRF_CUDA(copy_d2h(norm, p.norm.p + p0, sizeof(float) * Nb, s));
        }
        DCRF_CUDA(cudaStreamSynchronize(s));
        if (keys)
            for (int64_t v = 0; v < Mb; v++)
                for (int i = 0; i < d; i++) keys[v * d + i] = k8[(size_t)v * 8 + i];
        if (offsets)
            for (int64_t i = 0; i < Nb * d1; i++) offsets[i] -= (int32_t)v0;
        if (neighbours)
            for (int64_t i = 0; i < Mb * d1; i++) {
                neighbours[2 * i + 0] = nb[i].x >= 0 ? nb[i].x - (int32_t)v0 : -1;
                neighbours[2 * i + 1] = nb[i].y >= 0 ? nb[i].y - (int32_t)v0 : -1;
            }
    });
}

int dcrf_expf_ref(const float *x, float *y, int64_t n, int device) {
    return guarded([&] {
        DCRF_REQUIRE((x && y) || n == 0, DCRF_EINVAL, "NULL argument");
        DCRF_REQUIRE(n >= 0, DCRF_EINVAL, "n must be >= 0");
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
            throw Error{DCRF_ECUDA, "no CUDA device: dcrf_b200 has no CPU fallback"};
        if (device < 0) DCRF_CUDA(cudaGetDevice(&device));
        DCRF_REQUIRE(device < ndev, DCRF_EINVAL, "device index out of range");
        DeviceGuard guard(device);
        const std::shared_ptr<StreamSet> set = thread_set(device);
        cudaStream_t s = set->primary;
        DevBuf<float> dx, dy;
        dx.alloc((size_t)n, s);
        dy.alloc((size_t)n, s);
        DCRF_CUDA(copy_h2d(dx.p, x, sizeof(float) * n, s));
        launch_expf_ref(dx.p, dy.p, n, s);
        DCRF_CUDA(copy_d2h(y, dy.p, sizeof(float) * n, s));
        DCRF_CUDA(cudaStreamSynchronize(s));
    });
}

int dcrf_lattice_filter(dcrf_t *h, int kernel, const float *in, float *out, int value_size) {
    return guarded([&] {
        DCRF_REQUIRE(h && in && out, DCRF_EINVAL, "NULL argument");
        DCRF_REQUIRE(h->geom.B == 1, DCRF_ESTATE, "lattice_filter is single-image only");
        DCRF_REQUIRE(value_size >= 1 && value_size <= 128, DCRF_EINVAL, "value_size must be in [1, 128]");
        DeviceGuard guard(h->device);
        Pairwise &p = get_pw(h, kernel);
        cudaStream_t s = h->stream;
        const int64_t N = h->geom.Ntot;
        const int vs = value_size, vp = ((vs + 3) / 4) * 4;
        const bool seq = vs <= 2;
        DevBuf<float> ln, pm, a, b, sl;
        ln.alloc((size_t)N * vs, s);
        pm.alloc((size_t)N * vp, s);
        a.alloc((size_t)p.lat.M * vp, s);
        b.alloc((size_t)p.lat.M * vp, s);
        sl.alloc((size_t)N * vp, s);
        DCRF_CUDA(copy_h2d(ln.p, in, sizeof(float) * N * vs, s));
        launch_ln_to_pm(ln.p, pm.p, h->geom, vs, vp, s);
        const float *blurred = filter_to_lattice(h, p, pm.p, vp, false, seq, a.p, b.p);
        launch_slice_plain(p.lat, blurred, sl.p, N, vp, seq, s);
        launch_pm_to_ln(sl.p, ln.p, h->geom, vs, vp, s);
        DCRF_CUDA(copy_d2h(out, ln.p, sizeof(float) * N * vs, s));
        DCRF_CUDA(cudaStreamSynchronize(s));
    });
}

}  // extern "C"
