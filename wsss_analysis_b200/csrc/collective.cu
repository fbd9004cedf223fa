// collective.cu -- the one collective of the path at the C level: the int64 SUM all-reduce of the
// (C+1, C) confusion matrix behind mIoU (replaces the single-process
// accumulation of /root/reference/03b_irn/step/eval_sem_seg.py:41-50 when the image list is striped
// over the GPUs of a box).  NCCL over NVLink 5 / NVSwitch; integer addition is order independent, so
// the result is bit-identical to a single-process sum for any rank count.
//
// NCCL is bound at run time (dlopen of the libnccl.so.2 already loaded by PyTorch, or the one named by
// DCRF_NCCL_LIB): libdcrf_b200.so has no link-time dependency on it and single-GPU users never load it.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "common.cuh"

namespace dcrf {
namespace {

typedef struct { char internal[128]; } NcclUniqueId;  // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
typedef void *NcclComm;
enum { kNcclInt64 = 4, kNcclSum = 0 };  // ncclDataType_t / ncclRedOp_t values of nccl.h

struct NcclApi {
    int (*GetUniqueId)(NcclUniqueId *) = nullptr;
    int (*CommInitRank)(NcclComm *, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    int (*CommCount)(NcclComm, int *) = nullptr;
};

const NcclApi &nccl() {
    static NcclApi api;
    static std::once_flag once;
    static std::string err;
    std::call_once(once, [] {
        const char *names[] = {getenv("DCRF_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        void *lib = nullptr;
        for (const char *n : names) {
            if (!n || !*n) continue;
            lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) {
            err = "NCCL not found (import torch first, or set DCRF_NCCL_LIB to libnccl.so.2)";
            return;
        }
        api.GetUniqueId = (int (*)(NcclUniqueId *))dlsym(lib, "ncclGetUniqueId");
        api.CommInitRank = (int (*)(NcclComm *, int, NcclUniqueId, int))dlsym(lib, "ncclCommInitRank");
        api.CommDestroy = (int (*)(NcclComm))dlsym(lib, "ncclCommDestroy");
        api.AllReduce = (int (*)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t))dlsym(lib, "ncclAllReduce");
        api.GetErrorString = (const char *(*)(int))dlsym(lib, "ncclGetErrorString");
        api.CommCount = (int (*)(NcclComm, int *))dlsym(lib, "ncclCommCount");
        if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce) err = "NCCL symbols missing";
    });
    if (!err.empty()) throw Error{DCRF_ECUDA, err};
    return api;
}

void check_nccl(int rc, const char *what) {
    if (rc == 0) return;
    const NcclApi &a = nccl();
    throw Error{DCRF_ECUDA, std::string(what) + ": " + (a.GetErrorString ? a.GetErrorString(rc) : "NCCL error")};
}

template <typename F>
int guarded(F &&f) {
    try {
        f();
        return DCRF_OK;
    } catch (const Error &e) {
        set_error(e.msg);
        return e.code;
    } catch (const std::exception &e) {
        set_error(e.what());
        return DCRF_EINVAL;
    }
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

extern "C" {

int dcrf_nccl_unique_id(void *id_out) {
    return guarded([&] {
        DCRF_REQUIRE(id_out, DCRF_EINVAL, "NULL argument");
        NcclUniqueId id;
        check_nccl(nccl().GetUniqueId(&id), "ncclGetUniqueId");
        memcpy(id_out, &id, sizeof(id));
    });
}

int dcrf_nccl_comm_create(int n_ranks, int rank, const void *unique_id, int device, void **comm_out) {
    return guarded([&] {
        DCRF_REQUIRE(unique_id && comm_out, DCRF_EINVAL, "NULL argument");
        DCRF_REQUIRE(n_ranks >= 1 && rank >= 0 && rank < n_ranks, DCRF_EINVAL, "bad rank / n_ranks");
        DevGuard guard(device);
        NcclUniqueId id;
        memcpy(&id, unique_id, sizeof(id));
        NcclComm comm = nullptr;
        check_nccl(nccl().CommInitRank(&comm, n_ranks, id, rank), "ncclCommInitRank");
        *comm_out = comm;
    });
}

int dcrf_nccl_comm_destroy(void *comm) {
    return guarded([&] {
        DCRF_REQUIRE(comm, DCRF_EINVAL, "NULL communicator");
        check_nccl(nccl().CommDestroy((NcclComm)comm), "ncclCommDestroy");
    });
}

int dcrf_confusion_allreduce(void *comm, int64_t *conf, int64_t count, int device, void *stream) {
    return guarded([&] {
        DCRF_REQUIRE(comm && conf, DCRF_EINVAL, "NULL argument");
        DCRF_REQUIRE(count >= 0, DCRF_EINVAL, "count must be >= 0");
        if (count == 0) return;
        DevGuard guard(device);
        check_nccl(nccl().AllReduce(conf, conf, (size_t)count, kNcclInt64, kNcclSum, (NcclComm)comm, (cudaStream_t)stream),
                   "ncclAllReduce");
    });
}

}  // extern "C"
