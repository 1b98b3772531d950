// comm_api.cu -- the one exchange of the sharded path behind the C ABI: an NCCL all-gather of per-block
// packed sizes over NVLink / NVSwitch, after which every rank runs the same offset scan (go/block_index.go:16-35
// applied to the whole file) and knows where its bytes go.  Payload never crosses the interconnect.
//
// NCCL is loaded at run time (dlopen "libnccl.so.2"): a process that already holds an NCCL (torch's bundled one)
// gets that very library, a Go host gets the system one; libminnow_b200 itself links against neither.  Only stable C
// entry points are used (ncclGetUniqueId, ncclCommInitRank, ncclAllGather, ncclCommDestroy, ncclGetErrorString).
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "ctx.cuh"

namespace {

struct NcclUniqueId { char internal[128]; };              // ncclUniqueId
typedef struct ncclComm *ncclComm_t;
typedef int ncclResult_t;                                  // 0 = ncclSuccess
constexpr int NCCL_INT64 = 4;                              // ncclInt64

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(NcclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, NcclUniqueId, int) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    const char *why = "";
};

NcclApi &nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {getenv("MNW_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            if (!n) continue;
            api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) { api.why = "libnccl.so.2 not found (set MNW_NCCL_LIB)"; return; }
        api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.lib, "ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.lib, "ncclCommInitRank");
        api.AllGather = (decltype(api.AllGather))dlsym(api.lib, "ncclAllGather");
        api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
        api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.lib, "ncclGetErrorString");
        if (!api.GetUniqueId || !api.CommInitRank || !api.AllGather || !api.CommDestroy) {
            api.why = "libnccl lacks an expected entry point";
            api.lib = nullptr;
        }
    });
    return api;
}

int nccl_fail(mnw_ctx *ctx, const char *what, ncclResult_t r) {
    NcclApi &a = nccl();
    return mnw_fail(ctx, MNW_ERR_CUDA, "%s: %s", what, a.GetErrorString ? a.GetErrorString(r) : "NCCL error");
}

}  // namespace

extern "C" {

int mnw_comm_unique_id(mnw_comm_id *out) {
    static_assert(sizeof(mnw_comm_id) == sizeof(NcclUniqueId), "mnw_comm_id is an ncclUniqueId");
    if (!out) return MNW_ERR_ARG;
    NcclApi &a = nccl();
    if (!a.lib) return mnw_fail(nullptr, MNW_ERR_CUDA, "mnw_comm_unique_id: %s", a.why);
    NcclUniqueId id;
    const ncclResult_t r = a.GetUniqueId(&id);
    if (r != 0) return nccl_fail(nullptr, "ncclGetUniqueId", r);
    memcpy(out, &id, sizeof id);
    return MNW_OK;
}

int mnw_comm_init(mnw_ctx *ctx, const mnw_comm_id *id, int nranks, int rank) {
    if (!ctx || !id || nranks < 1 || rank < 0 || rank >= nranks) return mnw_fail(ctx, MNW_ERR_ARG, "mnw_comm_init: bad argument");
    (void)cudaSetDevice(ctx->device);
    NcclApi &a = nccl();
    if (!a.lib) return mnw_fail(ctx, MNW_ERR_CUDA, "mnw_comm_init: %s", a.why);
    if (ctx->comm) { a.CommDestroy((ncclComm_t)ctx->comm); ctx->comm = nullptr; }
    NcclUniqueId uid;
    memcpy(&uid, id, sizeof uid);
    ncclComm_t c = nullptr;
    const ncclResult_t r = a.CommInitRank(&c, nranks, uid, rank);
    if (r != 0) return nccl_fail(ctx, "ncclCommInitRank", r);
    ctx->comm = c; ctx->comm_ranks = nranks; ctx->comm_rank = rank;
    return MNW_OK;
}

int mnw_comm_destroy(mnw_ctx *ctx) {
    if (!ctx) return MNW_ERR_ARG;
    if (ctx->comm) {
        (void)cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->L.stream);
        nccl().CommDestroy((ncclComm_t)ctx->comm);
        ctx->comm = nullptr; ctx->comm_ranks = 0; ctx->comm_rank = 0;
    }
    return MNW_OK;
}

int mnw_comm_size(const mnw_ctx *ctx) { return ctx && ctx->comm ? ctx->comm_ranks : 1; }
int mnw_comm_rank(const mnw_ctx *ctx) { return ctx && ctx->comm ? ctx->comm_rank : 0; }

int mnw_allgather_sizes(mnw_ctx *ctx, const int64_t *local_dev, int64_t count, int64_t *all_dev) {
    if (!ctx || count < 0 || (count > 0 && (!local_dev || !all_dev))) return mnw_fail(ctx, MNW_ERR_ARG, "mnw_allgather_sizes: bad argument");
    (void)cudaSetDevice(ctx->device);
    if (count == 0) return MNW_OK;
    if (!ctx->comm) {   // a context without a communicator is a world of one
        CU(cudaMemcpyAsync(all_dev, local_dev, 8 * (size_t)count, cudaMemcpyDeviceToDevice, ctx->L.stream));
        return MNW_OK;
    }
    const ncclResult_t r = nccl().AllGather(local_dev, all_dev, (size_t)count, NCCL_INT64, (ncclComm_t)ctx->comm, ctx->L.stream);
    if (r != 0) return nccl_fail(ctx, "ncclAllGather", r);
    return MNW_OK;
}

int mnw_sharded_offsets_dev(mnw_ctx *ctx, const int64_t *local_nbytes_dev, int64_t count, int64_t *all_nbytes_dev,
                            int64_t *all_offsets_dev, int64_t *total_dev) {
    int rc = mnw_allgather_sizes(ctx, local_nbytes_dev, count, all_nbytes_dev);
    if (rc) return rc;
    return mnw_scan_offsets_dev(ctx, all_nbytes_dev, count * mnw_comm_size(ctx), 0, all_offsets_dev, total_dev);
}

}  // extern "C"
