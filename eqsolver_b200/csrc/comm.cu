// comm.cu -- the per-iteration exchange between the GPUs of one box: NCCL over NVLink 5 / NVSwitch.
//
// The reference is single-process and has no communication layer; this exists only because the
// population is sharded.  Messages are tiny (one record of 54+n doubles per rank, or one histogram), so the
// cost is launch + NVLink latency, not bandwidth.  NCCL is resolved at run time with dlopen so that
// (a) the library loads on a box without NCCL for single-GPU use and (b) inside a process that already
// loaded torch's bundled libnccl.so.2 the same copy is reused (same SONAME) instead of a second one.
#include <dlfcn.h>
#include <mutex>

#include "common.cuh"

namespace eqs {

namespace {

// the few NCCL declarations this file needs, restated from nccl.h (stable since NCCL 2.0)
typedef struct ncclComm *ncclComm_t;
typedef struct {
    char internal[128];
} ncclUniqueId;
typedef int ncclResult_t;
enum { ncclUint64 = 5, ncclFloat64 = 8 };
enum { ncclSum = 0 };

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

NcclApi &nccl()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) {
            api.error = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "?");
            return;
        }
        auto sym = [&](const char *s) {
            void *p = dlsym(api.handle, s);
            if (!p) api.error = std::string("libnccl lacks ") + s;
            return p;
        };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
        api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    });
    return api;
}

} // namespace

struct Comm {
    ncclComm_t comm = nullptr;
};

int comm_unique_id(void *out128, std::string &err)
{
    NcclApi &a = nccl();
    if (!a.error.empty()) {
        err = a.error;
        return EQS_ERR_EXTERNAL;
    }
    ncclUniqueId id;
    const ncclResult_t r = a.GetUniqueId(&id);
    if (r != 0) {
        err = std::string("ncclGetUniqueId: ") + a.GetErrorString(r);
        return EQS_ERR_EXTERNAL;
    }
    memcpy(out128, &id, sizeof id);
    return EQS_OK;
}

int comm_init(Ctx *ctx, int rank, int world, const void *id128)
{
    if (world < 1 || rank < 0 || rank >= world) return ctx->fail(EQS_ERR_INCORRECT_INPUT, "bad rank %d / world %d", rank, world);
    comm_destroy(ctx);
    ctx->rank = rank;
    ctx->world = world;
    if (world == 1) return EQS_OK;
    NcclApi &a = nccl();
    if (!a.error.empty()) return ctx->fail(EQS_ERR_EXTERNAL, "%s", a.error.c_str());
    EQS_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    Comm *c = new Comm;
    const ncclResult_t r = a.CommInitRank(&c->comm, world, id, rank);
    if (r != 0) {
        delete c;
        ctx->rank = 0;
        ctx->world = 1;
        return ctx->fail(EQS_ERR_EXTERNAL, "ncclCommInitRank: %s", a.GetErrorString(r));
    }
    ctx->comm = c;
    return EQS_OK;
}

void comm_destroy(Ctx *ctx)
{
    if (ctx->comm) {
        if (ctx->comm->comm) nccl().CommDestroy(ctx->comm->comm);
        delete ctx->comm;
        ctx->comm = nullptr;
    }
    ctx->rank = 0;
    ctx->world = 1;
}

int comm_allgather_f64(Ctx *ctx, const double *send, double *recv, int64_t count_per_rank, cudaStream_t s)
{
    if (!ctx->comm) return ctx->fail(EQS_ERR_EXTERNAL, "world > 1 but the context has no communicator");
    NcclApi &a = nccl();
    const ncclResult_t r = a.AllGather(send, recv, (size_t)count_per_rank, ncclFloat64, ctx->comm->comm, s);
    if (r != 0) return ctx->fail(EQS_ERR_EXTERNAL, "ncclAllGather: %s", a.GetErrorString(r));
    return EQS_OK;
}

int comm_allreduce_sum_u64(Ctx *ctx, unsigned long long *buf, int64_t count, cudaStream_t s)
{
    if (!ctx->comm) return ctx->fail(EQS_ERR_EXTERNAL, "world > 1 but the context has no communicator");
    NcclApi &a = nccl();
    const ncclResult_t r = a.AllReduce(buf, buf, (size_t)count, ncclUint64, ncclSum, ctx->comm->comm, s);
    if (r != 0) return ctx->fail(EQS_ERR_EXTERNAL, "ncclAllReduce: %s", a.GetErrorString(r));
    return EQS_OK;
}

} // namespace eqs
