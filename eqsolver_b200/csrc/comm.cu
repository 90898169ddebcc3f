// comm.cu -- the per-iteration exchange between the GPUs of one box over NVLink 5 / NVSwitch.
//
// The reference is single-process and has no communication layer; this exists only because the
// population is sharded.  Messages are tiny (one record of 54+n doubles per rank, or one histogram), so the
// cost is launch + NVLink latency, not bandwidth.  Two mechanisms:
//   * NCCL (bootstrap, init, sequential mode, CrossEntropy phases): ncclAllGather on the library's stream;
//   * peer-memory MAILBOXES (the synchronous ParticleSwarm step): every rank's mailbox is mapped into every peer
//     with CUDA IPC, and the step kernel's last CTA does the exchange itself with remote stores + release/acquire
//     flags (pso.cu: pso_p2p_exchange_apply) -- one kernel per pass at any world size, no collective launch.
//     Measured at N=2 on C2: 7 us per pass over the single-GPU kernel vs 11.5 us for NCCL + apply kernel.
// NCCL is resolved at run time with dlopen so that
// (a) the library loads on a box without NCCL for single-GPU use and (b) inside a process that already
// loaded torch's bundled libnccl.so.2 the same copy is reused (same SONAME) instead of a second one.
#include <dlfcn.h>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace eqs {

namespace {

// the few NCCL declarations this file needs, restated from nccl.h (stable since NCCL 2.0)
typedef struct ncclComm *ncclComm_t;
typedef struct {
    char internal[128];
} ncclUniqueId;
typedef int ncclResult_t;
enum { ncclUint8 = 1, ncclUint64 = 5, ncclFloat64 = 8 };
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
            const char *why = dlerror(); // a second call would return NULL: dlerror() clears the message it hands out
            api.error = std::string("cannot load libnccl.so.2: ") + (why ? why : "?");
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
    // peer mailboxes (see common.cuh)
    char *mailbox = nullptr;            // this rank's mailbox (cudaMalloc)
    std::vector<char *> peer;           // every rank's mailbox as mapped here (peer[rank] == mailbox)
    char **d_peer = nullptr;            // the same table on the device
    bool p2p = false;
};

namespace {

// allocate the mailbox, ship its IPC handle to every rank over NCCL, map the peers' mailboxes
int p2p_setup(Ctx *ctx, Comm *c, int rank, int world)
{
    if (world > MB_MAX_WORLD) return EQS_OK;
    const char *env = getenv("EQS_P2P");
    if (env && !strcmp(env, "0")) return EQS_OK;
    NcclApi &a = nccl();
    cudaStream_t s = ctx->stream;
    EQS_CUDA(ctx, cudaMalloc(&c->mailbox, MB_BYTES));
    EQS_CUDA(ctx, cudaMemsetAsync(c->mailbox, 0, MB_BYTES, s)); // header (counters, flags) and the tags of the LL region
    cudaIpcMemHandle_t mine;
    EQS_CUDA(ctx, cudaIpcGetMemHandle(&mine, c->mailbox));
    unsigned char *d_h = nullptr;
    EQS_CUDA(ctx, cudaMalloc(&d_h, sizeof(mine) * (size_t)(world + 1)));
    EQS_CUDA(ctx, cudaMemcpyAsync(d_h, &mine, sizeof(mine), cudaMemcpyHostToDevice, s));
    ncclResult_t r = a.AllGather(d_h, d_h + sizeof(mine), sizeof(mine), ncclUint8, c->comm, s);
    if (r != 0) return ctx->fail(EQS_ERR_EXTERNAL, "ncclAllGather (IPC handles): %s", a.GetErrorString(r));
    std::vector<cudaIpcMemHandle_t> all(world);
    EQS_CUDA(ctx, cudaMemcpyAsync(all.data(), d_h + sizeof(mine), sizeof(mine) * (size_t)world, cudaMemcpyDeviceToHost, s));
    EQS_CUDA(ctx, cudaStreamSynchronize(s));
    cudaFree(d_h);
    c->peer.assign(world, nullptr);
    bool ok = true;
    for (int q = 0; q < world; ++q) {
        if (q == rank) {
            c->peer[q] = c->mailbox;
            continue;
        }
        void *ptr = nullptr;
        if (cudaIpcOpenMemHandle(&ptr, all[q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError(); // clear: peer mapping is a capability, NCCL remains the exchange
            ok = false;
            break;
        }
        c->peer[q] = (char *)ptr;
    }
    // every rank must agree, otherwise some would wait on mailboxes nobody writes
    unsigned long long *d_ok = nullptr;
    EQS_CUDA(ctx, cudaMalloc(&d_ok, sizeof(unsigned long long)));
    const unsigned long long bad = ok ? 0ull : 1ull;
    EQS_CUDA(ctx, cudaMemcpyAsync(d_ok, &bad, sizeof bad, cudaMemcpyHostToDevice, s));
    r = a.AllReduce(d_ok, d_ok, 1, ncclUint64, ncclSum, c->comm, s);
    if (r != 0) return ctx->fail(EQS_ERR_EXTERNAL, "ncclAllReduce (IPC agreement): %s", a.GetErrorString(r));
    unsigned long long total_bad = 1;
    EQS_CUDA(ctx, cudaMemcpyAsync(&total_bad, d_ok, sizeof total_bad, cudaMemcpyDeviceToHost, s));
    EQS_CUDA(ctx, cudaStreamSynchronize(s));
    cudaFree(d_ok);
    if (total_bad != 0) return EQS_OK; // stays on NCCL
    EQS_CUDA(ctx, cudaMalloc(&c->d_peer, sizeof(char *) * (size_t)world));
    EQS_CUDA(ctx, cudaMemcpyAsync(c->d_peer, c->peer.data(), sizeof(char *) * (size_t)world, cudaMemcpyHostToDevice, s));
    EQS_CUDA(ctx, cudaStreamSynchronize(s));
    c->p2p = true;
    return EQS_OK;
}

} // namespace

bool comm_p2p(Ctx *ctx, char **mailbox, char *const **d_peers)
{
    if (!ctx->comm || !ctx->comm->p2p) return false;
    *mailbox = ctx->comm->mailbox;
    *d_peers = ctx->comm->d_peer;
    return true;
}

long long comm_p2p_timeout_ns()
{
    double seconds = 20.0;
    if (const char *e = getenv("EQS_P2P_TIMEOUT_S")) {
        const double v = atof(e);
        if (v > 0.0) seconds = v;
    }
    return (long long)(seconds * 1e9);
}

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
    const int st = p2p_setup(ctx, c, rank, world);
    if (st != EQS_OK) return st;
    return EQS_OK;
}

void comm_destroy(Ctx *ctx)
{
    if (ctx->comm) {
        Comm *c = ctx->comm;
        cudaDeviceSynchronize();
        for (size_t q = 0; q < c->peer.size(); ++q)
            if (c->peer[q] && c->peer[q] != c->mailbox) cudaIpcCloseMemHandle(c->peer[q]);
        if (c->d_peer) cudaFree(c->d_peer);
        if (c->mailbox) cudaFree(c->mailbox);
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

} // namespace eqs
