// common.cuh -- context, error plumbing, exchange-record layout, block-level reductions and scans.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include "../../include/eqsolver_b200.h"

// No C++ exception may cross the C ABI (include/eqsolver_b200.h): every extern "C" entry point with a body that can
// allocate runs between these two; std::bad_alloc and friends come back as EQS_ERR_EXTERNAL.
#define EQS_BOUNDARY_TRY                                                                                         \
    try {                                                                                                        \
        eqs::boundary_exception_clear();
#define EQS_BOUNDARY_CATCH                                                                                       \
    }                                                                                                            \
    catch (...) { return eqs::boundary_exception(); }

namespace eqs {

// inside a catch (...): the status for "a C++ exception reached the boundary"; the text is kept per thread
int boundary_exception() noexcept;
const char *boundary_exception_text() noexcept; // "" unless the last failure on this thread was such an exception
void boundary_exception_clear() noexcept;


// ---- exchange record (doubles) --------------------------------------------------------------------------
// One record per rank and exchange: the only data that ever crosses NVLink for ParticleSwarm.
//   [0] cost   [1] global particle index (as f64, exact below 2^53; -1 = none)   [2] ring pushes of the
//   init scan   [3] pad   [4..54) last <=50 pushed values in order   [54..104) the global indices of the particles
//   that pushed them (the ranks' pushes are merged by position)   [104..104+n) the particle's position
constexpr int REC_COST = 0, REC_IDX = 1, REC_PUSHES = 2, REC_RING = 4, REC_RPOS = REC_RING + EQS_STALL_WINDOW,
              REC_X = REC_RPOS + EQS_STALL_WINDOW;
__host__ __device__ inline int64_t record_doubles(int n) { return (int64_t)REC_X + n; }

// ---- context -------------------------------------------------------------------------------------------
struct Comm; // comm.cu (NCCL resolved at run time; no link-time dependency)

struct Ctx {
    int device = -1;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    Comm *comm = nullptr;
    int rank = 0, world = 1;
    // message behind the last non-zero status; solvers sharing the context may fail on different threads at once
    mutable std::mutex error_mutex;
    mutable std::string error;

    // Allocation caches.  The reference's benchmark shape (benchmarks/multi.rs:70-100) constructs a solver and runs a
    // whole default solve (100 particles, 50 passes) per sample: cudaMalloc / cudaMallocHost / cudaFree would cost
    // 100x the kernels.  One device arena and a few pinned control blocks are therefore kept per context.
    void *cached_arena = nullptr;
    size_t cached_arena_bytes = 0;
    std::vector<void *> pinned_free; // blocks of kPinnedBytes
    std::mutex cache_mutex;          // solvers sharing a context may be created / destroyed from different threads
    static constexpr size_t kPinnedBytes = 1024;

    static constexpr size_t kArenaCacheMax = (size_t)2 << 30;
    cudaError_t acquire_arena(size_t bytes, void **out, size_t *got)
    {
        std::lock_guard<std::mutex> lock(cache_mutex);
        if (cached_arena && cached_arena_bytes >= bytes) {
            *out = cached_arena;
            *got = cached_arena_bytes;
            cached_arena = nullptr;
            cached_arena_bytes = 0;
            return cudaSuccess;
        }
        *got = bytes;
        return cudaMalloc(out, bytes);
    }
    void release_arena(void *p, size_t bytes)
    {
        if (!p) return;
        std::lock_guard<std::mutex> lock(cache_mutex);
        // keep at most one arena, and only a modest one: the cache exists to take cudaMalloc / cudaFree out of
        // repeated small and medium solves, not to sit on tens of GB after one large solve
        if (bytes > cached_arena_bytes && bytes <= kArenaCacheMax) {
            if (cached_arena) cudaFree(cached_arena);
            cached_arena = p;
            cached_arena_bytes = bytes;
        } else {
            cudaFree(p);
        }
    }
    cudaError_t acquire_pinned(void **out)
    {
        std::lock_guard<std::mutex> lock(cache_mutex);
        if (!pinned_free.empty()) {
            *out = pinned_free.back();
            pinned_free.pop_back();
            return cudaSuccess;
        }
        return cudaMallocHost(out, kPinnedBytes);
    }
    void release_pinned(void *p)
    {
        std::lock_guard<std::mutex> lock(cache_mutex);
        if (p) pinned_free.push_back(p);
    }
    void drop_caches()
    {
        if (cached_arena) cudaFree(cached_arena);
        cached_arena = nullptr;
        cached_arena_bytes = 0;
        for (void *p : pinned_free) cudaFreeHost(p);
        pinned_free.clear();
    }

    int fail(int status, const char *fmt, ...) const
    {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        {
            std::lock_guard<std::mutex> lock(error_mutex);
            error = buf;
        }
        boundary_exception_clear();
        // a failed runtime call (e.g. cudaMalloc -> cudaErrorMemoryAllocation) stays recorded until it is read; clear
        // it, or the launch check of the NEXT, unrelated call on this thread reports it again
        (void)cudaGetLastError();
        return status;
    }
};

} // namespace eqs

// the opaque context of the C ABI
struct eqs_ctx {
    eqs::Ctx c;
};

namespace eqs {

#define EQS_CUDA(ctx, expr)                                                                                      \
    do {                                                                                                         \
        cudaError_t _e = (expr);                                                                                 \
        if (_e != cudaSuccess)                                                                                   \
            return (ctx)->fail(EQS_ERR_EXTERNAL, "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__,   \
                               __LINE__, cudaGetErrorString(_e));                                                \
    } while (0)

#define EQS_TRY(expr)                                                                                            \
    do {                                                                                                         \
        int _s = (expr);                                                                                         \
        if (_s != EQS_OK) return _s;                                                                             \
    } while (0)

// device scratch and CUDA events owned by one call: released on every way out of the function
template <class T> struct DevBuf {
    T *p = nullptr;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf()
    {
        if (p) cudaFree(p);
    }
    cudaError_t alloc(size_t count) { return cudaMalloc(&p, count * sizeof(T)); }
    operator T *() const { return p; }
};
struct Event {
    cudaEvent_t e = nullptr;
    Event() = default;
    Event(const Event &) = delete;
    Event &operator=(const Event &) = delete;
    ~Event()
    {
        if (e) cudaEventDestroy(e);
    }
    cudaError_t create() { return cudaEventCreate(&e); }
    operator cudaEvent_t() const { return e; }
};

// comm.cu
int comm_unique_id(void *out128, std::string &err);
int comm_init(Ctx *ctx, int rank, int world, const void *id128);
void comm_destroy(Ctx *ctx);
int comm_allgather_f64(Ctx *ctx, const double *send, double *recv, int64_t count_per_rank, cudaStream_t s);

// ---- peer-memory mailboxes (NVLink / NVSwitch, CUDA IPC): the exchanges fused into the kernels ----------------------
// One mailbox per rank, mapped into every peer: [exchange counter u64 | pad to 4 KB | words [2][world][MB_LL_CAP] u64].
// Every 8-byte word carries 32 bits of payload under a 32-bit tag (the exchange number), so data and "it has arrived"
// travel in ONE store: no fence, no flag, one NVLink one-way latency (ll_* below; NCCL calls this protocol LL).  The flag
// protocol this replaced in round 2 -- plain stores, __threadfence_system, a release store of the exchange number, an
// acquire spin -- paid the fence's NVLink round trip before the flag could leave: 7 us per exchange on two GPUs against 3
// (profiles/r02_sweep.md section 4).
constexpr size_t MB_LLSEQ = 0, MB_LL = 4096;
constexpr int MB_MAX_WORLD = 16;
constexpr long long MB_LL_CAP = 2 * (2 + 12800) + 4; // two words per 8-byte value: (cost, index, n <= 12800 coordinates)
constexpr size_t MB_BYTES = MB_LL + 2ull * MB_MAX_WORLD * MB_LL_CAP * sizeof(unsigned long long);
// returns false when the context has no peer mailboxes (single rank, IPC unavailable, or EQS_P2P=0)
bool comm_p2p(Ctx *ctx, char **mailbox, char *const **d_peers);
// How long a kernel waits in a mailbox for a peer before it gives up (the solve then fails with ExternalError instead of
// hanging the GPU): EQS_P2P_TIMEOUT_S seconds, default 20.  All ranks must therefore enqueue the same pass within that
// time of each other ("Co-scheduling").
long long comm_p2p_timeout_ns();

// Dynamic shared memory beyond this many bytes is requested with cudaFuncAttributeMaxDynamicSharedMemorySize.  The
// limit without the opt-in is 48 KB for STATIC + dynamic together, so the threshold leaves room for the kernels' static
// arrays (a few KB): asking at exactly 48 KB of dynamic memory fails the launch with cudaErrorInvalidValue.
constexpr size_t kSmemOptIn = 32 * 1024;

// ---- device helpers ------------------------------------------------------------------------------------
// Index checks of our own (compute-sanitizer is closed on the B200 pool, profiles/r02_sanitizer.md): a build with
// -DEQS_BOUNDS_CHECK (benchmarks/build_variant.py bounds -DEQS_BOUNDS_CHECK) asserts every computed offset of the
// population, cost, candidate, elite and histogram accesses against the extent of its buffer; a violation traps the
// kernel, which the next API call reports as ExternalError.  The default build compiles the checks away.
#ifdef EQS_BOUNDS_CHECK
#include <cassert>
#define EQS_CHECK(cond) assert(cond)
#else
#define EQS_CHECK(cond) ((void)0)
#endif

__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// ---- low-latency exchange: a double travels as two 8-byte words {tag, low half} and {tag, high half}; an aligned 8-byte
// store arrives whole, so a receiver that sees the tag of THIS exchange in a word has the payload of this exchange.  Every
// rank runs the same sequence of exchanges on a context (counter in its mailbox), two slots deep: a rank can be one
// exchange ahead of a peer, never two, because it needs that peer's words of exchange s+1, which the peer sends only
// after it has read what it wanted of exchange s.  No deadlock: a CTA only ever waits for kernels running on OTHER GPUs,
// and gives up after timeout_ns (the solve then fails with ExternalError instead of hanging the GPU).
struct LLExchange {
    char *mailbox;
    char *const *peers;
    int rank, world, slot;
    unsigned tag;
    unsigned long long seq;
};
// all threads of the calling CTA; ends with a barrier.  ll_end (one thread, after the last ll_get) closes the exchange.
__device__ __forceinline__ LLExchange ll_begin(char *mailbox, char *const *peers, int rank, int world)
{
    __shared__ unsigned long long s_llseq;
    if (threadIdx.x == 0) s_llseq = *reinterpret_cast<unsigned long long *>(mailbox + MB_LLSEQ) + 1;
    __syncthreads();
    LLExchange x;
    x.mailbox = mailbox;
    x.peers = peers;
    x.rank = rank;
    x.world = world;
    x.seq = s_llseq;
    x.slot = (int)(x.seq & 1ull);
    x.tag = (unsigned)(x.seq % 0xffffffffull) + 1u; // never 0 (the mailbox starts zeroed)
    return x;
}
__device__ __forceinline__ void ll_end(const LLExchange &x) { *reinterpret_cast<unsigned long long *>(x.mailbox + MB_LLSEQ) = x.seq; }
// 8-byte value number d of this rank's record, to every peer
__device__ __forceinline__ void ll_put_u64(const LLExchange &x, long long d, unsigned long long b)
{
    const unsigned long long t = (unsigned long long)x.tag << 32;
    const unsigned long long w0 = t | (b & 0xffffffffull), w1 = t | (b >> 32);
    for (int r = 0; r < x.world; ++r) {
        if (r == x.rank) continue;
        unsigned long long *dst = reinterpret_cast<unsigned long long *>(x.peers[r] + MB_LL) +
                                  ((size_t)x.slot * x.world + x.rank) * MB_LL_CAP + 2 * d;
        asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(dst), "l"(w0), "l"(w1) : "memory");
    }
}
__device__ __forceinline__ void ll_put(const LLExchange &x, int d, double v) { ll_put_u64(x, d, (unsigned long long)__double_as_longlong(v)); }
// one look at value number d of rank r's record in this rank's own mailbox: true when both halves carry this exchange's tag
__device__ __forceinline__ bool ll_peek_u64(const LLExchange &x, int r, long long d, unsigned long long &b)
{
    const unsigned long long *src = reinterpret_cast<const unsigned long long *>(x.mailbox + MB_LL) +
                                    ((size_t)x.slot * x.world + r) * MB_LL_CAP + 2 * d;
    // A weak load that bypasses L1 (.cg: the peers' stores land in this GPU's L2, the home of this memory).  It needs no
    // ordering of its own -- a word either carries this exchange's tag, and then its payload, or it is looked at again --
    // and unlike ld.relaxed.sys / ld.volatile, which the hardware completes one after the other, the eight looks of a trip
    // of ll_allgather_u64 overlap (on 8 GPUs a 2048-word histogram exchange took 26 us with the ordered loads).
    unsigned long long w0, w1;
    asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(src) : "memory");
    b = (w1 << 32) | (w0 & 0xffffffffull);
    return (unsigned)(w0 >> 32) == x.tag && (unsigned)(w1 >> 32) == x.tag;
}
// value number d of rank r's record (r != this rank); false after timeout_ns without it
__device__ __forceinline__ bool ll_get_u64(const LLExchange &x, int r, long long d, unsigned long long &b, long long timeout_ns)
{
    unsigned long long t0 = 0;
    for (unsigned spins = 0;; ++spins) {
        if (ll_peek_u64(x, r, d, b)) return true;
        if ((spins & 31u) == 31u) {
            const unsigned long long now = globaltimer_ns();
            if (t0 == 0) t0 = now;
            else if ((long long)(now - t0) > timeout_ns) return false;
        }
    }
}
__device__ __forceinline__ bool ll_get(const LLExchange &x, int r, int d, double &v, long long timeout_ns)
{
    unsigned long long b = 0;
    const bool ok = ll_get_u64(x, r, d, b, timeout_ns);
    v = __longlong_as_double((long long)b);
    return ok;
}

// ALLGATHER with the LL protocol, executed by ONE CTA (all of its threads call): `words` 8-byte words at `send` (global
// memory, complete for this CTA) from every rank into recv[r * stride + i] (this rank's local memory; its own words are
// copied).  A thread looks at eight of its words per trip before it waits for any of them, so the loads overlap.  Ends with
// a barrier; false (in every thread) when a peer's words did not arrive within timeout_ns.
__device__ __forceinline__ bool ll_allgather_u64(char *mailbox, char *const *peers, int rank, int world,
                                                 const unsigned long long *send, long long words, unsigned long long *recv,
                                                 long long stride, long long timeout_ns)
{
    __shared__ int s_ll_bad;
    if (threadIdx.x == 0) s_ll_bad = 0;
    const LLExchange X = ll_begin(mailbox, peers, rank, world);
    for (long long i = threadIdx.x; i < words; i += blockDim.x) {
        const unsigned long long v = __ldcg(send + i);
        ll_put_u64(X, i, v);
        recv[rank * stride + i] = v;
    }
    constexpr int U = 8;
    const long long total = words * (world - 1);
    for (long long base = threadIdx.x; base < total; base += (long long)blockDim.x * U) {
        unsigned long long b[U];
        bool ok[U];
        int rr[U];
        long long ii[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long idx = base + (long long)u * blockDim.x;
            ok[u] = true;
            if (idx < total) {
                const int pr = (int)(idx / words);
                rr[u] = pr < rank ? pr : pr + 1;
                ii[u] = idx % words;
                ok[u] = ll_peek_u64(X, rr[u], ii[u], b[u]);
            } else {
                rr[u] = -1;
                ii[u] = 0;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (rr[u] < 0) continue;
            if (!ok[u] && !ll_get_u64(X, rr[u], ii[u], b[u], timeout_ns)) s_ll_bad = 1;
            recv[rr[u] * stride + ii[u]] = b[u];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) ll_end(X);
    return s_ll_bad == 0;
}

__device__ __forceinline__ double sanitize_cost(double c) { return (c != c) ? __longlong_as_double(0x7ff0000000000000LL) : c; }

__device__ __forceinline__ double shfl_down_f64(double v, int delta)
{
    return __shfl_down_sync(0xffffffffu, v, delta);
}

// (cost, index) min with the lowest index winning ties -- the reference's strict `<` in particle order
struct MinLoc {
    double c;
    long long i;
};

__device__ __forceinline__ MinLoc minloc_better(MinLoc a, MinLoc b)
{
    const bool take_b = (b.c < a.c) || (b.c == a.c && b.i < a.i);
    return take_b ? b : a;
}

__device__ __forceinline__ MinLoc warp_minloc(MinLoc v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        MinLoc w;
        w.c = __shfl_down_sync(0xffffffffu, v.c, o);
        w.i = __shfl_down_sync(0xffffffffu, v.i, o);
        v = minloc_better(v, w);
    }
    return v;
}

// all threads of the block must call; result valid in thread 0.  smem: >= 32 MinLoc
__device__ __forceinline__ MinLoc block_minloc(MinLoc v, MinLoc *smem)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_minloc(v);
    __syncthreads();
    if (lane == 0) smem[wid] = v;
    __syncthreads();
    if (wid == 0) {
        MinLoc w = lane < nw ? smem[lane] : MinLoc{__longlong_as_double(0x7ff0000000000000LL), 0x7fffffffffffffffLL};
        w = warp_minloc(w);
        v = w;
    }
    return v;
}

__device__ __forceinline__ long long warp_min_i64(long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const long long w = __shfl_down_sync(0xffffffffu, v, o);
        v = w < v ? w : v;
    }
    return v;
}

// block-wide EXCLUSIVE scan with an associative op; every thread calls; `total` = reduction of the block.
// smem: >= 33 T.  Ordered by threadIdx.x.
template <class T, class Op> __device__ __forceinline__ T block_exclusive_scan(T v, T identity, Op op, T *smem, T &total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    T incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const T up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl = op(up, incl);
    }
    __syncthreads();
    if (lane == 31) smem[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        T w = lane < nw ? smem[lane] : identity;
        T winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const T up = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc = op(up, winc);
        }
        T wexc = __shfl_up_sync(0xffffffffu, winc, 1);
        if (lane == 0) wexc = identity;
        smem[lane] = wexc;
        if (lane == 31) smem[32] = winc;
    }
    __syncthreads();
    const T warp_prefix = smem[wid];
    total = smem[32];
    T excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = identity;
    return op(warp_prefix, excl);
}

struct OpMin {
    __device__ __forceinline__ double operator()(double a, double b) const { return b < a ? b : a; }
};
struct OpAddI {
    __device__ __forceinline__ int operator()(int a, int b) const { return a + b; }
};
struct OpAddLL {
    __device__ __forceinline__ long long operator()(long long a, long long b) const { return a + b; }
};

} // namespace eqs
