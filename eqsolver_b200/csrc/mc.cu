// mc.cu -- MonteCarlo integration over an orthotope on sm_100a.
//
// Replaces MonteCarlo::integrate_with_rng / integrate_with_sampler, reference src/integrators/monte_carlo.rs:143-212,
// and MeanVariance::from_iterator_using_welford, src/lib.rs:172-202.  Same sample -> evaluate -> reduce shape as the
// CrossEntropy sample kernel: draws (Philox stream 3, one block per dimension pair) and the integrand stay in
// registers, nothing is read from HBM.  The reference's sequential Welford recurrence becomes per-thread Welford over a
// grid-stride subsequence + Chan's pairwise merge (lanes, warps, CTAs, ranks, always in the same fixed order, so the result is
// deterministic); it agrees with the sequential recurrence to rounding (tests: 1e-12 on the mean).
#include <algorithm>
#include <cmath>
#include <functional>
#include <limits>
#include <string>
#include <vector>

#define EQS_COS_TABLE 1 // issue-bound sample kernels: constant-operand cos (devmath.cuh)
#include "common.cuh"
#include "objectives.cuh"
#include "philox.cuh"

namespace eqs {

namespace {

constexpr int MC_BLOCK = 256;
constexpr uint32_t STREAM_MC = 3u;

struct Moments { // count, mean, sum of squared deviations
    double n, mean, m2;
};

__host__ __device__ inline Moments merge(const Moments &a, const Moments &b)
{
    if (b.n == 0.0) return a;
    if (a.n == 0.0) return b;
    const double n = a.n + b.n, delta = b.mean - a.mean;
    return Moments{n, a.mean + delta * (b.n / n), a.m2 + b.m2 + delta * delta * (a.n * b.n / n)};
}

struct McDev {
    const double *from, *to; // [n]
    const double *replay;    // [count][n] or nullptr
    Moments *part;           // [grid]
    Moments *out;            // this rank's result
    unsigned int *counter;
    long long first, nloc;
    int n;
    PhiloxKey key;
};

template <class Obj, bool REPLAY> __global__ void __launch_bounds__(MC_BLOCK) mc_kernel(const McDev P)
{
    extern __shared__ double s_box[]; // from[n], scale[n]
    __shared__ Moments s_w[MC_BLOCK / 32];
    __shared__ int s_last;
    const int n = P.n, tid = threadIdx.x;
    for (int d = tid; d < n; d += blockDim.x) {
        s_box[d] = P.from[d];
        s_box[n + d] = P.to[d] - P.from[d]; // Uniform::new_inclusive(x, y): u * scale + low, :191-195
    }
    __syncthreads();
    Moments m{0.0, 0.0, 0.0};
    for (long long li = (long long)blockIdx.x * blockDim.x + tid; li < P.nloc; li += (long long)gridDim.x * blockDim.x) {
        const long long gi = P.first + li;
        double sample;
        {
            ObjAcc acc;
            double xs_prev = 0.0;
            (void)xs_prev;
            if constexpr (Obj::streaming) Obj::begin(acc, n);
            auto coord = [&](double u, int d) {
                const double x = u * s_box[n + d] + s_box[d];
                if constexpr (Obj::streaming) Obj::feed(acc, x, d);
            };
            if constexpr (REPLAY) {
                for (int d = 0; d < n; ++d) coord(P.replay[gi * n + d], d);
            } else {
                int d = 0;
                if constexpr (Obj::streaming) { // four coordinates (two Philox blocks) at a time: objectives.cuh feed_vec
                    for (; d + 4 <= n; d += 4) {
                        const uint4 w0 = philox_draw(P.key, STREAM_MC, 0u, gi, (uint32_t)(d >> 1));
                        const uint4 w1 = philox_draw(P.key, STREAM_MC, 0u, gi, (uint32_t)(d >> 1) + 1u);
                        const double u[4] = {u01(w0.x, w0.y), u01(w0.z, w0.w), u01(w1.x, w1.y), u01(w1.z, w1.w)};
                        double x[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) x[j] = u[j] * s_box[n + d + j] + s_box[d + j];
                        objective_feed_vec<Obj, 4>(acc, x, d);
                    }
                }
                for (; d < n; d += 2) {
                    const uint4 w = philox_draw(P.key, STREAM_MC, 0u, gi, (uint32_t)(d >> 1));
                    coord(u01(w.x, w.y), d);
                    if (d + 1 < n) coord(u01(w.z, w.w), d + 1);
                }
            }
            if constexpr (Obj::streaming) sample = Obj::end(acc, n);
            else sample = 0.0;
        }
        // Welford, lib.rs:184-186
        m.n += 1.0;
        const double delta = sample - m.mean;
        m.mean = m.mean + delta / m.n;
        m.m2 = m.m2 + delta * (sample - m.mean);
    }
    // lanes -> warp -> CTA, always merging the lower index range on the left
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        Moments b;
        b.n = __shfl_down_sync(0xffffffffu, m.n, o);
        b.mean = __shfl_down_sync(0xffffffffu, m.mean, o);
        b.m2 = __shfl_down_sync(0xffffffffu, m.m2, o);
        if ((tid & (2 * o - 1)) == 0) m = merge(m, b);
    }
    if ((tid & 31) == 0) s_w[tid >> 5] = m;
    __syncthreads();
    if (tid == 0) {
        Moments b = s_w[0];
        for (int w = 1; w < MC_BLOCK / 32; ++w) b = merge(b, s_w[w]);
        P.part[blockIdx.x] = b;
        __threadfence();
        s_last = atomicAdd(P.counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last && tid == 0) {
        __threadfence();
        Moments b{0.0, 0.0, 0.0};
        for (unsigned j = 0; j < gridDim.x; ++j) {
            Moments pj;
            pj.n = __ldcg(&P.part[j].n);
            pj.mean = __ldcg(&P.part[j].mean);
            pj.m2 = __ldcg(&P.part[j].m2);
            b = merge(b, pj);
        }
        *P.out = b;
        *P.counter = 0;
    }
}


// ---- MISER (miser.rs:301-443): all Welford batches (variance probes and leaves) of one recursion level ------------

constexpr int MS_BLOCK = 256;
// points per thread and chunk: a function of the batch size alone (small batches spread over more CTAs)
__host__ __device__ inline int miser_spt(long long count) { return count <= 65536 ? 1 : 4; }
__host__ __device__ inline long long miser_chunk(long long count) { return (long long)MS_BLOCK * miser_spt(count); }
constexpr uint32_t STREAM_MISER = 4u;

struct MiserBatch {      // `count` points uniform in the node's region, dimension d narrowed to one half (d < 0: leaf)
    long long count;
    long long chunk0;    // first chunk of this batch in the launch
    long long gbase;     // batch id << 40: high part of the Philox index
    double mid;          // the bisection coordinate (probes)
    int slot;            // row of the region arrays
    int d;
    int upper;           // 0: [from_d, mid], 1: [mid, to_d]
    uint32_t node_id;    // Philox counter word 3
};

struct MiserDev {
    const MiserBatch *batch;
    const double *rfrom, *rto; // [slots][n]
    Moments *part;             // [nchunks]
    Moments *out;              // [nbatch]
    long long nchunks;
    int nbatch, n;
    PhiloxKey key;
};

__device__ __forceinline__ Moments shfl_down(const Moments &m, int o)
{
    Moments b;
    b.n = __shfl_down_sync(0xffffffffu, m.n, o);
    b.mean = __shfl_down_sync(0xffffffffu, m.mean, o);
    b.m2 = __shfl_down_sync(0xffffffffu, m.m2, o);
    return b;
}

// one CTA per chunk of miser_chunk(count) points (grid-stride); the partition into chunks, threads and merge order depends on
// the batch sizes only, never on the grid, so a level is bit-reproducible
template <class Obj> __global__ void __launch_bounds__(MS_BLOCK) miser_chunk_kernel(const MiserDev P)
{
    extern __shared__ double s_box[]; // low[n], scale[n]
    __shared__ Moments s_w[MS_BLOCK / 32];
    __shared__ int s_b;
    const int n = P.n, tid = threadIdx.x;
    for (long long chunk = blockIdx.x; chunk < P.nchunks; chunk += gridDim.x) {
        if (tid == 0) { // the batch owning this chunk: last b with chunk0 <= chunk
            int lo = 0, hi = P.nbatch - 1;
            while (lo < hi) {
                const int md = (lo + hi + 1) >> 1;
                if (P.batch[md].chunk0 <= chunk) lo = md;
                else hi = md - 1;
            }
            s_b = lo;
        }
        __syncthreads();
        const MiserBatch B = P.batch[s_b];
        for (int d = tid; d < n; d += MS_BLOCK) {
            double lo = P.rfrom[(size_t)B.slot * n + d], hi = P.rto[(size_t)B.slot * n + d];
            if (d == B.d) { // lower_d_sampler / upper_d_sampler, miser.rs:366-367
                if (B.upper) lo = B.mid;
                else hi = B.mid;
            }
            s_box[d] = lo;
            s_box[n + d] = hi - lo;
        }
        __syncthreads();
        Moments m{0.0, 0.0, 0.0};
        const int spt = miser_spt(B.count);
        const long long base = (chunk - B.chunk0) * ((long long)MS_BLOCK * spt);
#pragma unroll 1
        for (int k = 0; k < spt; ++k) {
            const long long i = base + (long long)k * MS_BLOCK + tid;
            if (i >= B.count) break;
            ObjAcc acc;
            Obj::begin(acc, n);
            int d = 0;
            for (; d + 4 <= n; d += 4) {
                const uint4 w0 = philox_draw(P.key, STREAM_MISER, B.node_id, B.gbase | i, (uint32_t)(d >> 1));
                const uint4 w1 = philox_draw(P.key, STREAM_MISER, B.node_id, B.gbase | i, (uint32_t)(d >> 1) + 1u);
                const double u[4] = {u01(w0.x, w0.y), u01(w0.z, w0.w), u01(w1.x, w1.y), u01(w1.z, w1.w)};
                double x[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) x[j] = u[j] * s_box[n + d + j] + s_box[d + j];
                objective_feed_vec<Obj, 4>(acc, x, d);
            }
            for (; d < n; d += 2) {
                const uint4 w = philox_draw(P.key, STREAM_MISER, B.node_id, B.gbase | i, (uint32_t)(d >> 1));
                Obj::feed(acc, u01(w.x, w.y) * s_box[n + d] + s_box[d], d);
                if (d + 1 < n) Obj::feed(acc, u01(w.z, w.w) * s_box[n + d + 1] + s_box[d + 1], d + 1);
            }
            const double sample = Obj::end(acc, n);
            m.n += 1.0; // Welford, lib.rs:184-186
            const double delta = sample - m.mean;
            m.mean = m.mean + delta / m.n;
            m.m2 = m.m2 + delta * (sample - m.mean);
        }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const Moments b = shfl_down(m, o);
            if ((tid & (2 * o - 1)) == 0) m = merge(m, b);
        }
        if ((tid & 31) == 0) s_w[tid >> 5] = m;
        __syncthreads();
        if (tid == 0) {
            Moments b = s_w[0];
            for (int w = 1; w < MS_BLOCK / 32; ++w) b = merge(b, s_w[w]);
            P.part[chunk] = b;
        }
    }
}

// one warp per batch: lanes merge contiguous runs of chunk partials, then the lanes merge in order
__global__ void __launch_bounds__(MS_BLOCK) miser_merge_kernel(const MiserDev P)
{
    const int b = (int)(((long long)blockIdx.x * MS_BLOCK + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (b >= P.nbatch) return;
    const MiserBatch B = P.batch[b];
    const long long csz = miser_chunk(B.count), nch = (B.count + csz - 1) / csz, per = (nch + 31) / 32;
    Moments m{0.0, 0.0, 0.0};
    for (long long c = lane * per; c < nch && c < (lane + 1) * per; ++c) m = merge(m, P.part[B.chunk0 + c]);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const Moments x = shfl_down(m, o);
        if ((lane & (2 * o - 1)) == 0) m = merge(m, x);
    }
    if (lane == 0) P.out[b] = m;
}

} // namespace

int validate_mc_params(const eqs_mc_params *p, std::string *why)
{
    auto bad = [&](int st, const char *m) {
        if (why) *why = m;
        return st;
    };
    if (!p) return bad(EQS_ERR_INCORRECT_INPUT, "null parameters");
    if (p->n <= 0 || p->n > 12800) return bad(EQS_ERR_INCORRECT_INPUT, "dimension n out of range");
    if (p->integrand_id < 0 || p->integrand_id >= EQS_OBJ_COUNT) return bad(EQS_ERR_INCORRECT_INPUT, "unknown integrand id");
    if ((p->integrand_id == EQS_OBJ_THREE_CIRCLES || p->integrand_id == EQS_OBJ_SINCOS) && p->n != 2)
        return bad(EQS_ERR_INCORRECT_INPUT, "this integrand needs n = 2");
    if (!p->from || !p->to) return bad(EQS_ERR_INCORRECT_INPUT, "null bounds");
    double volume = 1.0;
    for (int d = 0; d < p->n; ++d) { // Uniform::new_inclusive -> ExternalError, monte_carlo.rs:171-172,191-195
        const double lo = p->from[d], hi = p->to[d];
        if (!(lo <= hi)) return bad(EQS_ERR_EXTERNAL, "low > high (or equal if exclusive) in uniform distribution");
        if (!std::isfinite(lo) || !std::isfinite(hi) || !std::isfinite(hi - lo))
            return bad(EQS_ERR_EXTERNAL, "Non-finite range in uniform distribution");
        volume *= hi - lo;
    }
    if (volume < 0.0) return bad(EQS_ERR_INCORRECT_INPUT, "the input volume should be non-negative"); // :148-152
    if (p->sample_count < 2) return bad(EQS_ERR_INCORRECT_INPUT, "the iterator must be over at least 2 elements"); // lib.rs:190-193
    return EQS_OK;
}


int validate_miser_params(const eqs_miser_params *p, std::string *why)
{
    auto bad = [&](int st, const char *m) {
        if (why) *why = m;
        return st;
    };
    if (!p) return bad(EQS_ERR_INCORRECT_INPUT, "null parameters");
    if (p->n <= 0 || p->n > 12800) return bad(EQS_ERR_INCORRECT_INPUT, "dimension n out of range");
    if (p->integrand_id < 0 || p->integrand_id >= EQS_OBJ_COUNT) return bad(EQS_ERR_INCORRECT_INPUT, "unknown integrand id");
    if ((p->integrand_id == EQS_OBJ_THREE_CIRCLES || p->integrand_id == EQS_OBJ_SINCOS) && p->n != 2)
        return bad(EQS_ERR_INCORRECT_INPUT, "this integrand needs n = 2");
    if (!p->from || !p->to) return bad(EQS_ERR_INCORRECT_INPUT, "null bounds");
    if (p->sample_count < 0 || p->sample_count > (1LL << 40)) return bad(EQS_ERR_INCORRECT_INPUT, "sample_count out of range");
    if (p->minimum_dimensional_sample_count < 0 || p->minimum_dimensional_sample_count > (1LL << 40) / p->n)
        return bad(EQS_ERR_INCORRECT_INPUT, "minimum_dimensional_sample_count out of range");
    if (p->maximum_depth < 0 || p->maximum_depth > 30) return bad(EQS_ERR_INCORRECT_INPUT, "maximum_depth out of range (0..30)");
    return EQS_OK;
}

namespace {

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

inline double host_u01(uint32_t w_lo, uint32_t w_hi)
{
    return (double)((((uint64_t)(w_hi & 0xFFFFFu)) << 32) | (uint64_t)w_lo) * 0x1.0p-52;
}

// Uniform::new_inclusive(lo, hi) of rand_distr: lo <= hi, both and their difference finite (miser.rs:296-299)
inline bool uniform_ok(double lo, double hi) { return lo <= hi && std::isfinite(lo) && std::isfinite(hi) && std::isfinite(hi - lo); }

struct MiserNode {
    std::vector<double> from, to;
    long long count = 0;      // points allotted (sample_count of miser_recurse)
    int depth_remaining = 0;
    uint32_t id = 1;          // root 1, children 2 id and 2 id + 1
    int status = EQS_OK;
    const char *why = "";
    bool leaf = false;
    int lower = -1, upper = -1; // children
    // probes
    std::vector<double> mid;
    int first_batch = -1;
    double mean = 0.0, variance = 0.0;
};

} // namespace

} // namespace eqs

using namespace eqs;

extern "C" {

int eqs_mc_validate(const eqs_mc_params *p) { return validate_mc_params(p, nullptr); }

int eqs_mc_integrate(eqs_ctx *ctx, const eqs_mc_params *p, double *mean, double *variance, eqs_report *report)
{
    EQS_BOUNDARY_TRY
    if (!ctx) return EQS_ERR_INCORRECT_INPUT;
    Ctx *c = &ctx->c;
    if (!mean || !variance) return c->fail(EQS_ERR_INCORRECT_INPUT, "null result pointers");
    std::string why;
    const int st = validate_mc_params(p, &why);
    if (st != EQS_OK) return c->fail(st, "%s", why.c_str());
    bool streaming = true;
    dispatch_objective(p->integrand_id, [&]<class Obj>() { streaming = Obj::streaming; });
    if (!streaming) return c->fail(EQS_ERR_INCORRECT_INPUT, "random-access integrands are not supported by the Monte Carlo kernel");
    EQS_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const int n = p->n, world = c->world, rank = c->rank;
    int64_t first = 0, nloc = 0;
    eqs_shard_range(p->sample_count, world, rank, &first, &nloc);
    const int grid = (int)std::min<long long>(std::max<long long>(1, (nloc + MC_BLOCK - 1) / MC_BLOCK), 8LL * c->sm_count);
    // arena: from, to, partials, result (+ all ranks' results), counter
    const size_t o_from = 0, o_to = (size_t)n * 8, o_part = 2 * (size_t)n * 8;
    const size_t o_out = o_part + (size_t)grid * sizeof(Moments), o_all = o_out + sizeof(Moments);
    const size_t o_cnt = o_all + (size_t)world * sizeof(Moments), total = o_cnt + 256;
    void *arena = nullptr;
    size_t got = 0;
    EQS_CUDA(c, c->acquire_arena(total, &arena, &got));
    char *base = (char *)arena;
    double *d_replay = nullptr;
    int rc = EQS_OK;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    auto cleanup = [&]() {
        cudaStreamSynchronize(s);
        if (d_replay) cudaFree(d_replay);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        c->release_arena(arena, got);
    };
#define MC_TRY(expr)                                                                                             \
    do {                                                                                                         \
        cudaError_t _e = (expr);                                                                                 \
        if (_e != cudaSuccess) {                                                                                 \
            cleanup();                                                                                           \
            return c->fail(EQS_ERR_EXTERNAL, "CUDA error %s: %s", cudaGetErrorName(_e), cudaGetErrorString(_e)); \
        }                                                                                                        \
    } while (0)
    MC_TRY(cudaMemcpyAsync(base + o_from, p->from, (size_t)n * 8, cudaMemcpyHostToDevice, s));
    MC_TRY(cudaMemcpyAsync(base + o_to, p->to, (size_t)n * 8, cudaMemcpyHostToDevice, s));
    MC_TRY(cudaMemsetAsync(base + o_cnt, 0, 256, s));
    if (p->replay_u) {
        const size_t bytes = (size_t)p->sample_count * n * 8;
        MC_TRY(cudaMalloc(&d_replay, bytes));
        MC_TRY(cudaMemcpyAsync(d_replay, p->replay_u, bytes, cudaMemcpyHostToDevice, s));
    }
    McDev D{};
    D.from = (const double *)(base + o_from);
    D.to = (const double *)(base + o_to);
    D.replay = d_replay;
    D.part = (Moments *)(base + o_part);
    D.out = (Moments *)(base + o_out);
    D.counter = (unsigned int *)(base + o_cnt);
    D.first = first;
    D.nloc = nloc;
    D.n = n;
    D.key = philox_key(p->seed);
    MC_TRY(cudaEventCreate(&e0));
    MC_TRY(cudaEventCreate(&e1));
    MC_TRY(cudaEventRecord(e0, s));
    const size_t smem = (size_t)2 * n * sizeof(double);
    dispatch_objective(p->integrand_id, [&]<class Obj>() {
        if constexpr (Obj::streaming) {
            auto run = [&](auto kernel) {
                if (smem > kSmemOptIn) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                kernel<<<grid, MC_BLOCK, smem, s>>>(D);
            };
            if (d_replay) run(mc_kernel<Obj, true>);
            else run(mc_kernel<Obj, false>);
        }
    });
    MC_TRY(cudaGetLastError());
    Moments *all = (Moments *)(base + o_all);
    if (world > 1) {
        rc = comm_allgather_f64(c, (const double *)D.out, (double *)all, 3, s);
        if (rc != EQS_OK) {
            cleanup();
            return rc;
        }
    }
    MC_TRY(cudaEventRecord(e1, s));
    std::vector<Moments> host((size_t)world);
    MC_TRY(cudaMemcpyAsync(host.data(), world > 1 ? (const void *)all : (const void *)D.out, sizeof(Moments) * (size_t)world,
                           cudaMemcpyDeviceToHost, s));
    MC_TRY(cudaStreamSynchronize(s));
    Moments m{0.0, 0.0, 0.0};
    for (int r = 0; r < world; ++r) m = merge(m, host[(size_t)r]); // ranks in index order
    double volume = 1.0;
    for (int d = 0; d < n; ++d) volume *= p->to[d] - p->from[d]; // (to - &from).product(), :197
    const double var = m.m2 / (m.n - 1.0);                        // bessel, lib.rs:195-199
    *mean = m.mean * volume;                                       // scale_mean, lib.rs:143-148
    *variance = var * volume * volume;
    if (report) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        *report = eqs_report{};
        report->iters_run = 1;
        report->evals = p->sample_count;
        report->best_cost = *mean;
        report->world = world;
        report->kernel_launches = 1;
        report->loop_ms = ms;
    }
    cleanup();
#undef MC_TRY
    return EQS_OK;
    EQS_BOUNDARY_CATCH
}


int eqs_miser_validate(const eqs_miser_params *p) { return validate_miser_params(p, nullptr); }

int eqs_miser_integrate(eqs_ctx *ctx, const eqs_miser_params *p, double *mean, double *variance, eqs_report *report)
{
    EQS_BOUNDARY_TRY
    if (!ctx) return EQS_ERR_INCORRECT_INPUT;
    Ctx *c = &ctx->c;
    if (!mean || !variance) return c->fail(EQS_ERR_INCORRECT_INPUT, "null result pointers");
    std::string why;
    const int vst = validate_miser_params(p, &why);
    if (vst != EQS_OK) return c->fail(vst, "%s", why.c_str());
    bool streaming = true;
    dispatch_objective(p->integrand_id, [&]<class Obj>() { streaming = Obj::streaming; });
    if (!streaming) return c->fail(EQS_ERR_INCORRECT_INPUT, "random-access integrands are not supported by the MISER kernel");
    EQS_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const int n = p->n;
    const PhiloxKey key = philox_key(p->seed);
    const long long min_count = p->minimum_dimensional_sample_count * n; // miser.rs:319
    const long long ctl = (long long)(2 * n) << 40, leaf_base = (long long)(2 * n + 1) << 40;

    std::vector<MiserNode> nodes(1);
    nodes[0].from.assign(p->from, p->from + n);
    nodes[0].to.assign(p->to, p->to + n);
    nodes[0].count = p->sample_count;
    nodes[0].depth_remaining = p->maximum_depth;
    std::vector<int> level{0};
    long long evals = 0, launches = 0;
    int levels = 0;
    float total_ms = 0.f;
    Event e0, e1;
    EQS_CUDA(c, e0.create());
    EQS_CUDA(c, e1.create());
    int rc = EQS_OK;

    while (!level.empty() && rc == EQS_OK) {
        ++levels;
        // ---- classify the nodes of this level, lay out their batches ------------------------------------------
        std::vector<MiserBatch> batches;
        std::vector<double> rfrom, rto;
        long long nchunks = 0;
        auto add_batch = [&](int slot, int d, int upper, double mid, long long count, long long gbase, uint32_t id) {
            MiserBatch b{};
            b.count = count;
            b.chunk0 = nchunks;
            b.gbase = gbase;
            b.mid = mid;
            b.slot = slot;
            b.d = d;
            b.upper = upper;
            b.node_id = id;
            batches.push_back(b);
            nchunks += (count + miser_chunk(count) - 1) / miser_chunk(count);
            evals += count;
        };
        for (size_t li = 0; li < level.size(); ++li) {
            MiserNode &nd = nodes[(size_t)level[li]];
            const int slot = (int)li;
            rfrom.insert(rfrom.end(), nd.from.begin(), nd.from.end());
            rto.insert(rto.end(), nd.to.begin(), nd.to.end());
            bool ok = true;
            for (int d = 0; d < n; ++d) ok = ok && uniform_ok(nd.from[d], nd.to[d]); // samplers, :320-324
            if (!ok) {
                nd.status = EQS_ERR_EXTERNAL;
                nd.why = "low > high (or non-finite) in uniform distribution";
                continue;
            }
            if (nd.depth_remaining == 0 || nd.count < min_count) { // plain Monte Carlo, :335-349
                nd.leaf = true;
                const long long cnt = std::max(nd.count, min_count);
                if (cnt < 2) {
                    nd.status = EQS_ERR_INCORRECT_INPUT;
                    nd.why = "the iterator must be over at least 2 elements"; // lib.rs:190-193
                    continue;
                }
                nd.first_batch = (int)batches.size();
                add_batch(slot, -1, 0, 0.0, cnt, leaf_base, nd.id);
                continue;
            }
            nd.count = std::max<long long>(nd.count, 2); // :351
            if (!uniform_ok(-p->dither, p->dither)) {   // dither_sampler, :361
                nd.status = EQS_ERR_EXTERNAL;
                nd.why = "low > high (or non-finite) in uniform distribution (dither)";
                continue;
            }
            nd.mid.resize((size_t)n);
            for (int d = 0; d < n && ok; ++d) {
                const uint4 w = philox_draw(key, STREAM_MISER, nd.id, ctl | (long long)(1 + d), 0u);
                const double r = host_u01(w.x, w.y) * (p->dither - (-p->dither)) + (-p->dither);
                const double delta = nd.to[d] - nd.from[d];
                const double mid = nd.from[d] + delta * (0.5 + r); // :364-365
                nd.mid[(size_t)d] = mid;
                ok = uniform_ok(nd.from[d], mid) && uniform_ok(mid, nd.to[d]); // :366-367
            }
            if (!ok) {
                nd.status = EQS_ERR_EXTERNAL;
                nd.why = "low > high (or non-finite) in uniform distribution (bisection)";
                continue;
            }
            nd.first_batch = (int)batches.size();
            for (int d = 0; d < n; ++d)
                for (int up = 0; up < 2; ++up)
                    add_batch(slot, d, up, nd.mid[(size_t)d], nd.count, (long long)(2 * d + up) << 40, nd.id);
        }

        // ---- one launch for every batch of the level ----------------------------------------------------------
        std::vector<Moments> out(batches.size());
        if (!batches.empty()) {
            const size_t b_bytes = batches.size() * sizeof(MiserBatch), r_bytes = rfrom.size() * 8;
            const size_t o_b = 0, o_rf = align_up(b_bytes, 256), o_rt = o_rf + align_up(r_bytes, 256);
            const size_t o_part = o_rt + align_up(r_bytes, 256), o_out = o_part + align_up((size_t)nchunks * sizeof(Moments), 256);
            const size_t total = o_out + align_up(batches.size() * sizeof(Moments), 256);
            void *arena = nullptr;
            size_t got = 0;
            EQS_CUDA(c, c->acquire_arena(total, &arena, &got));
            char *base = (char *)arena;
            auto try_cuda = [&](cudaError_t e) {
                if (e != cudaSuccess && rc == EQS_OK)
                    rc = c->fail(EQS_ERR_EXTERNAL, "CUDA error %s: %s", cudaGetErrorName(e), cudaGetErrorString(e));
            };
            try_cuda(cudaMemcpyAsync(base + o_b, batches.data(), b_bytes, cudaMemcpyHostToDevice, s));
            try_cuda(cudaMemcpyAsync(base + o_rf, rfrom.data(), r_bytes, cudaMemcpyHostToDevice, s));
            try_cuda(cudaMemcpyAsync(base + o_rt, rto.data(), r_bytes, cudaMemcpyHostToDevice, s));
            MiserDev D{};
            D.batch = (const MiserBatch *)(base + o_b);
            D.rfrom = (const double *)(base + o_rf);
            D.rto = (const double *)(base + o_rt);
            D.part = (Moments *)(base + o_part);
            D.out = (Moments *)(base + o_out);
            D.nchunks = nchunks;
            D.nbatch = (int)batches.size();
            D.n = n;
            D.key = key;
            try_cuda(cudaEventRecord(e0, s));
            const int grid = (int)std::min<long long>(nchunks, 16LL * c->sm_count);
            const size_t smem = (size_t)2 * n * sizeof(double);
            if (rc == EQS_OK)
                dispatch_objective(p->integrand_id, [&]<class Obj>() {
                    if constexpr (Obj::streaming) {
                        auto kernel = miser_chunk_kernel<Obj>;
                        if (smem > kSmemOptIn) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                        kernel<<<grid, MS_BLOCK, smem, s>>>(D);
                    }
                });
            if (rc == EQS_OK) miser_merge_kernel<<<(D.nbatch * 32 + MS_BLOCK - 1) / MS_BLOCK, MS_BLOCK, 0, s>>>(D);
            try_cuda(cudaGetLastError());
            try_cuda(cudaEventRecord(e1, s));
            try_cuda(cudaMemcpyAsync(out.data(), D.out, out.size() * sizeof(Moments), cudaMemcpyDeviceToHost, s));
            try_cuda(cudaStreamSynchronize(s));
            if (rc == EQS_OK) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, e0, e1);
                total_ms += ms;
                launches += 2;
            }
            c->release_arena(arena, got);
            if (rc != EQS_OK) break;
        }

        // ---- decide: leaves are done, inner nodes choose their bisection and hand points to two children ------
        std::vector<int> next;
        for (size_t li = 0; li < level.size(); ++li) {
            const int ni = level[li];
            if (nodes[(size_t)ni].status != EQS_OK) continue;
            if (nodes[(size_t)ni].leaf) {
                MiserNode &nd = nodes[(size_t)ni];
                const Moments &m = out[(size_t)nd.first_batch];
                double volume = 1.0;
                for (int d = 0; d < n; ++d) volume = volume * (nd.to[d] - nd.from[d]); // fold(T::one(), T::mul), :337-341
                nd.mean = m.mean * volume;                                             // scale_mean
                nd.variance = (m.m2 / m.n) * volume * volume;                          // no Bessel correction, :345
                continue;
            }
            const MiserNode cur = nodes[(size_t)ni]; // copy: nodes grows below
            double best_variance = std::numeric_limits<double>::max(), best_lower = best_variance, best_upper = best_variance; // :356-358
            const uint4 w = philox_draw(key, STREAM_MISER, cur.id, ctl, 0u);
            int best_d = std::min(n - 1, (int)(host_u01(w.x, w.y) * (double)n));       // uniform(0, n - 1), :359
            double best_mid = cur.from[best_d] + (cur.to[best_d] - cur.from[best_d]) * 0.5; // :360
            for (int d = 0; d < n; ++d) {
                const Moments &ml = out[(size_t)cur.first_batch + 2 * d], &mu = out[(size_t)cur.first_batch + 2 * d + 1];
                const double vl = ml.m2 / (ml.n - 1.0), vu = mu.m2 / (mu.n - 1.0);      // Bessel-corrected, :382-392
                const double tv = vl + vu;
                if (tv < best_variance) { // :396-402
                    best_variance = tv;
                    best_lower = vl;
                    best_upper = vu;
                    best_d = d;
                    best_mid = cur.mid[(size_t)d];
                }
            }
            const double beta = 1.0 / (1.0 + p->alpha);                                  // :407
            const double la = std::pow(best_lower, beta), ua = std::pow(best_upper, beta);
            const double sum = la + ua;
            const double nl = (double)cur.count * (la / sum), nu = (double)cur.count * (ua / sum); // :411-419
            auto as_count = [](double v, long long *o) { // Float::to_usize: NaN, <= -1 and >= 2^64 have no value
                if (!(v > -1.0 && v < 18446744073709551616.0)) return false;
                *o = v < 0.0 ? 0 : (long long)std::min(v, 9.0e18);
                return true;
            };
            long long cl = 0, cu = 0;
            if (!as_count(nl, &cl) || !as_count(nu, &cu)) {
                nodes[(size_t)ni].status = EQS_ERR_TYPE_CONVERSION;
                nodes[(size_t)ni].why = "a variance share of the points is not a count";
                continue;
            }
            for (int up = 0; up < 2; ++up) { // lower first, :422-440
                MiserNode ch;
                ch.from = cur.from;
                ch.to = cur.to;
                if (up) ch.from[(size_t)best_d] = best_mid;
                else ch.to[(size_t)best_d] = best_mid;
                ch.count = up ? cu : cl;
                ch.depth_remaining = cur.depth_remaining - 1;
                ch.id = 2 * cur.id + (uint32_t)up;
                nodes.push_back(std::move(ch));
                (up ? nodes[(size_t)ni].upper : nodes[(size_t)ni].lower) = (int)nodes.size() - 1;
                next.push_back((int)nodes.size() - 1);
            }
        }
        level.swap(next);
        if (nodes.size() > (size_t)1 << 22) { // the host holds the recursion tree; 4M regions is far beyond any sane call
            rc = c->fail(EQS_ERR_INCORRECT_INPUT, "MISER recursion spans more than 2^22 regions; lower maximum_depth");
            break;
        }
    }
    if (rc != EQS_OK) return rc;

    // ---- results in the recursion's own order: a node's error first, then lower, then upper (`?`, :422-440) ------
    struct Ret {
        int status;
        const char *why;
        double mean, variance;
    };
    std::function<Ret(int)> collect = [&](int ni) -> Ret {
        const MiserNode &nd = nodes[(size_t)ni];
        if (nd.status != EQS_OK) return Ret{nd.status, nd.why, 0.0, 0.0};
        if (nd.leaf) return Ret{EQS_OK, "", nd.mean, nd.variance};
        const Ret lo = collect(nd.lower);
        if (lo.status != EQS_OK) return lo;
        const Ret up = collect(nd.upper);
        if (up.status != EQS_OK) return up;
        return Ret{EQS_OK, "", lo.mean + up.mean, lo.variance + up.variance}; // MeanVariance::add, :442
    };
    const Ret r = collect(0);
    if (r.status != EQS_OK) return c->fail(r.status, "%s", r.why);
    *mean = r.mean;
    *variance = r.variance;
    if (report) {
        *report = eqs_report{};
        report->iters_run = levels;
        report->evals = evals;
        report->best_cost = r.mean;
        report->world = 1;
        report->kernel_launches = launches;
        report->loop_ms = total_ms;
    }
    return EQS_OK;
    EQS_BOUNDARY_CATCH
}

} // extern "C"
