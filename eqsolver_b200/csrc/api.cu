// api.cu -- the extern "C" boundary (include/eqsolver_b200.h): context, diagnostics, ParticleSwarm entry points.
// CrossEntropy entry points live in ce.cu.  No exception crosses the boundary.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <exception>
#include <new>

#define EQS_DEVMATH_TABLES // draws_kernel samples normals: log_unit_tab's shared-memory tables (devmath.cuh)
#include "common.cuh"
#include "objectives.cuh"
#include "philox.cuh"
#include "pso.cuh"

using namespace eqs;

struct eqs_pso {
    PsoSolver s;
    explicit eqs_pso(Ctx *c) : s(c) {}
};

static thread_local std::string g_create_error;
static thread_local char g_boundary_text[256] = "";

namespace eqs {
int boundary_exception() noexcept
{
    try {
        throw;
    } catch (const std::bad_alloc &) {
        snprintf(g_boundary_text, sizeof(g_boundary_text), "out of host memory");
    } catch (const std::exception &e) {
        snprintf(g_boundary_text, sizeof(g_boundary_text), "C++ exception: %s", e.what());
    } catch (...) {
        snprintf(g_boundary_text, sizeof(g_boundary_text), "unknown C++ exception");
    }
    return EQS_ERR_EXTERNAL;
}
const char *boundary_exception_text() noexcept { return g_boundary_text; }
void boundary_exception_clear() noexcept { g_boundary_text[0] = 0; }
} // namespace eqs


namespace {

constexpr int64_t kDiagMaxCount = (int64_t)1 << 36; // values per call of a diagnostic entry point (grid and size_t arithmetic stay in range)

template <class Obj> __global__ void objective_eval_kernel(const double *x, int n, long long count, double *out)
{
    // x is particle-major [count][n] as the caller (and the reference's closure) sees it
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = objective_on<Obj>(StridedView{x + i * n, 1}, n);
}

__global__ void philox_blocks_kernel(const uint32_t *ctr, PhiloxKey key, long long count, uint32_t *out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint4 r = philox4x32_10(ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3], key);
    out[4 * i] = r.x;
    out[4 * i + 1] = r.y;
    out[4 * i + 2] = r.z;
    out[4 * i + 3] = r.w;
}

__global__ void cos_kernel(const double *x, long long count, double *out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = eqs_cos(x[i]);
}

__global__ void draws_kernel(int kind, PhiloxKey key, uint32_t t, long long first, long long count, int n, double *a, double *b)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (kind == 2) devmath_tables_load(); // every thread of the block: it ends with a barrier
    if (i >= count) return;
    const long long gi = first + i;
    if (kind == 2) {
        for (int p = 0; 2 * p < n; ++p) {
            double z0, z1;
            normal_pair(philox_draw(key, STREAM_CE, t, gi, (uint32_t)p), z0, z1);
            a[i * n + 2 * p] = z0;
            if (2 * p + 1 < n) a[i * n + 2 * p + 1] = z1;
        }
        return;
    }
    for (int d = 0; d < n; ++d) {
        const uint4 w = philox_draw(key, kind == 0 ? STREAM_PSO_INIT : STREAM_PSO_STEP, t, gi, (uint32_t)d);
        a[i * n + d] = u01(w.x, w.y);
        b[i * n + d] = u01(w.z, w.w);
    }
}

// DFMA issue-rate probe: 8 independent dependent-chains per thread, fully unrolled
__global__ void __launch_bounds__(256) dfma_kernel(double *out, int iters, double a, double b)
{
    double r0 = threadIdx.x, r1 = r0 + 1, r2 = r0 + 2, r3 = r0 + 3, r4 = r0 + 4, r5 = r0 + 5, r6 = r0 + 6, r7 = r0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            r0 = fma(r0, a, b); r1 = fma(r1, a, b); r2 = fma(r2, a, b); r3 = fma(r3, a, b);
            r4 = fma(r4, a, b); r5 = fma(r5, a, b); r6 = fma(r6, a, b); r7 = fma(r7, a, b);
        }
    }
    const double s = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
    if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) copy_kernel(const double2 *__restrict__ src, double2 *__restrict__ dst, long long n2)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

// the ParticleSwarm step's memory traffic with the arithmetic removed: a thread walks the 4-dimension groups of its
// particle in the [groups][ld][4] layout, 256-bit loads of three buffers, 256-bit stores back into two of them
// (in place, like x and v), two groups in flight -- the ceiling this access pattern can reach on the machine
__device__ __forceinline__ void ld4(const double *p, double (&r)[4])
{
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r[0]), "=d"(r[1]), "=d"(r[2]), "=d"(r[3]) : "l"(p));
}
__device__ __forceinline__ void st4(double *p, const double (&r)[4])
{
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(r[0]), "d"(r[1]), "d"(r[2]), "d"(r[3]));
}
__global__ void __launch_bounds__(256, 2) pattern_kernel(double *a, double *b, const double *c, long long ld, int groups)
{
    for (long long li = (long long)blockIdx.x * blockDim.x + threadIdx.x; li < ld; li += (long long)gridDim.x * blockDim.x) {
        for (int q = 0; q + 2 <= groups; q += 2) {
            double x[2][4], v[2][4], p[2][4];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const long long off = ((long long)(q + k) * ld + li) * 4;
                ld4(a + off, x[k]);
                ld4(b + off, v[k]);
                ld4(c + off, p[k]);
            }
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const long long off = ((long long)(q + k) * ld + li) * 4;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    v[k][j] = v[k][j] + p[k][j];
                    x[k][j] = x[k][j] + v[k][j];
                }
                st4(a + off, x[k]);
                st4(b + off, v[k]);
            }
        }
    }
}

} // namespace

extern "C" {

int eqs_abi_version(void) { return EQS_ABI_VERSION; }

int eqs_ctx_create(int device, eqs_ctx **out)
{
    EQS_BOUNDARY_TRY
    if (!out) return EQS_ERR_INCORRECT_INPUT;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        // no CPU fallback: fail loudly
        g_create_error = std::string("no CUDA device available (") + (e == cudaSuccess ? "device count 0" : cudaGetErrorString(e)) +
                         "); eqsolver_b200 has no CPU path";
        return EQS_ERR_EXTERNAL;
    }
    if (device < 0 || device >= count) {
        g_create_error = "device ordinal out of range";
        return EQS_ERR_INCORRECT_INPUT;
    }
    eqs_ctx *ctx = new (std::nothrow) eqs_ctx;
    if (!ctx) return EQS_ERR_EXTERNAL;
    ctx->c.device = device;
    if ((e = cudaSetDevice(device)) != cudaSuccess ||
        (e = cudaDeviceGetAttribute(&ctx->c.sm_count, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&ctx->c.stream, cudaStreamNonBlocking)) != cudaSuccess) {
        g_create_error = std::string("CUDA error: ") + cudaGetErrorString(e);
        delete ctx;
        return EQS_ERR_EXTERNAL;
    }
    *out = ctx;
    return EQS_OK;
    EQS_BOUNDARY_CATCH
}

int eqs_ctx_destroy(eqs_ctx *ctx)
{
    EQS_BOUNDARY_TRY
    if (!ctx) return EQS_OK;
    cudaSetDevice(ctx->c.device);
    comm_destroy(&ctx->c);
    if (ctx->c.stream) cudaStreamSynchronize(ctx->c.stream);
    ctx->c.drop_caches();
    if (ctx->c.stream) cudaStreamDestroy(ctx->c.stream);
    delete ctx;
    return EQS_OK;
    EQS_BOUNDARY_CATCH
}

const char *eqs_last_error(const eqs_ctx *ctx)
{
    // a C++ exception caught at the boundary (out of host memory ...) never reached Ctx::fail: its text is per thread
    if (boundary_exception_text()[0]) return boundary_exception_text();
    if (!ctx) return g_create_error.c_str();
    // hand out a per-thread copy: another thread may fail on the same context while the caller reads the message
    static thread_local std::string copy;
    try {
        std::lock_guard<std::mutex> lock(ctx->c.error_mutex);
        copy = ctx->c.error;
    } catch (...) {
        return "out of host memory";
    }
    return copy.c_str();
}

int eqs_ctx_sm_count(const eqs_ctx *ctx, int *out)
{
    EQS_BOUNDARY_TRY
    if (!ctx || !out) return EQS_ERR_INCORRECT_INPUT;
    *out = ctx->c.sm_count;
    return EQS_OK;
    EQS_BOUNDARY_CATCH
}

int eqs_comm_unique_id(void *out128)
{
    EQS_BOUNDARY_TRY
    if (!out128) return EQS_ERR_INCORRECT_INPUT;
    return comm_unique_id(out128, g_create_error);
    EQS_BOUNDARY_CATCH
}

int eqs_ctx_comm_init(eqs_ctx *ctx, int rank, int world, const void *id128)
{
    EQS_BOUNDARY_TRY
    if (!ctx || (world > 1 && !id128)) return EQS_ERR_INCORRECT_INPUT;
    return comm_init(&ctx->c, rank, world, id128);
    EQS_BOUNDARY_CATCH
}

int eqs_ctx_rank(const eqs_ctx *ctx, int *rank, int *world)
{
    EQS_BOUNDARY_TRY
    if (!ctx) return EQS_ERR_INCORRECT_INPUT;
    if (rank) *rank = ctx->c.rank;
    if (world) *world = ctx->c.world;
    return EQS_OK;
    EQS_BOUNDARY_CATCH
}

int eqs_shard_range(int64_t count, int world, int rank, int64_t *first, int64_t *local)
{
    EQS_BOUNDARY_TRY
    if (count < 0 || world < 1 || rank < 0 || rank >= world) return EQS_ERR_INCORRECT_INPUT;
    const int64_t base = count / world, rem = count % world;
    const int64_t f = rank * base + std::min<int64_t>(rank, rem);
    if (first) *first = f;
    if (local) *local = base + (rank < rem ? 1 : 0);
    return EQS_OK;
    EQS_BOUNDARY_CATCH
}

int64_t eqs_record_doubles(int32_t n) { return record_doubles(n); }

// ---- diagnostics ---------------------------------------------------------------------------------------

int eqs_objective_eval(eqs_ctx *ctx, int32_t objective_id, const double *x, int32_t n, int64_t count, double *out)
{
    EQS_BOUNDARY_TRY
    if (!ctx || !x || !out || n <= 0 || count < 0) return EQS_ERR_INCORRECT_INPUT;
    Ctx *c = &ctx->c;
    if ((objective_id == EQS_OBJ_THREE_CIRCLES || objective_id == EQS_OBJ_SINCOS) && n != 2)
        return c->fail(EQS_ERR_INCORRECT_INPUT, "this objective needs n = 2");
    if (count == 0) return EQS_OK;
    EQS_CUDA(c, cudaSetDevice(c->device));
    if (count > kDiagMaxCount / n) return c->fail(EQS_ERR_INCORRECT_INPUT, "count * n is limited to 2^36 values");
    DevBuf<double> dx, dout;
    EQS_CUDA(c, dx.alloc((size_t)n * count));
    EQS_CUDA(c, dout.alloc((size_t)count));
    EQS_CUDA(c, cudaMemcpyAsync(dx, x, sizeof(double) * n * count, cudaMemcpyHostToDevice, c->stream));
    double *const px = dx, *const po = dout;
    const bool known = dispatch_objective(objective_id, [&]<class Obj>() {
        objective_eval_kernel<Obj><<<(unsigned)((count + 127) / 128), 128, 0, c->stream>>>(px, n, count, po);
    });
    if (!known) return c->fail(EQS_ERR_INCORRECT_INPUT, "unknown objective id %d", objective_id);
    EQS_CUDA(c, cudaMemcpyAsync(out, dout, sizeof(double) * count, cudaMemcpyDeviceToHost, c->stream));
    EQS_CUDA(c, cudaStreamSynchronize(c->stream));
    return EQS_OK;
    EQS_BOUNDARY_CATCH
}

int eqs_philox_blocks(eqs_ctx *ctx, const uint32_t *ctr, const uint32_t key[2], int64_t count, uint32_t *out)
{
    EQS_BOUNDARY_TRY
    if (!ctx || !ctr || !key || !out || count < 0) return EQS_ERR_INCORRECT_INPUT;
    Ctx *c = &ctx->c;
    if (count == 0) return EQS_OK;
    EQS_CUDA(c, cudaSetDevice(c->device));
    if (count > kDiagMaxCount / 4) return c->fail(EQS_ERR_INCORRECT_INPUT, "count is limited to 2^34 blocks");
    DevBuf<uint32_t> dc, dout;
    EQS_CUDA(c, dc.alloc(4 * (size_t)count));
    EQS_CUDA(c, dout.alloc(4 * (size_t)count));
    EQS_CUDA(c, cudaMemcpyAsync(dc, ctr, 16 * count, cudaMemcpyHostToDevice, c->stream));
    philox_blocks_kernel<<<(unsigned)((count + 127) / 128), 128, 0, c->stream>>>(dc, philox_key(key[0], key[1]), count, dout);
    EQS_CUDA(c, cudaGetLastError());
    EQS_CUDA(c, cudaMemcpyAsync(out, dout, 16 * count, cudaMemcpyDeviceToHost, c->stream));
    EQS_CUDA(c, cudaStreamSynchronize(c->stream));
    return EQS_OK;
    EQS_BOUNDARY_CATCH
}

int eqs_draws(eqs_ctx *ctx, int32_t kind, uint64_t seed, int64_t t, int64_t first, int64_t count, int32_t n, double *a, double *b)
{
    EQS_BOUNDARY_TRY
    if (!ctx || !a || kind < 0 || kind > 2 || (kind != 2 && !b) || n <= 0 || count < 0) return EQS_ERR_INCORRECT_INPUT;
    Ctx *c = &ctx->c;
    if (count == 0) return EQS_OK;
    EQS_CUDA(c, cudaSetDevice(c->device));
    if (count > kDiagMaxCount / n) return c->fail(EQS_ERR_INCORRECT_INPUT, "count * n is limited to 2^36 values");
    DevBuf<double> da, db;
    const size_t bytes = sizeof(double) * n * count;
    EQS_CUDA(c, da.alloc((size_t)n * count));
    EQS_CUDA(c, db.alloc((size_t)n * count));
    draws_kernel<<<(unsigned)((count + 127) / 128), 128, 0, c->stream>>>(kind, philox_key(seed), (uint32_t)t, first, count, n, da, db);
    EQS_CUDA(c, cudaGetLastError());
    EQS_CUDA(c, cudaMemcpyAsync(a, da, bytes, cudaMemcpyDeviceToHost, c->stream));
    if (kind != 2) EQS_CUDA(c, cudaMemcpyAsync(b, db, bytes, cudaMemcpyDeviceToHost, c->stream));
    EQS_CUDA(c, cudaStreamSynchronize(c->stream));
    return EQS_OK;
    EQS_BOUNDARY_CATCH
}

int eqs_device_cos(eqs_ctx *ctx, const double *x, int64_t count, double *out)
{
    EQS_BOUNDARY_TRY
    if (!ctx || !x || !out || count < 0) return EQS_ERR_INCORRECT_INPUT;
    Ctx *c = &ctx->c;
    if (count == 0) return EQS_OK;
    EQS_CUDA(c, cudaSetDevice(c->device));
    if (count > kDiagMaxCount) return c->fail(EQS_ERR_INCORRECT_INPUT, "count is limited to 2^36 values");
    DevBuf<double> dx, dout;
    EQS_CUDA(c, dx.alloc((size_t)count));
    EQS_CUDA(c, dout.alloc((size_t)count));
    EQS_CUDA(c, cudaMemcpyAsync(dx, x, sizeof(double) * count, cudaMemcpyHostToDevice, c->stream));
    cos_kernel<<<(unsigned)((count + 255) / 256), 256, 0, c->stream>>>(dx, count, dout);
    EQS_CUDA(c, cudaGetLastError());
    EQS_CUDA(c, cudaMemcpyAsync(out, dout, sizeof(double) * count, cudaMemcpyDeviceToHost, c->stream));
    EQS_CUDA(c, cudaStreamSynchronize(c->stream));
    return EQS_OK;
    EQS_BOUNDARY_CATCH
}

int eqs_measure_fp64(eqs_ctx *ctx, double *tflops)
{
    EQS_BOUNDARY_TRY
    if (!ctx || !tflops) return EQS_ERR_INCORRECT_INPUT;
    Ctx *c = &ctx->c;
    EQS_CUDA(c, cudaSetDevice(c->device));
    DevBuf<double> dout;
    EQS_CUDA(c, dout.alloc(1));
    Event e0, e1;
    EQS_CUDA(c, e0.create());
    EQS_CUDA(c, e1.create());
    const int grid = c->sm_count * 8, iters = 2048;
    dfma_kernel<<<grid, 256, 0, c->stream>>>(dout, 64, 1.0000001, 1e-9); // warm-up
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0, c->stream);
        dfma_kernel<<<grid, 256, 0, c->stream>>>(dout, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1, c->stream);
        EQS_CUDA(c, cudaEventSynchronize(e1));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 8 * 16 * (double)iters * 256.0 * grid;
        best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    *tflops = best;
    return EQS_OK;
    EQS_BOUNDARY_CATCH
}

int eqs_measure_copy(eqs_ctx *ctx, int64_t bytes, double *gbps)
{
    EQS_BOUNDARY_TRY
    if (!ctx || !gbps || bytes < 16) return EQS_ERR_INCORRECT_INPUT;
    Ctx *c = &ctx->c;
    EQS_CUDA(c, cudaSetDevice(c->device));
    const long long n2 = bytes / 16;
    DevBuf<double2> a, b;
    EQS_CUDA(c, a.alloc((size_t)n2));
    EQS_CUDA(c, b.alloc((size_t)n2));
    EQS_CUDA(c, cudaMemsetAsync(a, 0, n2 * 16, c->stream));
    Event e0, e1;
    EQS_CUDA(c, e0.create());
    EQS_CUDA(c, e1.create());
    const int grid = c->sm_count * 8;
    copy_kernel<<<grid, 256, 0, c->stream>>>(a, b, n2);
    double best = 0.0;
    for (int rep = 0; rep < 10; ++rep) {
        cudaEventRecord(e0, c->stream);
        copy_kernel<<<grid, 256, 0, c->stream>>>(a, b, n2);
        cudaEventRecord(e1, c->stream);
        EQS_CUDA(c, cudaEventSynchronize(e1));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        best = std::max(best, 2.0 * n2 * 16 / (ms * 1e-3) / 1e9);
    }
    *gbps = best;
    return EQS_OK;
    EQS_BOUNDARY_CATCH
}

int eqs_measure_pattern(eqs_ctx *ctx, int32_t n, int64_t particles, double *gbps)
{
    EQS_BOUNDARY_TRY
    if (!ctx || !gbps || n < 8 || (n & 7) || particles < 1) return EQS_ERR_INCORRECT_INPUT;
    Ctx *c = &ctx->c;
    EQS_CUDA(c, cudaSetDevice(c->device));
    const long long ld = (particles + 31) / 32 * 32;
    const int groups = n / 4;
    const size_t bytes = (size_t)groups * ld * 4 * sizeof(double);
    DevBuf<double> buf[3];
    for (auto &p : buf) {
        EQS_CUDA(c, p.alloc(bytes / sizeof(double)));
        EQS_CUDA(c, cudaMemsetAsync(p, 0, bytes, c->stream));
    }
    Event e0, e1;
    EQS_CUDA(c, e0.create());
    EQS_CUDA(c, e1.create());
    const int grid = (int)std::min<long long>(2LL * c->sm_count, (ld + 255) / 256);
    double best = 0.0;
    for (int rep = 0; rep < 12; ++rep) { // rep 0 warms up
        cudaEventRecord(e0, c->stream);
        pattern_kernel<<<grid, 256, 0, c->stream>>>(buf[0], buf[1], buf[2], ld, groups);
        cudaEventRecord(e1, c->stream);
        EQS_CUDA(c, cudaEventSynchronize(e1));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0) best = std::max(best, 5.0 * (double)bytes / (ms * 1e-3) / 1e9);
    }
    // EQS_PATTERN_SECONDS=t: the SUSTAINED figure instead of the burst -- launches back to back for about t seconds
    // between two events (what the board delivers for this traffic once the power cap has settled the clocks)
    if (const char *e = getenv("EQS_PATTERN_SECONDS")) {
        const double seconds = atof(e);
        if (seconds > 0.0) {
            const long long launches = std::max<long long>(8, (long long)(seconds * best * 1e9 / (5.0 * (double)bytes)));
            cudaEventRecord(e0, c->stream);
            for (long long k = 0; k < launches; ++k) pattern_kernel<<<grid, 256, 0, c->stream>>>(buf[0], buf[1], buf[2], ld, groups);
            cudaEventRecord(e1, c->stream);
            EQS_CUDA(c, cudaEventSynchronize(e1));
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            best = 5.0 * (double)bytes * (double)launches / (ms * 1e-3) / 1e9;
        }
    }
    *gbps = best;
    return EQS_OK;
    EQS_BOUNDARY_CATCH
}

// ---- ParticleSwarm -------------------------------------------------------------------------------------

int eqs_pso_validate(const eqs_pso_params *p) { return validate_pso_params(p, nullptr); }

int eqs_pso_create(eqs_ctx *ctx, const eqs_pso_params *p, eqs_pso **out)
{
    EQS_BOUNDARY_TRY
    if (!ctx || !out) return EQS_ERR_INCORRECT_INPUT;
    *out = nullptr;
    eqs_pso *h = new (std::nothrow) eqs_pso(&ctx->c);
    if (!h) return ctx->c.fail(EQS_ERR_EXTERNAL, "out of host memory");
    const int st = h->s.create(p);
    if (st != EQS_OK) {
        delete h;
        return st;
    }
    *out = h;
    return EQS_OK;
    EQS_BOUNDARY_CATCH
}

int eqs_pso_destroy(eqs_pso *h)
{
    EQS_BOUNDARY_TRY
    delete h;
    return EQS_OK;
    EQS_BOUNDARY_CATCH
}

#define H_OR_FAIL(h)                                                                                             \
    if (!(h)) return EQS_ERR_INCORRECT_INPUT

int eqs_pso_init(eqs_pso *h, const double *x0)
{
    EQS_BOUNDARY_TRY
    H_OR_FAIL(h);
    if (!x0) return h->s.ctx_->fail(EQS_ERR_INCORRECT_INPUT, "null x0");
    return h->s.init(x0);
    EQS_BOUNDARY_CATCH
}

int eqs_pso_step(eqs_pso *h, int64_t iters, const double *replay_step)
{
    EQS_BOUNDARY_TRY
    H_OR_FAIL(h);
    if (iters < 0) return h->s.ctx_->fail(EQS_ERR_INCORRECT_INPUT, "negative iteration count");
    return h->s.step(iters, replay_step);
    EQS_BOUNDARY_CATCH
}

int eqs_pso_sync(eqs_pso *h)
{
    EQS_BOUNDARY_TRY
    H_OR_FAIL(h);
    return h->s.sync();
    EQS_BOUNDARY_CATCH
}

int eqs_pso_last_step_ms(eqs_pso *h, double *ms, int64_t *kernel_launches)
{
    EQS_BOUNDARY_TRY
    H_OR_FAIL(h);
    Ctx *c = h->s.ctx_;
    EQS_CUDA(c, cudaEventSynchronize(h->s.ev1_));
    float f = 0.f;
    EQS_CUDA(c, cudaEventElapsedTime(&f, h->s.ev0_, h->s.ev1_));
    if (ms) *ms = f;
    if (kernel_launches) *kernel_launches = h->s.last_launches_;
    return EQS_OK;
    EQS_BOUNDARY_CATCH
}

int eqs_pso_read_gbest(eqs_pso *h, double *g, double *cost, int64_t *iters_run, int32_t *stalled, int64_t *improvements)
{
    EQS_BOUNDARY_TRY
    H_OR_FAIL(h);
    return h->s.read_gbest(g, cost, iters_run, stalled, improvements);
    EQS_BOUNDARY_CATCH
}

int eqs_pso_read_state(eqs_pso *h, double *x, double *v, double *pbest, double *pbest_cost)
{
    EQS_BOUNDARY_TRY
    H_OR_FAIL(h);
    return h->s.read_state(x, v, pbest, pbest_cost);
    EQS_BOUNDARY_CATCH
}

int eqs_pso_read_ring(eqs_pso *h, double *ring, int64_t *ring_i)
{
    EQS_BOUNDARY_TRY
    H_OR_FAIL(h);
    return h->s.read_ring(ring, ring_i);
    EQS_BOUNDARY_CATCH
}

int eqs_pso_local_range(eqs_pso *h, int64_t *first, int64_t *local)
{
    EQS_BOUNDARY_TRY
    H_OR_FAIL(h);
    if (first) *first = h->s.d_.first;
    if (local) *local = h->s.d_.nloc;
    return EQS_OK;
    EQS_BOUNDARY_CATCH
}

int eqs_pso_local_indices(eqs_pso *h, int64_t *gidx)
{
    EQS_BOUNDARY_TRY
    H_OR_FAIL(h);
    if (!gidx) return EQS_ERR_INCORRECT_INPUT;
    return h->s.local_indices(gidx);
    EQS_BOUNDARY_CATCH
}

int eqs_pso_init_local(eqs_pso *h, const double *x0, int32_t phase)
{
    EQS_BOUNDARY_TRY
    H_OR_FAIL(h);
    if (phase < 0 || phase > 1 || (phase == 0 && !x0)) return h->s.ctx_->fail(EQS_ERR_INCORRECT_INPUT, "bad init phase");
    return h->s.init_local(x0, phase);
    EQS_BOUNDARY_CATCH
}

int eqs_pso_init_apply(eqs_pso *h)
{
    EQS_BOUNDARY_TRY
    H_OR_FAIL(h);
    return h->s.init_apply();
    EQS_BOUNDARY_CATCH
}

int eqs_pso_step_local(eqs_pso *h, const double *replay_step)
{
    EQS_BOUNDARY_TRY
    H_OR_FAIL(h);
    return h->s.step_local(replay_step);
    EQS_BOUNDARY_CATCH
}

int eqs_pso_step_apply(eqs_pso *h)
{
    EQS_BOUNDARY_TRY
    H_OR_FAIL(h);
    return h->s.step_apply();
    EQS_BOUNDARY_CATCH
}

int eqs_pso_get_record(eqs_pso *h, double *record)
{
    EQS_BOUNDARY_TRY
    H_OR_FAIL(h);
    if (!record) return EQS_ERR_INCORRECT_INPUT;
    return h->s.get_record(record);
    EQS_BOUNDARY_CATCH
}

int eqs_pso_set_records(eqs_pso *h, const double *records)
{
    EQS_BOUNDARY_TRY
    H_OR_FAIL(h);
    if (!records) return EQS_ERR_INCORRECT_INPUT;
    return h->s.set_records(records);
    EQS_BOUNDARY_CATCH
}

int eqs_pso_solve(eqs_ctx *ctx, const eqs_pso_params *p, const double *x0, double *out_x, eqs_report *report)
{
    EQS_BOUNDARY_TRY
    if (!ctx) return EQS_ERR_INCORRECT_INPUT;
    if (!x0 || !out_x) return ctx->c.fail(EQS_ERR_INCORRECT_INPUT, "null x0 / out_x");
    PsoSolver s(&ctx->c);
    EQS_TRY(s.create(p));
    return s.solve(x0, out_x, report);
    EQS_BOUNDARY_CATCH
}

} // extern "C"
