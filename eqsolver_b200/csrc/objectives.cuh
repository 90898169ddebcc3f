// objectives.cuh -- ahead-of-time compiled objective functors, selected by eqs_objective ID.
//
// Replaces the user closure `F: Fn(VectorType<T,D>) -> T` of the reference
// (particle_swarm.rs:92,415; cross_entropy.rs:55,273), which cannot run on the device.
//
// A functor is either STREAMING -- begin / feed(x_d, d) for d = 0..n-1 in order / end -- so that the fused
// step kernel evaluates it while the coordinates are still in registers, or RANDOM-ACCESS -- eval(x, n) with
// x[d] reading coordinate d of this particle back from memory through a view of the device layout.  The operation
// order of every functor equals the oracle's (oracle/eqs_oracle.c: orc_objective); the library is compiled
// with -fmad=false because rustc never contracts a*b+c.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/eqsolver_b200.h"
#include "devmath.cuh"

// Where eqs_cos takes its coefficients from (devmath.cuh): 0 = literals (pso.cu, api.cu), 1 = __constant__ operands
// (ce.cu, mc.cu define it before including this header).  The functors live in an inline namespace named after the
// choice, so the two instantiations are distinct entities.
#ifndef EQS_COS_TABLE
#define EQS_COS_TABLE 0
#endif

namespace eqs {
#if EQS_COS_TABLE
inline namespace cos_table {
#else
inline namespace cos_literal {
#endif

__device__ __forceinline__ double obj_cos(double y) { return eqs_cos_t<(EQS_COS_TABLE != 0)>(y); }
template <int M> __device__ __forceinline__ void obj_cos_vec(const double (&y)[M], double (&out)[M])
{
    eqs_cos_vec<(EQS_COS_TABLE != 0), M>(y, out);
}
// the range test deferred to the end of the particle (devmath.cuh: DEFER)
template <int M> __device__ __forceinline__ void obj_cos_vec_defer(const double (&y)[M], double (&out)[M], unsigned &hi_max)
{
    eqs_cos_vec<(EQS_COS_TABLE != 0), M, true>(y, out, &hi_max);
}

struct ObjAcc {
    double a, b, p;
};

// views of one particle's coordinates in the device layouts
struct StridedView { // coordinate d at base[d * stride]: contiguous vectors (stride 1), dimension-major columns
    const double *base;
    long long stride;
    __device__ __forceinline__ double operator[](int d) const { return base[(long long)d * stride]; }
};
struct GroupedView { // ParticleSwarm population: [n/4][ld][4], base = buffer + 4 * particle, pitch = 4 * ld
    const double *base;
    long long pitch;
    __device__ __forceinline__ double operator[](int d) const { return base[(long long)(d >> 2) * pitch + (d & 3)]; }
};

constexpr double kTwoPi = 2.0 * 3.14159265358979323846; // (2.0 * PI), exact doubling
constexpr double kE = 2.718281828459045235360287;

template <int ID> struct Objective;

template <> struct Objective<EQS_OBJ_SPHERE> {
    static constexpr bool streaming = true;
    __device__ __forceinline__ static void begin(ObjAcc &s, int) { s.a = 0.0; }
    __device__ __forceinline__ static void feed(ObjAcc &s, double x, int) { s.a += x * x; }
    __device__ __forceinline__ static double end(const ObjAcc &s, int) { return s.a; }
};

template <> struct Objective<EQS_OBJ_ROSENBROCK> {
    static constexpr bool streaming = true;
    __device__ __forceinline__ static void begin(ObjAcc &s, int) { s.a = 0.0; s.p = 0.0; }
    __device__ __forceinline__ static void feed(ObjAcc &s, double x, int d)
    {
        if (d > 0) {
            const double t = x - s.p * s.p;
            const double u = 1.0 - s.p;
            s.a += 100.0 * (t * t) + u * u;
        }
        s.p = x;
    }
    __device__ __forceinline__ static double end(const ObjAcc &s, int) { return s.a; }
};

// examples/global_optimisation.rs:9-17
template <> struct Objective<EQS_OBJ_RASTRIGIN> {
    static constexpr bool streaming = true;
    __device__ __forceinline__ static void begin(ObjAcc &s, int n) { s.a = 10.0 * (double)n; }
    __device__ __forceinline__ static void feed(ObjAcc &s, double x, int) { s.a += x * x - 10.0 * obj_cos(kTwoPi * x); }
    // M consecutive coordinates at once: same operations per coordinate, accumulated in coordinate order
    template <int M> __device__ __forceinline__ static void feed_vec(ObjAcc &s, const double (&x)[M], int)
    {
        double y[M], c[M];
#pragma unroll
        for (int j = 0; j < M; ++j) y[j] = kTwoPi * x[j];
        obj_cos_vec<M>(y, c);
#pragma unroll
        for (int j = 0; j < M; ++j) s.a += x[j] * x[j] - 10.0 * c[j];
    }
    // same, the cosine's range test deferred (hi_max: devmath.cuh, eqs_cos_vec DEFER)
    template <int M> __device__ __forceinline__ static void feed_vec_defer(ObjAcc &s, const double (&x)[M], int, unsigned &hi_max)
    {
        double y[M], c[M];
#pragma unroll
        for (int j = 0; j < M; ++j) y[j] = kTwoPi * x[j];
        obj_cos_vec_defer<M>(y, c, hi_max);
#pragma unroll
        for (int j = 0; j < M; ++j) s.a += x[j] * x[j] - 10.0 * c[j];
    }
    __device__ __forceinline__ static double end(const ObjAcc &s, int) { return s.a; }
};

template <> struct Objective<EQS_OBJ_ACKLEY> {
    static constexpr bool streaming = true;
    __device__ __forceinline__ static void begin(ObjAcc &s, int) { s.a = 0.0; s.b = 0.0; }
    __device__ __forceinline__ static void feed(ObjAcc &s, double x, int)
    {
        s.a += x * x;
        s.b += obj_cos(kTwoPi * x);
    }
    template <int M> __device__ __forceinline__ static void feed_vec(ObjAcc &s, const double (&x)[M], int)
    {
        double y[M], c[M];
#pragma unroll
        for (int j = 0; j < M; ++j) y[j] = kTwoPi * x[j];
        obj_cos_vec<M>(y, c);
#pragma unroll
        for (int j = 0; j < M; ++j) {
            s.a += x[j] * x[j];
            s.b += c[j];
        }
    }
    template <int M> __device__ __forceinline__ static void feed_vec_defer(ObjAcc &s, const double (&x)[M], int, unsigned &hi_max)
    {
        double y[M], c[M];
#pragma unroll
        for (int j = 0; j < M; ++j) y[j] = kTwoPi * x[j];
        obj_cos_vec_defer<M>(y, c, hi_max);
#pragma unroll
        for (int j = 0; j < M; ++j) {
            s.a += x[j] * x[j];
            s.b += c[j];
        }
    }
    __device__ __forceinline__ static double end(const ObjAcc &s, int n)
    {
        const double a = -20.0 * exp(-0.2 * sqrt(s.a / (double)n));
        const double b = exp(s.b / (double)n);
        return a - b + 20.0 + kE;
    }
};

// examples/global_optimiser_as_solver.rs:22-42: || F(v) - b ||, circles (3,5,r3) (1,0,r4) (6,2,r2); n = 2
template <> struct Objective<EQS_OBJ_THREE_CIRCLES> {
    static constexpr bool streaming = true;
    __device__ __forceinline__ static void begin(ObjAcc &s, int) { s.a = 0.0; s.p = 0.0; }
    __device__ __forceinline__ static void feed(ObjAcc &s, double x, int d)
    {
        if (d == 0) s.p = x;
        if (d == 1) {
            const double x0 = s.p, x1 = x;
            const double e0 = ((x0 - 3.0) * (x0 - 3.0) + (x1 - 5.0) * (x1 - 5.0)) - 9.0;
            const double e1 = ((x0 - 1.0) * (x0 - 1.0) + (x1 - 0.0) * (x1 - 0.0)) - 16.0;
            const double e2 = ((x0 - 6.0) * (x0 - 6.0) + (x1 - 2.0) * (x1 - 2.0)) - 4.0;
            s.a = ((0.0 + e0 * e0) + e1 * e1) + e2 * e2;
        }
    }
    __device__ __forceinline__ static double end(const ObjAcc &s, int) { return sqrt(s.a); }
};

// benchmarks/multi.rs:106-115,137: out_i = sum_{j>=i} x_j^2 (a fresh forward sum per i), cost = ||out||.
// O(n^2) by construction in the reference; kept so, for bit parity of each partial sum.
template <> struct Objective<EQS_OBJ_HEAVY> {
    static constexpr bool streaming = false;
    // U of the forward sums run side by side: each is still 0 + x_i^2 + x_{i+1}^2 + .. in that order (same bits), but a
    // thread has U independent DADD chains in flight instead of one (the kernels that call this run one thread per
    // particle, so a lone chain is pure FP64 latency), and x_j^2 is formed once per U sums.
    template <class X> __device__ __forceinline__ static double eval(const X &x, int n)
    {
        constexpr int U = 8;
        double acc = 0.0;
        int i = 0;
        for (; i + U <= n; i += U) {
            double o[U];
#pragma unroll
            for (int k = 0; k < U; ++k) o[k] = 0.0;
#pragma unroll
            for (int t = 0; t < U; ++t) { // x_{i+t} belongs to the sums that start at or before it
                const double xj = x[i + t];
                const double q = xj * xj;
#pragma unroll
                for (int k = 0; k <= t; ++k) o[k] += q;
            }
            for (int j = i + U; j < n; ++j) {
                const double xj = x[j];
                const double q = xj * xj;
#pragma unroll
                for (int k = 0; k < U; ++k) o[k] += q;
            }
#pragma unroll
            for (int k = 0; k < U; ++k) acc += o[k] * o[k];
        }
        for (; i < n; ++i) {
            double o = 0.0;
            for (int j = i; j < n; ++j) {
                const double xj = x[j];
                o += xj * xj;
            }
            acc += o * o;
        }
        return sqrt(acc);
    }
};

// integrands of the reference's integrator tests (tests/integrators.rs:11-17)
template <> struct Objective<EQS_OBJ_GAUSSIAN> { // (-x * x).exp(); the squares are summed for n > 1
    static constexpr bool streaming = true;
    __device__ __forceinline__ static void begin(ObjAcc &s, int) { s.a = 0.0; }
    __device__ __forceinline__ static void feed(ObjAcc &s, double x, int) { s.a += x * x; }
    __device__ __forceinline__ static double end(const ObjAcc &s, int) { return exp(-s.a); }
};

template <> struct Objective<EQS_OBJ_SINCOS> { // (v[1].cos() + v[0]).sin(), n = 2
    static constexpr bool streaming = true;
    __device__ __forceinline__ static void begin(ObjAcc &s, int) { s.a = 0.0; s.p = 0.0; }
    __device__ __forceinline__ static void feed(ObjAcc &s, double x, int d)
    {
        if (d == 0) s.p = x;
        if (d == 1) s.a = sin(cos(x) + s.p);
    }
    __device__ __forceinline__ static double end(const ObjAcc &s, int) { return s.a; }
};

// integrand of the reference's integrator benchmark (benchmarks/integrators.rs:69-75):
// sum += (-1f64).powi(i) * x.powi(i) for i in 0..1000; powi with a run-time exponent is compiler-rt's __powidf2
__device__ __forceinline__ double powi_rt(double a, int b)
{
    double r = 1.0;
    for (;;) {
        if (b & 1) r *= a;
        b >>= 1;
        if (b == 0) break;
        a *= a;
    }
    return r;
}
template <> struct Objective<EQS_OBJ_SERIES> {
    static constexpr bool streaming = true;
    __device__ __forceinline__ static void begin(ObjAcc &s, int) { s.a = 0.0; }
    __device__ __forceinline__ static void feed(ObjAcc &s, double x, int)
    {
        double sum = 0.0;
#pragma unroll 4 // four independent square-and-multiply chains in flight; the sum itself stays in order
        for (int i = 0; i < 1000; ++i) sum += powi_rt(-1.0, i) * powi_rt(x, i);
        s.a += sum;
    }
    __device__ __forceinline__ static double end(const ObjAcc &s, int) { return s.a; }
};

} // inline namespace
} // namespace eqs

// the user hook: csrc/user_objective.cuh, or the header named by the build (EQS_USER_OBJECTIVE_HEADER, build.py)
#ifdef EQS_USER_OBJECTIVE_HEADER
#include EQS_USER_OBJECTIVE_HEADER
#else
#include "user_objective.cuh"
#endif

namespace eqs {
#if EQS_COS_TABLE
inline namespace cos_table {
#else
inline namespace cos_literal {
#endif
template <> struct Objective<EQS_OBJ_USER> : public UserObjective {};

// M consecutive coordinates d0 .. d0+M-1 of a streaming functor: its feed_vec when it has one, else feed in order
template <class Obj, int M> __device__ __forceinline__ void objective_feed_vec(ObjAcc &acc, const double (&x)[M], int d0)
{
    if constexpr (requires(ObjAcc &a) { Obj::template feed_vec<M>(a, x, d0); }) {
        Obj::template feed_vec<M>(acc, x, d0);
    } else {
#pragma unroll
        for (int j = 0; j < M; ++j) Obj::feed(acc, x[j], d0 + j);
    }
}

// The same for callers that can evaluate a particle a second time: functors with a feed_vec_defer skip their per-call
// range tests and report through hi_max instead (test it against kCosHiLimit after the last coordinate; if it fired, the
// accumulated value is meaningless and objective_on() below gives the right one).
template <class Obj, int M>
__device__ __forceinline__ void objective_feed_vec_defer(ObjAcc &acc, const double (&x)[M], int d0, unsigned &hi_max)
{
    if constexpr (requires(ObjAcc &a, unsigned &h) { Obj::template feed_vec_defer<M>(a, x, d0, h); }) {
        Obj::template feed_vec_defer<M>(acc, x, d0, hi_max);
    } else {
        objective_feed_vec<Obj, M>(acc, x, d0);
    }
}

// evaluate any functor through a view (eqs_objective_eval, f(x0), random-access functors)
template <class Obj, class X> __device__ __forceinline__ double objective_on(const X &x, int n)
{
    if constexpr (Obj::streaming) {
        ObjAcc s;
        Obj::begin(s, n);
        for (int d = 0; d < n; ++d) Obj::feed(s, x[d], d);
        return Obj::end(s, n);
    } else {
        return Obj::eval(x, n);
    }
}

// host-side dispatch over the objective ID: f.template operator()<Objective<ID>>()
template <class F> inline bool dispatch_objective(int id, F &&f)
{
    switch (id) {
    case EQS_OBJ_SPHERE: f.template operator()<Objective<EQS_OBJ_SPHERE>>(); return true;
    case EQS_OBJ_ROSENBROCK: f.template operator()<Objective<EQS_OBJ_ROSENBROCK>>(); return true;
    case EQS_OBJ_RASTRIGIN: f.template operator()<Objective<EQS_OBJ_RASTRIGIN>>(); return true;
    case EQS_OBJ_ACKLEY: f.template operator()<Objective<EQS_OBJ_ACKLEY>>(); return true;
    case EQS_OBJ_THREE_CIRCLES: f.template operator()<Objective<EQS_OBJ_THREE_CIRCLES>>(); return true;
    case EQS_OBJ_HEAVY: f.template operator()<Objective<EQS_OBJ_HEAVY>>(); return true;
    case EQS_OBJ_SERIES: f.template operator()<Objective<EQS_OBJ_SERIES>>(); return true;
    case EQS_OBJ_USER: f.template operator()<Objective<EQS_OBJ_USER>>(); return true;
    case EQS_OBJ_GAUSSIAN: f.template operator()<Objective<EQS_OBJ_GAUSSIAN>>(); return true;
    case EQS_OBJ_SINCOS: f.template operator()<Objective<EQS_OBJ_SINCOS>>(); return true;
    default: return false;
    }
}
} // inline namespace
} // namespace eqs
