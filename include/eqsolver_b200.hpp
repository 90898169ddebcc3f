// eqsolver_b200.hpp -- header-only C++ mirror of eqsolver::global_optimisers over the C ABI (eqsolver_b200.h).
//
// Same builder surface as the Rust reference (AzeezDa/eqsolver v0.4.1):
//   ParticleSwarm::create(objective, lower, upper)   ~ ParticleSwarm::new(f, lower, upper) -> SolverResult<Self>
//                                                       (particle_swarm.rs:119-153)
//   .with_*(..)                                      ~ :179-333 (setters returning &mut Self)
//   .solve(x0) -> SolverResult<Vector>               ~ :356-431
//   CrossEntropy(objective).with_*(..).solve(x0)     ~ cross_entropy.rs:75-85, 108-223, 245-306
// The closure `f` is an Objective ID (device functors are compiled ahead of time).  Link with -leqsolver_b200.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <memory>
#include <string>
#include <utility>
#include <variant>
#include <vector>

#include "eqsolver_b200.h"

namespace eqsolver {

constexpr double DEFAULT_TOL = 1e-6;      // lib.rs:13
constexpr int64_t DEFAULT_ITERMAX = 50;   // lib.rs:16

enum class Objective : int32_t { Sphere = 0, Rosenbrock, Rastrigin, Ackley, ThreeCircles, Heavy, User, Gaussian, SinCos, Series };

// SolverError, lib.rs:19-44
struct SolverError {
    enum Kind { MaxIterReached = 1, NotANumber, IncorrectInput, BadJacobian, TypeConversionError, ExternalError } kind;
    std::string message;
};

template <class T> class SolverResult {
public:
    SolverResult(T v) : v_(std::move(v)) {}
    SolverResult(SolverError e) : v_(std::move(e)) {}
    bool is_ok() const { return v_.index() == 0; }
    const T &unwrap() const { return std::get<0>(v_); }
    const SolverError &unwrap_err() const { return std::get<1>(v_); }

private:
    std::variant<T, SolverError> v_;
};

using Vector = std::vector<double>;

// one GPU: stream, device arena, communicator
class Context {
public:
    static SolverResult<std::shared_ptr<Context>> create(int device = 0)
    {
        eqs_ctx *raw = nullptr;
        const int st = eqs_ctx_create(device, &raw);
        if (st != EQS_OK) return SolverError{static_cast<SolverError::Kind>(st), eqs_last_error(nullptr)};
        return std::shared_ptr<Context>(new Context(raw));
    }
    ~Context() { eqs_ctx_destroy(raw_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    eqs_ctx *raw() const { return raw_; }
    SolverError error(int st) const { return SolverError{static_cast<SolverError::Kind>(st), eqs_last_error(raw_)}; }

private:
    explicit Context(eqs_ctx *raw) : raw_(raw) {}
    eqs_ctx *raw_;
};

namespace global_optimisers {

enum class PsoMode : int32_t { SequentialExact = EQS_PSO_SEQUENTIAL_EXACT, Synchronous = EQS_PSO_SYNCHRONOUS };
enum class MeanDivisor : int32_t { ReferenceTen = EQS_CE_DIV_REFERENCE_TEN, EliteCount = EQS_CE_DIV_ELITE_COUNT };

class ParticleSwarm {
public:
    // ParticleSwarm::new: Err(ExternalError) when a bounds pair is not lower <= upper (particle_swarm.rs:124-139)
    static SolverResult<ParticleSwarm> create(Objective objective, Vector lower_bounds, Vector upper_bounds)
    {
        ParticleSwarm s(objective, std::move(lower_bounds), std::move(upper_bounds));
        if (s.lower_.size() != s.upper_.size())
            return SolverError{SolverError::IncorrectInput, "bounds differ in length"};
        const eqs_pso_params p = s.params();
        const int st = eqs_pso_validate(&p);
        if (st == EQS_ERR_EXTERNAL)
            return SolverError{SolverError::ExternalError, "low > high (or equal if exclusive) in uniform distribution"};
        if (st != EQS_OK) return SolverError{static_cast<SolverError::Kind>(st), "dimension, objective or bounds rejected"};
        return s;
    }

    ParticleSwarm &with_inertia_weight(double v) { w_ = v; return *this; }
    ParticleSwarm &with_cognitive_coefficient(double v) { c1_ = v; return *this; }
    ParticleSwarm &with_social_coefficient(double v) { c2_ = v; return *this; }
    ParticleSwarm &with_tol(double v) { tol_ = v; return *this; }
    ParticleSwarm &with_particle_count(int64_t v) { particle_count_ = v; return *this; }
    ParticleSwarm &with_iter_max(int64_t v) { iter_max_ = v; return *this; }
    ParticleSwarm &with_seed(uint64_t v) { seed_ = v; return *this; }
    ParticleSwarm &with_mode(PsoMode m) { mode_ = m; return *this; }
    ParticleSwarm &with_context(std::shared_ptr<Context> c) { ctx_ = std::move(c); return *this; }

    // solve with an explicit generator state, the twin of integrate_with_rng (integrators/mod.rs:56-61)
    SolverResult<Vector> solve_with_seed(const Vector &x0, uint64_t seed, eqs_report *report = nullptr)
    {
        seed_ = seed;
        return solve(x0, report);
    }
    // solve on caller-supplied draws instead of a generator (the replay mode of the parity tests as a feature; the
    // closest thing the reference has is integrate_with_rng, integrators/mod.rs:56-61).  init_draws [N][2][n]: u in
    // [0, 1) for the start positions, then the velocities (the samplers of particle_swarm.rs:367,373 scale them);
    // step_draws [passes][N][n][2]: (r_p, r_g) of :407-408.  Runs min(passes, iter_max) passes, fewer when the stall
    // test (:433-441) fires.
    SolverResult<Vector> solve_with_draws(const Vector &x0, const Vector &init_draws, const Vector &step_draws) const
    {
        const size_t n = lower_.size(), N = static_cast<size_t>(particle_count_);
        if (x0.size() != n) return SolverError{SolverError::IncorrectInput, "x0 and bounds differ in length"};
        if (init_draws.size() != N * 2 * n) return SolverError{SolverError::IncorrectInput, "init_draws must hold particle_count * 2n draws"};
        const size_t per_pass = N * n * 2;
        if (per_pass == 0 || step_draws.size() % per_pass) return SolverError{SolverError::IncorrectInput, "step_draws must hold whole passes of particle_count * n * 2 draws"};
        std::shared_ptr<Context> ctx = ctx_;
        if (!ctx) {
            auto made = Context::create(0);
            if (!made.is_ok()) return made.unwrap_err();
            ctx = made.unwrap();
        }
        eqs_pso_params p = params();
        p.replay_init = init_draws.data();
        eqs_pso *h = nullptr;
        int st = eqs_pso_create(ctx->raw(), &p, &h);
        if (st != EQS_OK) return ctx->error(st);
        Vector out(n);
        st = eqs_pso_init(h, x0.data());
        const int64_t passes = std::min<int64_t>(static_cast<int64_t>(step_draws.size() / per_pass), iter_max_);
        for (int64_t t = 0; st == EQS_OK && t < passes; ++t) st = eqs_pso_step(h, 1, step_draws.data() + t * per_pass);
        if (st == EQS_OK) st = eqs_pso_read_gbest(h, out.data(), nullptr, nullptr, nullptr, nullptr);
        const SolverError err = st == EQS_OK ? SolverError{} : ctx->error(st);
        eqs_pso_destroy(h);
        if (st != EQS_OK) return err;
        return out;
    }
    // solve(x0), particle_swarm.rs:356-431: the global best position
    SolverResult<Vector> solve(const Vector &x0, eqs_report *report = nullptr) const
    {
        if (x0.size() != lower_.size()) return SolverError{SolverError::IncorrectInput, "x0 and bounds differ in length"};
        std::shared_ptr<Context> ctx = ctx_;
        if (!ctx) {
            auto made = Context::create(0);
            if (!made.is_ok()) return made.unwrap_err();
            ctx = made.unwrap();
        }
        Vector out(x0.size());
        const eqs_pso_params p = params();
        const int st = eqs_pso_solve(ctx->raw(), &p, x0.data(), out.data(), report);
        if (st != EQS_OK) return ctx->error(st);
        return out;
    }

private:
    ParticleSwarm(Objective o, Vector lo, Vector hi) : objective_(o), lower_(std::move(lo)), upper_(std::move(hi)) {}
    eqs_pso_params params() const
    {
        eqs_pso_params p{};
        p.n = static_cast<int32_t>(lower_.size());
        p.objective_id = static_cast<int32_t>(objective_);
        p.mode = static_cast<int32_t>(mode_);
        p.exchange = EQS_EXCHANGE_AUTO;
        p.particle_count = particle_count_;
        p.iter_max = iter_max_;
        p.tol = tol_;
        p.w = w_;
        p.c1 = c1_;
        p.c2 = c2_;
        p.lower = lower_.data();
        p.upper = upper_.data();
        p.seed = seed_;
        p.virtual_world = 1;
        return p;
    }
    Objective objective_;
    Vector lower_, upper_;
    double w_ = 0.5, c1_ = 1.0, c2_ = 1.0, tol_ = DEFAULT_TOL; // particle_swarm.rs:8-10
    int64_t iter_max_ = DEFAULT_ITERMAX, particle_count_ = 100; // :11
    uint64_t seed_ = 0;
    PsoMode mode_ = PsoMode::SequentialExact;
    std::shared_ptr<Context> ctx_;
};

class CrossEntropy {
public:
    explicit CrossEntropy(Objective objective) : objective_(objective) {} // CrossEntropy::new, cross_entropy.rs:75-85

    CrossEntropy &with_tol(double v) { tol_ = v; return *this; }
    CrossEntropy &with_iter_max(int64_t v) { iter_max_ = v; return *this; }
    CrossEntropy &with_sample_size(int64_t v) { sample_size_ = v; return *this; }
    CrossEntropy &with_importance_selection_size(int64_t v) { elite_ = v; return *this; }
    CrossEntropy &with_std_dev(Vector v) { std_dev_ = std::move(v); return *this; }
    CrossEntropy &with_seed(uint64_t v) { seed_ = v; return *this; }
    CrossEntropy &with_mean_divisor(MeanDivisor m) { divisor_ = m; return *this; }
    CrossEntropy &with_context(std::shared_ptr<Context> c) { ctx_ = std::move(c); return *this; }

    SolverResult<Vector> solve_with_seed(const Vector &x0, uint64_t seed, eqs_report *report = nullptr)
    {
        seed_ = seed;
        return solve(x0, report);
    }
    // solve on caller-supplied standard normals `normals` [passes][N][n] (the z of Normal::sample, cross_entropy.rs:270):
    // runs min(passes, iter_max - 1) passes (`iter` starts at 1, :253) or stops when max |sigma| <= tol
    SolverResult<Vector> solve_with_draws(const Vector &x0, const Vector &normals) const
    {
        if (!std_dev_.empty() && std_dev_.size() != x0.size())
            return SolverError{SolverError::IncorrectInput, "std_dev and x0 differ in length"};
        const size_t per_pass = static_cast<size_t>(sample_size_) * x0.size();
        if (per_pass == 0 || normals.size() % per_pass) return SolverError{SolverError::IncorrectInput, "normals must hold whole passes of sample_size * n draws"};
        std::shared_ptr<Context> ctx = ctx_;
        if (!ctx) {
            auto made = Context::create(0);
            if (!made.is_ok()) return made.unwrap_err();
            ctx = made.unwrap();
        }
        const eqs_ce_params p = params(x0.size());
        eqs_ce *h = nullptr;
        int st = eqs_ce_create(ctx->raw(), &p, x0.data(), &h);
        if (st != EQS_OK) return ctx->error(st);
        Vector out(x0.size());
        const int64_t passes = std::min<int64_t>(static_cast<int64_t>(normals.size() / per_pass), std::max<int64_t>(iter_max_ - 1, 0));
        for (int64_t t = 0; st == EQS_OK && t < passes; ++t) st = eqs_ce_step(h, 1, normals.data() + t * per_pass);
        if (st == EQS_OK) st = eqs_ce_read_state(h, out.data(), nullptr, nullptr, nullptr);
        const SolverError err = st == EQS_OK ? SolverError{} : ctx->error(st);
        eqs_ce_destroy(h);
        if (st != EQS_OK) return err;
        return out;
    }
    // solve(x0), cross_entropy.rs:245-306: the final mean; the reference's panics become NotANumber
    SolverResult<Vector> solve(const Vector &x0, eqs_report *report = nullptr) const
    {
        if (!std_dev_.empty() && std_dev_.size() != x0.size())
            return SolverError{SolverError::IncorrectInput, "std_dev and x0 differ in length"};
        std::shared_ptr<Context> ctx = ctx_;
        if (!ctx) {
            auto made = Context::create(0);
            if (!made.is_ok()) return made.unwrap_err();
            ctx = made.unwrap();
        }
        const eqs_ce_params p = params(x0.size());
        Vector out(x0.size());
        const int st = eqs_ce_solve(ctx->raw(), &p, x0.data(), out.data(), report);
        if (st != EQS_OK) return ctx->error(st);
        return out;
    }

private:
    eqs_ce_params params(size_t n) const
    {
        eqs_ce_params p{};
        p.n = static_cast<int32_t>(n);
        p.objective_id = static_cast<int32_t>(objective_);
        p.mean_divisor = static_cast<int32_t>(divisor_);
        p.exchange = EQS_EXCHANGE_AUTO;
        p.sample_size = sample_size_;
        p.elite_count = elite_;
        p.iter_max = iter_max_;
        p.tol = tol_;
        p.std_dev = std_dev_.empty() ? nullptr : std_dev_.data();
        p.seed = seed_;
        p.virtual_world = 1;
        return p;
    }
    Objective objective_;
    Vector std_dev_;
    double tol_ = DEFAULT_TOL;
    int64_t iter_max_ = DEFAULT_ITERMAX, sample_size_ = 100, elite_ = 10; // cross_entropy.rs:8-9
    uint64_t seed_ = 0;
    MeanDivisor divisor_ = MeanDivisor::ReferenceTen;
    std::shared_ptr<Context> ctx_;
};

} // namespace global_optimisers

namespace integrators {

constexpr int64_t DEFAULT_SAMPLE_COUNT = 1000; // integrators/mod.rs:18

// MeanVariance<T>, lib.rs:56-60
struct MeanVariance {
    double mean, variance;
    double standard_deviation() const { return std::sqrt(variance); } // lib.rs:133-135
};

// MonteCarlo, integrators/monte_carlo.rs:85-212; the integrand is an objective ID, the rng a Philox seed
class MonteCarlo {
public:
    explicit MonteCarlo(Objective integrand) : integrand_(integrand) {}
    MonteCarlo &with_sample_count(int64_t v) { sample_count_ = v; return *this; } // :111-114
    MonteCarlo &with_context(std::shared_ptr<Context> c) { ctx_ = std::move(c); return *this; }

    // integrate_with_rng, :181-212
    SolverResult<MeanVariance> integrate_with_seed(const Vector &from, const Vector &to, uint64_t seed = 0, eqs_report *report = nullptr)
    {
        if (from.size() != to.size()) return SolverError{SolverError::IncorrectInput, "bounds differ in length"};
        auto ctx = ctx_;
        if (!ctx) {
            auto made = Context::create(0);
            if (!made.is_ok()) return made.unwrap_err();
            ctx = made.unwrap();
        }
        eqs_mc_params p{};
        p.n = static_cast<int32_t>(from.size());
        p.integrand_id = static_cast<int32_t>(integrand_);
        p.sample_count = sample_count_;
        p.from = from.data();
        p.to = to.data();
        p.seed = seed;
        MeanVariance out{};
        const int st = eqs_mc_integrate(ctx->raw(), &p, &out.mean, &out.variance, report);
        if (st != EQS_OK) return ctx->error(st);
        return out;
    }
    // integrate, integrators/mod.rs:95-98
    SolverResult<double> integrate(const Vector &from, const Vector &to)
    {
        auto r = integrate_with_seed(from, to);
        if (!r.is_ok()) return r.unwrap_err();
        return r.unwrap().mean;
    }

private:
    Objective integrand_;
    int64_t sample_count_ = DEFAULT_SAMPLE_COUNT;
    std::shared_ptr<Context> ctx_;
};

// MISER, integrators/miser.rs:93-293 (defaults :13-23); the recursion :301-443 runs level by level in the library
class MISER {
public:
    explicit MISER(Objective integrand) : integrand_(integrand) {}
    MISER &with_sample_count(int64_t v) { sample_count_ = v; return *this; }
    MISER &with_minimum_dimensional_sample_count(int64_t v) { min_dim_ = v; return *this; }
    MISER &with_maximum_depth(int32_t v) { depth_ = v; return *this; }
    MISER &with_alpha(double v) { alpha_ = v; return *this; }
    MISER &with_dither(double v) { dither_ = v; return *this; }
    MISER &with_context(std::shared_ptr<Context> c) { ctx_ = std::move(c); return *this; }

    // integrate_with_rng, :277-292
    SolverResult<MeanVariance> integrate_with_seed(const Vector &from, const Vector &to, uint64_t seed = 0, eqs_report *report = nullptr)
    {
        if (from.size() != to.size()) return SolverError{SolverError::IncorrectInput, "bounds differ in length"};
        auto ctx = ctx_;
        if (!ctx) {
            auto made = Context::create(0);
            if (!made.is_ok()) return made.unwrap_err();
            ctx = made.unwrap();
        }
        eqs_miser_params p{};
        p.n = static_cast<int32_t>(from.size());
        p.integrand_id = static_cast<int32_t>(integrand_);
        p.sample_count = sample_count_;
        p.minimum_dimensional_sample_count = min_dim_;
        p.maximum_depth = depth_;
        p.alpha = alpha_;
        p.dither = dither_;
        p.from = from.data();
        p.to = to.data();
        p.seed = seed;
        MeanVariance out{};
        const int st = eqs_miser_integrate(ctx->raw(), &p, &out.mean, &out.variance, report);
        if (st != EQS_OK) return ctx->error(st);
        return out;
    }
    SolverResult<double> integrate(const Vector &from, const Vector &to)
    {
        auto r = integrate_with_seed(from, to);
        if (!r.is_ok()) return r.unwrap_err();
        return r.unwrap().mean;
    }

private:
    Objective integrand_;
    int64_t sample_count_ = DEFAULT_SAMPLE_COUNT, min_dim_ = 32;
    int32_t depth_ = 5;
    double alpha_ = 2.0, dither_ = 0.05;
    std::shared_ptr<Context> ctx_;
};

} // namespace integrators
} // namespace eqsolver
