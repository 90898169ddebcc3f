// C++ twin of the reference's doctests / examples for this path, through include/eqsolver_b200.hpp.
//   test_builder host   : host-only behaviour (no GPU needed)
//   test_builder gpu    : solves on cuda:0 (examples/global_optimisation.rs, global_optimiser_as_solver.rs style)
#include <cmath>
#include <cstdio>
#include <cstring>

#include "eqsolver_b200.hpp"

using namespace eqsolver;
using namespace eqsolver::global_optimisers;

#define CHECK(cond)                                                                                              \
    do {                                                                                                         \
        if (!(cond)) {                                                                                           \
            std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond);                                          \
            return 1;                                                                                            \
        }                                                                                                        \
    } while (0)

static int host_tests()
{
    // ParticleSwarm::new -> Err(ExternalError) for lower > upper (particle_swarm.rs:124-139)
    auto bad = ParticleSwarm::create(Objective::Rastrigin, {1.0, 0.0}, {0.0, 1.0});
    CHECK(!bad.is_ok() && bad.unwrap_err().kind == SolverError::ExternalError);
    auto bad2 = ParticleSwarm::create(Objective::ThreeCircles, {0.0, 0.0, 0.0}, {1.0, 1.0, 1.0});
    CHECK(!bad2.is_ok() && bad2.unwrap_err().kind == SolverError::IncorrectInput);
    auto ok = ParticleSwarm::create(Objective::Rastrigin, Vector(4, -100.0), Vector(4, 100.0));
    CHECK(ok.is_ok());
    CHECK(eqs_abi_version() == EQS_ABI_VERSION);
    int64_t first = -1, local = -1;
    CHECK(eqs_shard_range(10, 3, 2, &first, &local) == EQS_OK && first == 7 && local == 3);
    std::printf("host tests ok\n");
    return 0;
}

static int gpu_tests()
{
    auto made = Context::create(0);
    CHECK(made.is_ok());
    auto ctx = made.unwrap();
    // doctest of particle_swarm.rs:337-355: 4-D Rastrigin, guess 80, bounds +-100 -- runs and returns Ok
    {
        auto ps = ParticleSwarm::create(Objective::Rastrigin, Vector(4, -100.0), Vector(4, 100.0)).unwrap();
        auto r = ps.with_context(ctx).solve(Vector(4, 80.0));
        CHECK(r.is_ok() && r.unwrap().size() == 4);
    }
    // examples/global_optimiser_as_solver.rs:44-49 with the optimum of tests/multi.rs:70
    {
        auto ps = ParticleSwarm::create(Objective::ThreeCircles, {1.0, 1.0}, {7.0, 7.0}).unwrap();
        eqs_report rep{};
        auto r = ps.with_context(ctx).with_particle_count(500).with_iter_max(300).with_tol(0.0).with_seed(2).solve({4.5, 2.5}, &rep);
        CHECK(r.is_ok());
        CHECK(std::fabs(r.unwrap()[0] - 4.217265312839526) < 1e-6 && std::fabs(r.unwrap()[1] - 2.317879970005811) < 1e-6);
        CHECK(rep.iters_run == 300 && rep.evals == 1 + 500 * 301);
    }
    // CrossEntropy on the same problem
    {
        auto r = CrossEntropy(Objective::ThreeCircles).with_context(ctx).with_sample_size(2000).with_importance_selection_size(50)
                     .with_iter_max(100).with_mean_divisor(MeanDivisor::EliteCount).with_seed(2).solve({4.5, 2.5});
        CHECK(r.is_ok());
        CHECK(std::fabs(r.unwrap()[0] - 4.217265312839526) < 1e-5 && std::fabs(r.unwrap()[1] - 2.317879970005811) < 1e-5);
    }
    // where the reference panics: non-finite standard deviation -> NotANumber
    {
        auto r = CrossEntropy(Objective::Sphere).with_context(ctx).with_std_dev({1.0, INFINITY}).solve({0.0, 0.0});
        CHECK(!r.is_ok() && r.unwrap_err().kind == SolverError::NotANumber);
    }
    // tests/integrators.rs:87-106: the two known integrals at the default 1000 samples, tolerance 0.1
    {
        using eqsolver::integrators::MonteCarlo;
        auto g = MonteCarlo(Objective::Gaussian).with_context(ctx).integrate({0.0}, {1.0});
        CHECK(g.is_ok() && std::fabs(g.unwrap() - 0.746824132812427025) <= 0.1);
        auto sc = MonteCarlo(Objective::SinCos).with_context(ctx).with_sample_count(1 << 20).integrate_with_seed({0.0, 0.0}, {1.0, 1.0}, 7);
        CHECK(sc.is_ok() && std::fabs(sc.unwrap().mean - 0.9248464799519) <= 1e-2);
        // tests/integrators.rs:108-116
        auto mi = eqsolver::integrators::MISER(Objective::SinCos).with_context(ctx).integrate_with_seed({0.0, 0.0}, {1.0, 1.0}, 1729);
        CHECK(mi.is_ok() && std::fabs(mi.unwrap().mean - 0.9248464799519) <= 0.001);
        auto bad = MonteCarlo(Objective::Gaussian).with_context(ctx).integrate({1.0}, {0.0});
        CHECK(!bad.is_ok() && bad.unwrap_err().kind == SolverError::ExternalError);
    }
    // solve_with_draws: the same draws give the same answer, other draws another; a transparent case checked by hand:
    // every step draw 0 leaves v = w v, so with w = 0 the particles stand still and gbest is the best start position
    {
        const int n = 3;
        const int64_t N = 64, passes = 4;
        Vector init(N * 2 * n), steps(passes * N * n * 2, 0.0);
        for (size_t i = 0; i < init.size(); ++i) init[i] = std::fmod(0.6180339887 * (double)(i + 1), 1.0);
        auto ps = ParticleSwarm::create(Objective::Sphere, Vector(n, -1.0), Vector(n, 1.0)).unwrap();
        ps.with_context(ctx).with_particle_count(N).with_iter_max(passes).with_tol(0.0).with_inertia_weight(0.0);
        auto a = ps.solve_with_draws(Vector(n, 0.9), init, steps), b = ps.solve_with_draws(Vector(n, 0.9), init, steps);
        CHECK(a.is_ok() && b.is_ok() && a.unwrap() == b.unwrap());
        double best = 1e300;
        Vector arg(n);
        for (int64_t i = 0; i < N; ++i) { // the particles after the first pass: x0 + v0 (w = 0 keeps them there)
            double c = 0.0;
            Vector x(n);
            for (int d = 0; d < n; ++d) {
                const double pos = init[i * 2 * n + d] * 2.0 + -1.0, vel = init[i * 2 * n + n + d] * 4.0 + -2.0;
                (void)vel;
                x[d] = pos;
                c += pos * pos;
            }
            if (c < best) { best = c; arg = x; }
        }
        CHECK(a.unwrap() == arg);
        auto bad = ps.solve_with_draws(Vector(n, 0.9), Vector(5, 0.0), steps);
        CHECK(!bad.is_ok() && bad.unwrap_err().kind == SolverError::IncorrectInput);
        // CrossEntropy: normals all zero -> every sample is the mean, sigma collapses to 0 and the loop ends
        Vector z(2 * 50 * n, 0.0);
        auto ce = CrossEntropy(Objective::Sphere).with_context(ctx).with_sample_size(50).with_importance_selection_size(5)
                      .with_std_dev(Vector(n, 1.0)).with_mean_divisor(MeanDivisor::EliteCount).solve_with_draws(Vector(n, 0.25), z);
        CHECK(ce.is_ok() && ce.unwrap() == Vector(n, 0.25));
    }
    std::printf("gpu tests ok\n");
    return 0;
}

int main(int argc, char **argv)
{
    if (argc > 1 && !std::strcmp(argv[1], "gpu")) return gpu_tests();
    return host_tests();
}
