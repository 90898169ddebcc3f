"""GPU: MonteCarlo integration against the CPU oracle (sequential Welford, lib.rs:172-202)."""
import numpy as np
import pytest

import eqsolver_b200 as E
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def test_reference_integrals_like_the_reference_tests(ctx):
    # tests/integrators.rs:87-106: default 1000 samples, tolerance 0.1 on the known integrals
    r = E.MonteCarlo.new(E.Objective.GAUSSIAN).with_context(ctx).integrate_with_seed(0.0, 1.0, 1729)
    assert abs(r.mean - 0.746824132812427025) <= 0.1
    r = E.MonteCarlo.new(E.Objective.SINCOS).with_context(ctx).integrate_with_seed([0.0, 0.0], [1.0, 1.0], 1729)
    assert abs(r.mean - 0.9248464799519) <= 0.1
    # and much tighter with the sample counts a GPU is for
    r = (E.MonteCarlo.new(E.Objective.GAUSSIAN).with_sample_count(1 << 24).with_context(ctx)
         .integrate_with_seed(0.0, 1.0, 1729))
    assert abs(r.mean - 0.746824132812427025) <= 5 * np.sqrt(r.variance / (1 << 24))
    assert abs(E.MonteCarlo.new(E.Objective.SINCOS).with_sample_count(1 << 22).with_context(ctx)
               .integrate([0.0, 0.0], [1.0, 1.0]) - 0.9248464799519) < 1e-3


@pytest.mark.parametrize("integrand,n,lo,hi,exact", [(E.Objective.SPHERE, 5, -1.0, 2.0, True),
                                                      (E.Objective.ROSENBROCK, 4, -2.0, 2.0, True),
                                                      (E.Objective.GAUSSIAN, 1, 0.0, 1.0, False),
                                                      (E.Objective.GAUSSIAN, 3, -1.0, 1.0, False),
                                                      (E.Objective.SINCOS, 2, 0.0, 1.0, False),
                                                      (E.Objective.RASTRIGIN, 6, -5.12, 5.12, False)])
@pytest.mark.parametrize("count", [2, 1000, 100003])
def test_philox_parity_vs_oracle(ctx, integrand, n, lo, hi, exact, count):
    """Same Philox draws on both sides; parallel Welford merge vs the reference's sequential recurrence."""
    lower, upper = [lo] * n, [hi] * n
    got = E.MonteCarlo.new(integrand).with_sample_count(count).with_context(ctx).integrate_with_seed(lower, upper, 99)
    st, mean, var = O.mc_integrate(int(integrand), lower, upper, count, 99)
    assert st == 0
    assert abs(got.mean - mean) <= 1e-12 * max(1.0, abs(mean))
    assert abs(got.variance - var) <= 1e-10 * max(1.0, abs(var))


def test_replay_draws_and_determinism(ctx):
    n, count = 3, 5000
    u = np.random.Generator(np.random.Philox(1729)).random((count, n))
    lower, upper = [-1.0, 0.0, 2.0], [1.0, 0.5, 3.5]
    mc = E.MonteCarlo.new(E.Objective.ROSENBROCK).with_sample_count(count).with_context(ctx)
    a = mc.integrate_with_seed(lower, upper, replay_u=u)
    b = mc.integrate_with_seed(lower, upper, replay_u=u)
    assert a == b                                   # fixed merge order: bit-reproducible
    st, mean, var = O.mc_integrate(O.ROSENBROCK, lower, upper, count, replay_u=u)
    assert abs(a.mean - mean) <= 1e-12 * abs(mean) and abs(a.variance - var) <= 1e-10 * abs(var)
    # a constant integrand (zero-width box): exactly f(x) * 0 volume
    r = E.MonteCarlo.new(E.Objective.SPHERE).with_context(ctx).integrate_with_seed([1.0, 2.0], [1.0, 2.0], 3)
    assert r.mean == 0.0 and r.variance == 0.0


def test_errors_like_the_reference(ctx):
    with pytest.raises(E.ExternalError):     # Uniform::new_inclusive(from > to), monte_carlo.rs:171-172
        E.MonteCarlo.new(E.Objective.GAUSSIAN).with_context(ctx).integrate_with_seed(1.0, 0.0)
    with pytest.raises(E.IncorrectInput):    # "the iterator must be over at least 2 elements", lib.rs:190-193
        E.MonteCarlo.new(E.Objective.GAUSSIAN).with_sample_count(1).with_context(ctx).integrate_with_seed(0.0, 1.0)
    with pytest.raises(E.IncorrectInput):
        E.MonteCarlo.new(E.Objective.SINCOS).with_context(ctx).integrate_with_seed(0.0, 1.0)   # needs n = 2
    with pytest.raises(E.IncorrectInput):
        E.MonteCarlo.new(E.Objective.HEAVY).with_context(ctx).integrate_with_seed([0.0] * 3, [1.0] * 3)


# ---- MISER -----------------------------------------------------------------------------------

def test_miser_reference_test_and_defaults(ctx):
    # tests/integrators.rs:108-116: default parameters on sin(cos y + x) over the unit square, tolerance 0.001
    r = E.MISER.new(E.Objective.SINCOS).with_context(ctx).integrate_with_seed([0.0, 0.0], [1.0, 1.0], 1729)
    assert abs(r.mean - 0.9248464799519) <= 0.001
    assert abs(E.MISER.new(E.Objective.GAUSSIAN).with_context(ctx).integrate(0.0, 1.0) - 0.746824132812427025) <= 0.005


@pytest.mark.parametrize("integrand,lower,upper", [(E.Objective.SINCOS, [0.0, 0.0], [1.0, 1.0]),
                                                    (E.Objective.GAUSSIAN, [0.0], [1.0]),
                                                    (E.Objective.GAUSSIAN, [-1.0, 0.0, -2.0], [1.0, 2.0, 0.5]),
                                                    (E.Objective.SPHERE, [0.0, 0.0, 0.0], [1.0, 2.0, 3.0]),
                                                    (E.Objective.ROSENBROCK, [-2.0] * 5, [2.0] * 5),
                                                    (E.Objective.RASTRIGIN, [-5.12] * 4, [5.12] * 4)])
@pytest.mark.parametrize("seed", [0, 1729])
def test_miser_matches_the_recursive_oracle(ctx, integrand, lower, upper, seed):
    """Level-by-level GPU evaluation vs the depth-first oracle: same tree (same points spent), same sums."""
    mi = E.MISER.new(integrand).with_context(ctx)
    got = mi.integrate_with_seed(lower, upper, seed)
    st, mean, var, evals = O.miser_integrate(int(integrand), lower, upper, seed=seed)
    assert st == 0 and mi.last_report.evals == evals        # identical bisections and point allotments
    assert abs(got.mean - mean) <= 1e-12 * max(1.0, abs(mean))
    assert abs(got.variance - var) <= 1e-9 * max(1e-12, abs(var))
    assert got == mi.integrate_with_seed(lower, upper, seed)  # bit-reproducible


@pytest.mark.parametrize("kw", [dict(sample_count=200000, maximum_depth=8), dict(sample_count=5000, maximum_depth=0),
                                dict(sample_count=50000, min_dim_count=4, alpha=0.5, dither=0.2),
                                dict(sample_count=3, maximum_depth=3, min_dim_count=1)])
def test_miser_parameters(ctx, kw):
    lower, upper = [0.0, 0.0], [1.0, 1.0]
    mi = E.MISER.new(E.Objective.SINCOS).with_context(ctx).with_sample_count(kw["sample_count"]) \
        .with_maximum_depth(kw.get("maximum_depth", 5)) \
        .with_minimum_dimensional_sample_count(kw.get("min_dim_count", 32)).with_alpha(kw.get("alpha", 2.0)) \
        .with_dither(kw.get("dither", 0.05))
    got = mi.integrate_with_seed(lower, upper, 7)
    st, mean, var, evals = O.miser_integrate(O.SINCOS, lower, upper, sample_count=kw["sample_count"],
                                             maximum_depth=kw.get("maximum_depth", 5),
                                             min_dim_count=kw.get("min_dim_count", 32), alpha=kw.get("alpha", 2.0),
                                             dither=kw.get("dither", 0.05), seed=7)
    assert st == 0 and mi.last_report.evals == evals
    assert abs(got.mean - mean) <= 1e-12 and abs(got.variance - var) <= 1e-9 * abs(var)


def test_miser_errors_like_the_reference(ctx):
    mk = lambda obj=E.Objective.SPHERE: E.MISER.new(obj).with_context(ctx)
    with pytest.raises(E.ExternalError):        # uniform(lower, upper), miser.rs:296-299,320-324
        mk().integrate_with_seed([0.0, 0.0], [1.0, -1.0])
    with pytest.raises(E.ExternalError):        # uniform(-dither, dither), :361
        mk().with_dither(-0.1).integrate_with_seed([0.0, 0.0], [1.0, 1.0])
    with pytest.raises(E.ExternalError):        # a dithered mid-point outside the region, :366-367
        mk().with_dither(0.7).integrate_with_seed([0.0, 0.0], [1.0, 1.0], 3)
    with pytest.raises(E.TypeConversionError):  # constant integrand on a zero-width box: 0/0 points, :413-419
        mk().integrate_with_seed([0.0, 0.0], [0.0, 0.0])
    with pytest.raises(E.IncorrectInput):       # a leaf with fewer than 2 points, lib.rs:190-193
        mk().with_minimum_dimensional_sample_count(0).with_sample_count(1).with_maximum_depth(0) \
            .integrate_with_seed([0.0, 0.0], [1.0, 1.0])
    with pytest.raises(E.IncorrectInput):
        mk(E.Objective.HEAVY).integrate_with_seed([0.0] * 3, [1.0] * 3)
    for (kw, want) in [(dict(dither=-0.1), O.ERR_EXTERNAL), (dict(dither=0.7, seed=3), O.ERR_EXTERNAL)]:
        assert O.miser_integrate(O.SPHERE, [0.0, 0.0], [1.0, 1.0], **kw)[0] == want
