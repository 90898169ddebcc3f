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
