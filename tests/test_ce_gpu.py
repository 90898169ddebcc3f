"""GPU: the CUDA CrossEntropy path (through the C ABI) against the CPU oracle on identical draws.

Tolerances (SURVEY.md §7.2 hard part 4): the elite SET is bit-exact absent ties at the k-th cost (ties: lowest
sample index first on both sides); mu and sigma are sums over the elites taken in sample order on the GPU and in
cost order in the reference, so they agree to 1e-12 scaled by the magnitude of the elite coordinates.
"""
import numpy as np
import pytest

import eqsolver_b200 as E
from conftest import from_bits, golden
from oracle import oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-12


def close(a, b, scale, what=""):
    np.testing.assert_allclose(a, b, rtol=RTOL, atol=RTOL * scale, err_msg=what)


def test_golden_trajectories_replay(ctx):
    for g in golden("ce_golden.json"):
        z = np.array([[[from_bits(b) for b in row] for row in it] for it in g["normals"]])
        ce = (E.CrossEntropy.new(g["objective"]).with_sample_size(g["N"]).with_importance_selection_size(g["k"])
              .with_iter_max(g["iter_max"]).with_std_dev(g["sigma0"]).with_context(ctx)
              .with_mean_divisor(E.MeanDivisor.REFERENCE_TEN if g["divisor_ten"] else E.MeanDivisor.ELITE_COUNT))
        run = ce.start(g["x0"])
        scale = 10.0
        for t, tr in enumerate(g["trace"]):
            run.step(1, z[t:t + 1])
            mu, sg, it, active = run.state()
            assert it == t + 1
            close(mu, [from_bits(b) for b in tr["mu"]], scale, f"mu pass {t}")
            close(sg, [from_bits(b) for b in tr["sigma"]], scale, f"sigma pass {t}")
            assert sorted(run.elites()) == sorted(tr["elites"])
        assert not run.state()[3]  # iter reached iter_max: loop condition false (cross_entropy.rs:255)
        before = run.state()
        run.step(1, z[:1])         # a pass after the loop ended is a no-op
        after = run.state()
        assert np.array_equal(before[0], after[0]) and before[2] == after[2]


@pytest.mark.parametrize("objective", [E.Objective.SPHERE, E.Objective.ROSENBROCK, E.Objective.RASTRIGIN,
                                       E.Objective.ACKLEY, E.Objective.USER, E.Objective.HEAVY])
@pytest.mark.parametrize("n", [8, 7])
def test_philox_parity_per_iteration(ctx, objective, n):
    """Both sides derive the SAME normals from Philox counters (device log/sincospi vs glibc: <= 4e-15)."""
    N, k, iters = 6000, 60, 6
    x0, s0 = np.full(n, 3.0), np.full(n, 2.0)
    ce = (E.CrossEntropy.new(objective).with_sample_size(N).with_importance_selection_size(k).with_iter_max(iters + 1)
          .with_std_dev(s0).with_mean_divisor(E.MeanDivisor.ELITE_COUNT).with_seed(1729).with_tol(0.0)
          .with_context(ctx))
    run = ce.start(x0)
    o = O.Ce(int(objective), x0, sample_size=N, elite_count=k, iter_max=iters + 1, std_dev=s0,
             mean_divisor=O.DIV_ELITE_COUNT, seed=1729, tol=0.0)
    for t in range(iters):
        run.step(1)
        assert o.step() == 0
        oc, gc = o.costs(), run.costs()
        np.testing.assert_allclose(gc, oc, rtol=1e-11, atol=1e-11, err_msg=f"costs pass {t}")
        oe, ge = set(o.elites().tolist()), set(run.elites().tolist())
        if oe != ge:  # only legal when the k-th and (k+1)-th costs are within rounding of each other
            srt = np.sort(oc)
            assert abs(srt[k] - srt[k - 1]) <= 1e-11 * max(1.0, abs(srt[k])), f"elite sets differ at pass {t}"
            pytest.skip("near-tie at the k-th cost: trajectories legitimately diverge")
        mu, sg, it, _ = run.state()
        omu, osg = o.state()
        close(mu, omu, 5.0, f"mu pass {t}")
        close(sg, osg, 5.0, f"sigma pass {t}")
        assert it == o.iters()


@pytest.mark.parametrize("n", [8, 7, 5])
def test_cosine_arguments_out_of_range_in_some_coordinates(ctx, n):
    """The sample kernel evaluates four cosines in lock step and repairs the out-of-range ones (|2 pi x| >= 8e5 goes to
    the full-range cosine) AFTER the arithmetic (devmath.cuh: eqs_cos_vec, LATE).  Ackley's exp(mean cos) shows every
    cosine: means of 1e16 .. 2e17 in some coordinates, ordinary ones in the others, costs against the oracle's (glibc)."""
    N, k = 3000, 30
    x0 = np.array([1e16, 0.3, -2e17, 1.0, 0.5, 3e15, 0.1, -7e16])[:n]
    s0 = np.full(n, 1.0)
    ce = (E.CrossEntropy.new(E.Objective.ACKLEY).with_sample_size(N).with_importance_selection_size(k).with_iter_max(3)
          .with_std_dev(s0).with_mean_divisor(E.MeanDivisor.ELITE_COUNT).with_seed(99).with_tol(0.0).with_context(ctx))
    run = ce.start(x0)
    o = O.Ce(int(E.Objective.ACKLEY), x0, sample_size=N, elite_count=k, iter_max=3, std_dev=s0,
             mean_divisor=O.DIV_ELITE_COUNT, seed=99, tol=0.0)
    run.step(1)
    assert o.step() == 0
    gc, oc = run.costs(), o.costs()
    assert np.isfinite(gc).all()
    assert np.ptp(oc) > 0.5  # the cosines of the huge coordinates differ from sample to sample: the check has teeth
    np.testing.assert_allclose(gc, oc, rtol=1e-12, atol=1e-12)


def test_ties_at_kth_cost_lowest_index_first(ctx):
    """Duplicate draws => exactly equal costs; the k-th cost is shared and the tie must split by sample index."""
    n, N, k = 4, 96, 11
    rng = np.random.default_rng(3)
    base = rng.standard_normal((N // 4, n))
    z = np.repeat(base, 4, axis=0)[None]          # every cost appears 4 times
    ce = (E.CrossEntropy.new(E.Objective.SPHERE).with_sample_size(N).with_importance_selection_size(k)
          .with_iter_max(3).with_mean_divisor(E.MeanDivisor.ELITE_COUNT).with_context(ctx))
    run = ce.start(np.zeros(n))
    o = O.Ce(O.SPHERE, np.zeros(n), sample_size=N, elite_count=k, iter_max=3, mean_divisor=O.DIV_ELITE_COUNT,
             replay_normals=z)
    run.step(1, z); o.step()
    assert sorted(run.elites().tolist()) == sorted(o.elites().tolist())
    close(run.state()[0], o.state()[0], 3.0); close(run.state()[1], o.state()[1], 3.0)
    # all costs equal: everything is a tie
    z1 = np.repeat(rng.standard_normal((1, n)), N, axis=0)[None]
    run = ce.start(np.zeros(n))
    run.step(1, z1)
    assert run.elites().tolist() == list(range(k))
    # identical elites: sigma is what the reference's sequential sums make of it (not exactly 0: 11 x / 11 != x)
    o1 = O.Ce(O.SPHERE, np.zeros(n), sample_size=N, elite_count=k, iter_max=3, mean_divisor=O.DIV_ELITE_COUNT,
              replay_normals=z1)
    o1.step()
    assert np.array_equal(run.state()[0].view(np.uint64), o1.state()[0].view(np.uint64))
    assert np.array_equal(run.state()[1].view(np.uint64), o1.state()[1].view(np.uint64))
    assert run.state()[1].max() < 1e-15


def test_reference_default_solve(ctx):
    """Reference defaults (N=100, k=10, divisor 10, <=49 passes) like benchmarks/multi.rs:86-100."""
    n = 30
    x0 = np.full(n, 0.4)
    ce = E.CrossEntropy.new(E.Objective.ROSENBROCK).with_std_dev(np.ones(n)).with_seed(5).with_context(ctx)
    got = ce.solve(x0)
    o = O.Ce(O.ROSENBROCK, x0, std_dev=np.ones(n), seed=5)
    want = o.solve()
    assert ce.last_report.iters_run == o.iters()
    assert ce.last_report.evals == 100 * o.iters()
    close(got, want, 1.0)


def test_edge_sizes(ctx):
    for (N, k, n) in [(2, 2, 1), (3, 2, 2), (5, 5, 3), (1023, 2, 5), (1024, 1024, 2), (1025, 513, 9), (4097, 33, 1)]:
        x0, s0 = np.full(n, 1.0), np.full(n, 0.5)
        ce = (E.CrossEntropy.new(E.Objective.SPHERE).with_sample_size(N).with_importance_selection_size(k)
              .with_iter_max(4).with_std_dev(s0).with_mean_divisor(E.MeanDivisor.ELITE_COUNT).with_seed(N)
              .with_tol(0.0).with_context(ctx))
        run = ce.start(x0)
        o = O.Ce(O.SPHERE, x0, sample_size=N, elite_count=k, iter_max=4, std_dev=s0, mean_divisor=O.DIV_ELITE_COUNT,
                 seed=N, tol=0.0)
        for _ in range(3):
            run.step(1); o.step()
            assert set(run.elites().tolist()) == set(o.elites().tolist()), (N, k, n)
            close(run.state()[0], o.state()[0], 2.0, str((N, k, n)))
            close(run.state()[1], o.state()[1], 2.0, str((N, k, n)))


def test_errors_where_the_reference_panics(ctx):
    with pytest.raises(E.IncorrectInput):
        E.CrossEntropy.new(E.Objective.SPHERE).with_importance_selection_size(1).with_context(ctx).solve([0.0, 0.0])
    with pytest.raises(E.IncorrectInput):
        (E.CrossEntropy.new(E.Objective.SPHERE).with_sample_size(5).with_importance_selection_size(6)
         .with_context(ctx).solve([0.0, 0.0]))
    # non-finite sigma: Normal::new(..).unwrap() panics (cross_entropy.rs:259)
    with pytest.raises(E.NotANumber):
        E.CrossEntropy.new(E.Objective.SPHERE).with_std_dev([1.0, float("inf")]).with_context(ctx).solve([0.0, 0.0])
    # NaN cost: partial_cmp().unwrap() panics (cross_entropy.rs:276).  x0 = NaN makes every cost NaN.
    with pytest.raises(E.NotANumber):
        E.CrossEntropy.new(E.Objective.SPHERE).with_context(ctx).solve([float("nan"), 0.0])
    # divisor 10 with k = 100 (reference-exact, F7) overflows within the 49 passes -> non-finite sigma -> panic
    with pytest.raises(E.NotANumber):
        (E.CrossEntropy.new(E.Objective.RASTRIGIN).with_sample_size(1000).with_importance_selection_size(100)
         .with_std_dev([2.0] * 8).with_iter_max(500).with_context(ctx).solve([3.0] * 8))
    # iter_max <= 1: the loop never runs, Ok(x0) even with a bad sigma
    got = E.CrossEntropy.new(E.Objective.SPHERE).with_iter_max(1).with_std_dev([float("inf")] * 2).with_context(ctx).solve([1.0, 2.0])
    assert got.tolist() == [1.0, 2.0]


def test_converges_to_known_optima(ctx):
    mu = (E.CrossEntropy.new(E.Objective.RASTRIGIN).with_sample_size(20000).with_importance_selection_size(200)
          .with_std_dev([2.0] * 8).with_iter_max(100).with_mean_divisor(E.MeanDivisor.ELITE_COUNT).with_seed(1729)
          .with_context(ctx).solve([3.0] * 8))
    assert np.abs(mu).max() < 1e-5
    mu = (E.CrossEntropy.new(E.Objective.THREE_CIRCLES).with_sample_size(2000).with_importance_selection_size(50)
          .with_iter_max(100).with_mean_divisor(E.MeanDivisor.ELITE_COUNT).with_seed(2).with_context(ctx)
          .solve([4.5, 2.5]))
    assert np.allclose(mu, [4.217265312839526, 2.317879970005811], atol=1e-5)  # tests/multi.rs:70


def test_virtual_ranks_match_single_rank(ctx):
    n, N, k, world, iters = 6, 5001, 77, 3, 4
    x0, s0 = np.full(n, 3.0), np.full(n, 2.0)
    base = lambda: (E.CrossEntropy.new(E.Objective.ROSENBROCK).with_sample_size(N).with_importance_selection_size(k)
                    .with_iter_max(iters + 1).with_std_dev(s0).with_mean_divisor(E.MeanDivisor.ELITE_COUNT)
                    .with_seed(11).with_tol(0.0).with_context(ctx))
    single = base().start(x0)
    ranks = [base().with_virtual_rank(r, world).start(x0) for r in range(world)]
    assert sum(r.local for r in ranks) == N
    for t in range(iters):
        single.step(1)
        for ph in range(E.CeRun.N_PHASES):
            for r in ranks:
                r.phase(ph)
            if ranks[0].exchange_words(ph):
                words = np.stack([r.get_exchange(ph) for r in ranks])
                for r in ranks:
                    r.set_exchange(ph, words)
        smu, ssg, sit, sact = single.state()
        el = np.sort(np.concatenate([r.elites() for r in ranks]))
        assert el.tolist() == np.sort(single.elites()).tolist()
        for r in ranks:
            mu, sg, it, act = r.state()
            close(mu, smu, 5.0, f"mu pass {t}"); close(sg, ssg, 5.0, f"sigma pass {t}")
            assert it == sit and act == sact
    # mu/sigma agree to rounding (summation order differs with the rank count), so the costs do too
    costs = np.concatenate([r.costs() for r in ranks])
    np.testing.assert_allclose(costs, single.costs(), rtol=1e-10)


def test_full_size_c3_properties(ctx):
    """BASELINE config 3 at full size (Rastrigin n=64, 2^22 samples, 1% elites): size-independent checks."""
    n, N = 64, 1 << 22
    k = N // 100
    x0, s0 = np.full(n, 3.0), np.full(n, 2.0)
    ce = (E.CrossEntropy.new(E.Objective.RASTRIGIN).with_sample_size(N).with_importance_selection_size(k)
          .with_iter_max(21).with_std_dev(s0).with_mean_divisor(E.MeanDivisor.ELITE_COUNT).with_seed(1729)
          .with_tol(0.0).with_context(ctx))
    run = ce.start(x0)
    run.step(1)
    costs, el = run.costs(), run.elites()
    assert el.size == k and np.all(np.diff(el) > 0)                 # exactly k elites, ascending sample order
    mask = np.zeros(N, dtype=bool); mask[el] = True
    assert costs[mask].max() <= costs[~mask].min()                    # the elites are the k smallest costs
    assert costs[mask].max() == np.partition(costs, k - 1)[k - 1]     # and the threshold is the k-th order statistic
    # regenerate the elites on the host from their counters and refit: mu/sigma of the GPU within 1e-12
    X = np.stack([x0 + s0 * O.ce_normals(1729, 0, int(i), n) for i in el])
    mu, sg, it, _ = run.state()
    close(mu, X.mean(axis=0), 5.0, "mu")
    close(sg, X.std(axis=0, ddof=1), 5.0, "sigma")
    # sampled costs against the oracle's objective
    for j in np.random.default_rng(1).integers(0, N, 64):
        x = x0 + s0 * O.ce_normals(1729, 0, int(j), n)
        assert abs(costs[j] - O.objective(O.RASTRIGIN, x)) <= 1e-10 * costs[j]
    run.step(3)
    mu2, sg2, it2, _ = run.state()
    assert it2 == 4 and sg2.max() < sg.max() < 2.0                  # the distribution keeps contracting


def test_small_kernel_replay_is_bit_exact(ctx):
    """The single-CTA kernel sums the elites in cost order like the reference: with replayed normals and a
    transcendental-free objective mu and sigma are bit-identical to the oracle (and to the golden fixtures)."""
    from conftest import assert_bit_equal
    for g in golden("ce_golden.json"):
        z = np.array([[[from_bits(b) for b in row] for row in it] for it in g["normals"]])
        ce = (E.CrossEntropy.new(g["objective"]).with_sample_size(g["N"]).with_importance_selection_size(g["k"])
              .with_iter_max(g["iter_max"]).with_std_dev(g["sigma0"]).with_context(ctx)
              .with_mean_divisor(E.MeanDivisor.REFERENCE_TEN if g["divisor_ten"] else E.MeanDivisor.ELITE_COUNT))
        run = ce.start(g["x0"])
        for t, tr in enumerate(g["trace"]):
            run.step(1, z[t:t + 1])
            mu, sg, it, _ = run.state()
            assert_bit_equal(mu, [from_bits(b) for b in tr["mu"]], f"mu pass {t}")
            assert_bit_equal(sg, [from_bits(b) for b in tr["sigma"]], f"sigma pass {t}")
            assert run.last_step_ms()[1] == 1      # one launch


@pytest.mark.parametrize("objective", [E.Objective.ROSENBROCK, E.Objective.RASTRIGIN, E.Objective.HEAVY])
def test_small_kernel_equals_phase_pipeline(ctx, objective, monkeypatch):
    n, N, k, iters = 9, 777, 31, 6
    x0, s0 = np.full(n, 3.0), np.full(n, 2.0)
    mk = lambda: (E.CrossEntropy.new(objective).with_sample_size(N).with_importance_selection_size(k)
                  .with_iter_max(iters + 1).with_std_dev(s0).with_mean_divisor(E.MeanDivisor.ELITE_COUNT).with_seed(5)
                  .with_tol(0.0).with_context(ctx))
    small = mk().start(x0)
    monkeypatch.setenv("EQS_CE_NO_SMALL", "1")
    big = mk().start(x0)
    for t in range(iters):
        monkeypatch.delenv("EQS_CE_NO_SMALL", raising=False)
        small.step(1)
        monkeypatch.setenv("EQS_CE_NO_SMALL", "1")
        big.step(1)
        assert small.elites().tolist() == big.elites().tolist()
        np.testing.assert_allclose(small.costs(), big.costs(), rtol=1e-10)
        close(small.state()[0], big.state()[0], 5.0, f"mu pass {t}")
        close(small.state()[1], big.state()[1], 5.0, f"sigma pass {t}")
        assert small.state()[2:] == big.state()[2:]
    assert small.last_step_ms()[1] == 1 and big.last_step_ms()[1] == 8


def test_seeded_and_replay_solves_are_public_api(ctx):
    """CrossEntropy.solve_with_seed / solve_with_draws."""
    n, N, k, iters = 4, 500, 10, 8
    x0 = np.full(n, 3.0)
    ce = E.CrossEntropy.new(E.Objective.SPHERE).with_sample_size(N).with_importance_selection_size(k) \
        .with_iter_max(iters).with_tol(0.0).with_std_dev(np.full(n, 2.0)).with_context(ctx)
    a = ce.solve_with_seed(x0, 5)
    assert np.array_equal(a, ce.solve_with_seed(x0, 5)) and not np.array_equal(a, ce.solve_with_seed(x0, 6))
    z = np.random.Generator(np.random.Philox(3)).standard_normal((iters - 1, N, n))
    got = ce.solve_with_draws(x0, z)
    o = O.Ce(O.SPHERE, x0, sample_size=N, elite_count=k, iter_max=iters, tol=0.0, std_dev=np.full(n, 2.0),
             replay_normals=z)
    mu = o.solve()
    np.testing.assert_allclose(got, mu, rtol=1e-12, atol=1e-13)


def test_full_size_c5_per_gpu_properties(ctx):
    """BASELINE config 5 at its per-GPU size (Rosenbrock n=1024, 2^23 samples, 1 % elites): size-independent checks."""
    n, N = 1024, 1 << 23
    k = N // 100
    x0, s0 = np.zeros(n), np.full(n, 2.0)
    ce = (E.CrossEntropy.new(E.Objective.ROSENBROCK).with_sample_size(N).with_importance_selection_size(k)
          .with_iter_max(6).with_std_dev(s0).with_mean_divisor(E.MeanDivisor.ELITE_COUNT).with_seed(1729)
          .with_tol(0.0).with_context(ctx))
    run = ce.start(x0)
    run.step(1)
    costs, el = run.costs(), run.elites()
    assert el.size == k and np.all(np.diff(el) > 0)
    mask = np.zeros(N, dtype=bool); mask[el] = True
    assert costs[mask].max() <= costs[~mask].min()
    assert costs[mask].max() == np.partition(costs, k - 1)[k - 1]
    for j in np.random.default_rng(2).integers(0, N, 16):             # sampled costs against the oracle
        x = x0 + s0 * O.ce_normals(1729, 0, int(j), n)
        assert abs(costs[j] - O.objective(O.ROSENBROCK, x)) <= 1e-12 * costs[j]
    X = np.stack([x0 + s0 * O.ce_normals(1729, 0, int(i), n) for i in el])   # refit on the host from the counters
    mu, sg, it, _ = run.state()
    close(mu, X.mean(axis=0), 5.0, "mu")
    close(sg, X.std(axis=0, ddof=1), 5.0, "sigma")
    assert it == 1
    run.close()
