"""GPU: the CrossEntropy select / compaction / moment pipeline (csrc/ce.cu) on the inputs that stress it.

The pass is eight launches whose stages are joined by the last CTA of each kernel (range of the cost keys -> radix levels
over that range -> candidate list -> elite list in sample order -> two moment passes).  Checked here, through the C ABI
and against the CPU oracle: the fused pass equals the phased pass bit for bit; the multi-kernel path on sizes the
single-CTA kernel normally takes; key distributions that defeat a naive radix select (all keys equal, a handful of
distinct keys, costs that cluster more and more tightly as the solver converges, +inf costs); and the error semantics
of cross_entropy.rs:255-276 (a NaN cost on ONE rank stops EVERY rank in the same pass; a non-finite sigma is an error
only if another pass would sample with it).
"""
import os

import numpy as np
import pytest

import eqsolver_b200 as E
from conftest import assert_bit_equal
from oracle import oracle as O

pytestmark = pytest.mark.gpu


class env:
    """Environment switches the library reads when a solver is created / stepped."""

    def __init__(self, **kv):
        self.kv = kv

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        os.environ.update(self.kv)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def builder(ctx, obj, N, k, iters, s0, seed=1729):
    return (E.CrossEntropy.new(obj).with_sample_size(N).with_importance_selection_size(k).with_iter_max(iters + 1)
            .with_std_dev(s0).with_mean_divisor(E.MeanDivisor.ELITE_COUNT).with_seed(seed).with_tol(0.0).with_context(ctx))


@pytest.mark.parametrize("objective,n,N,k", [(E.Objective.ROSENBROCK, 7, 20011, 200), (E.Objective.RASTRIGIN, 12, 70001, 701),
                                             (E.Objective.USER, 3, 4099, 41)])
def test_fused_pass_equals_phased_pass_bit_for_bit(ctx, objective, n, N, k):
    """One rank: the last-CTA stages (default) and the same device functions run as ten phases with the exchange outside
    (EQS_CE_PHASED, what virtual ranks and the NCCL fallback use) must agree in every bit, pass after pass."""
    x0, s0, iters = np.full(n, 3.0), np.full(n, 2.0), 5
    fused = builder(ctx, objective, N, k, iters, s0).start(x0)
    with env(EQS_CE_PHASED="1"):
        phased = builder(ctx, objective, N, k, iters, s0).start(x0)
    for t in range(iters):
        fused.step(1)
        phased.step(1)
        (mu, sg, it, act), (pmu, psg, pit, pact) = fused.state(), phased.state()
        assert_bit_equal(mu, pmu, f"mu pass {t}")
        assert_bit_equal(sg, psg, f"sigma pass {t}")
        assert (it, act) == (pit, pact)
        assert fused.elites().tolist() == phased.elites().tolist()
    assert fused.last_step_ms()[1] == 8          # launches of one fused pass
    fused.close(); phased.close()


def test_multi_kernel_path_on_small_sizes(ctx):
    """EQS_CE_NO_SMALL routes sizes the single-CTA kernel normally takes through the eight-launch pass: fewer samples than
    one tile, than one warp, k = N, k = 2."""
    with env(EQS_CE_NO_SMALL="1"):
        for (N, k, n) in [(2, 2, 1), (3, 2, 2), (5, 5, 3), (31, 7, 2), (33, 33, 1), (257, 100, 4), (1023, 2, 5),
                          (1024, 1024, 2), (1025, 513, 9), (4097, 33, 1)]:
            x0, s0 = np.full(n, 1.0), np.full(n, 0.5)
            run = builder(ctx, E.Objective.SPHERE, N, k, 3, s0, seed=N).start(x0)
            o = O.Ce(O.SPHERE, x0, sample_size=N, elite_count=k, iter_max=4, std_dev=s0, mean_divisor=O.DIV_ELITE_COUNT,
                     seed=N, tol=0.0)
            for t in range(3):
                run.step(1)
                assert o.step() == 0
                assert sorted(run.elites().tolist()) == sorted(o.elites().tolist()), (N, k, n, t)
                np.testing.assert_allclose(run.state()[0], o.state()[0], rtol=1e-12, atol=2e-12, err_msg=str((N, k, n)))
                np.testing.assert_allclose(run.state()[1], o.state()[1], rtol=1e-12, atol=2e-12, err_msg=str((N, k, n)))
            assert run.last_step_ms()[1] == 8
            run.close()


def replay_case(ctx, obj, z, x0, s0, k, passes=1):
    """Replayed normals.  Pass 0 has bit-identical inputs on both sides, so the costs of a transcendental-free objective
    and therefore the elite set equal the oracle's exactly.  From then on mu and sigma differ in their last bits (the sums
    are taken in another order), so every pass is checked against the definition applied to the GPU's OWN numbers: the
    elite set is the k smallest of the costs it produced (ties: lowest index), and mu / sigma are the moments of
    x = mu + sigma z over that set."""
    N, n = z.shape[1], z.shape[2]
    run = (E.CrossEntropy.new(obj).with_sample_size(N).with_importance_selection_size(k).with_iter_max(passes + 1)
           .with_std_dev(s0).with_mean_divisor(E.MeanDivisor.ELITE_COUNT).with_tol(0.0).with_context(ctx)).start(x0)
    o = O.Ce(int(obj), x0, sample_size=N, elite_count=k, iter_max=passes + 1, std_dev=s0, mean_divisor=O.DIV_ELITE_COUNT,
             tol=0.0, replay_normals=z[:1])
    for t in range(passes):
        mu0, sg0, _, _ = run.state()
        run.step(1, z[t:t + 1])
        costs = run.costs()
        got = np.sort(run.elites())
        if t == 0:
            assert o.step() == 0
            assert_bit_equal(costs, o.costs(), "costs pass 0")
            assert got.tolist() == sorted(o.elites().tolist()), "elite set pass 0 (oracle)"
        want = np.sort(np.lexsort((np.arange(N), costs))[:k])      # by cost, then by index
        assert got.tolist() == want.tolist(), f"elite set pass {t}"
        x = mu0 + sg0 * z[t, want]                                   # the elites' coordinates, regenerated on the host
        mu = x.sum(axis=0) / k
        sg = np.sqrt(((x - mu) ** 2).sum(axis=0) / (k - 1))
        scale = float(np.abs(mu).max() + np.abs(sg).max() + 1e-300)
        np.testing.assert_allclose(run.state()[0], mu, rtol=1e-11, atol=1e-12 * scale, err_msg=f"mu pass {t}")
        np.testing.assert_allclose(run.state()[1], sg, rtol=1e-9, atol=1e-12 * scale, err_msg=f"sigma pass {t}")
    run.close()


def test_few_distinct_keys_and_all_keys_equal(ctx):
    """Massive ties: the key range collapses (all equal: one level, every sample a tie) or holds a handful of values
    (the bucket of the k-th key keeps thousands of candidates down to the last level)."""
    rng = np.random.default_rng(11)
    N, n = 6000, 2
    z = rng.choice([-1.0, 0.0, 0.5, 1.0], size=(1, N, n))                 # at most 9 distinct Sphere costs
    for k in (2, 777, 3000, 5999, 6000):
        replay_case(ctx, E.Objective.SPHERE, z, np.zeros(n), np.ones(n), k)
    z1 = np.repeat(rng.standard_normal((1, 1, n)), N, axis=1)            # one cost for everybody
    for k in (2, 100, 6000):
        replay_case(ctx, E.Objective.SPHERE, z1, np.zeros(n), np.ones(n), k)
    # one outlier stretches the range over hundreds of binades: everything else lands in a single bin of level A
    z2 = rng.standard_normal((1, N, n)) * 1e-3
    z2[0, 4321, :] = 1e150
    replay_case(ctx, E.Objective.SPHERE, z2, np.zeros(n), np.ones(n), 60)
    # +inf costs are ordinary keys (they sort last); with k > the finite ones, infinities are elite, lowest index first
    z3 = rng.standard_normal((1, N, n))
    z3[0, ::3, 0] = 1e200                                                  # x^2 overflows
    N3 = z3.shape[1]
    run = (E.CrossEntropy.new(E.Objective.SPHERE).with_sample_size(N3).with_importance_selection_size(N3 - 1500)
           .with_iter_max(2).with_mean_divisor(E.MeanDivisor.ELITE_COUNT).with_tol(0.0).with_context(ctx)).start(np.zeros(n))
    o = O.Ce(O.SPHERE, np.zeros(n), sample_size=N3, elite_count=N3 - 1500, iter_max=2, mean_divisor=O.DIV_ELITE_COUNT,
             tol=0.0, replay_normals=z3)
    run.step(1, z3)
    assert o.step() == 0
    assert sorted(run.elites().tolist()) == sorted(o.elites().tolist())
    run.close()


def test_costs_cluster_as_the_solver_converges(ctx):
    """Thirty replayed passes of Sphere: sigma falls by orders of magnitude, the costs of a pass end up within a few
    thousand ulps of each other, and the elite set must still be exactly the oracle's in every pass."""
    rng = np.random.default_rng(5)
    N, n, k, passes = 8192, 3, 82, 30
    z = rng.standard_normal((passes, N, n))
    replay_case(ctx, E.Objective.SPHERE, z, np.full(n, 1.0), np.full(n, 2.0), k, passes)


def virtual_pass(ranks, until=None):
    """One phased pass over virtual ranks; returns after phase `until` (inclusive) if given."""
    for ph in range(E.CeRun.N_PHASES):
        for r in ranks:
            r.phase(ph)
        if ranks[0].exchange_words(ph):
            words = np.stack([r.get_exchange(ph) for r in ranks])
            for r in ranks:
                r.set_exchange(ph, words)
        if until is not None and ph == until:
            return


def test_nan_cost_on_one_rank_stops_every_rank_in_the_same_pass(ctx):
    """SERIES(x) = sum (-1)^i x^i is NaN for x > 2.03 (inf - inf).  With mu = 0, sigma = 0.6 that is a 3.4 sigma event:
    find a seed where exactly ONE of three virtual ranks draws such a sample.  Every rank must then report NotANumber
    for this pass with mu, sigma and the iteration count untouched -- the rank that saw nothing as well (before, it went
    on and waited for the failed rank in the next exchange)."""
    n, N, k, world = 1, 3000, 30, 3
    x0, s0 = np.zeros(n), np.full(n, 0.6)
    for seed in range(200):
        ranks = [builder(ctx, E.Objective.SERIES, N, k, 5, s0, seed=seed).with_virtual_rank(r, world).start(x0)
                 for r in range(world)]
        for r in ranks:
            r.phase(0)
        nan_ranks = [int(np.isnan(r.costs()).any()) for r in ranks]
        if sum(nan_ranks) != 1:
            for r in ranks:
                r.close()
            continue
        # finish phase 0's exchange and run the rest of the pass
        words = np.stack([r.get_exchange(0) for r in ranks])
        for r in ranks:
            r.set_exchange(0, words)
        for ph in range(1, E.CeRun.N_PHASES):
            for r in ranks:
                r.phase(ph)
            if ranks[0].exchange_words(ph):
                words = np.stack([r.get_exchange(ph) for r in ranks])
                for r in ranks:
                    r.set_exchange(ph, words)
        for r in ranks:
            with pytest.raises(E.NotANumber):
                r.state()
            r.close()
        # one real rank: the same seed fails in solve() and in the oracle
        with pytest.raises(E.NotANumber):
            builder(ctx, E.Objective.SERIES, N, k, 5, s0, seed=seed).solve(x0)
        with pytest.raises(FloatingPointError):
            O.Ce(O.SERIES, x0, sample_size=N, elite_count=k, iter_max=6, std_dev=s0, mean_divisor=O.DIV_ELITE_COUNT,
                 seed=seed, tol=0.0).solve()
        return
    pytest.fail("no seed with exactly one NaN rank among 200")


@pytest.mark.parametrize("N", [64, 5000])   # the single-CTA kernel and the multi-kernel pass
def test_non_finite_sigma_is_an_error_only_if_the_loop_goes_on(ctx, N):
    """cross_entropy.rs:255-259: Normal::new(mu, sigma).unwrap() is what panics, at the top of the NEXT pass.  Elites
    around 1e200 make sum (x - mu)^2 overflow: sigma = inf after the first pass.  iter_max = 2 ends the loop there ->
    Ok(mus); iter_max = 3 samples again -> the reference panics, here NotANumber.  Same on the oracle."""
    n, k = 2, 10
    x0, s0 = np.full(n, 1e200), np.full(n, 1e200)
    mk = lambda iter_max: (E.CrossEntropy.new(E.Objective.SPHERE).with_sample_size(N).with_importance_selection_size(k)
                           .with_iter_max(iter_max).with_std_dev(s0).with_mean_divisor(E.MeanDivisor.ELITE_COUNT)
                           .with_seed(3).with_context(ctx))
    got = mk(2).solve(x0)
    want = O.Ce(O.SPHERE, x0, sample_size=N, elite_count=k, iter_max=2, std_dev=s0, mean_divisor=O.DIV_ELITE_COUNT,
                seed=3).solve()
    assert np.isfinite(got).all()
    np.testing.assert_allclose(got, want, rtol=1e-12)
    with pytest.raises(E.NotANumber):
        mk(3).solve(x0)
    with pytest.raises(FloatingPointError):
        O.Ce(O.SPHERE, x0, sample_size=N, elite_count=k, iter_max=3, std_dev=s0, mean_divisor=O.DIV_ELITE_COUNT,
             seed=3).solve()
