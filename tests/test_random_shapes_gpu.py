"""GPU: a seeded sweep over random shapes (dimension, population, objective, mode, seed) against the CPU oracle -- the
ragged corners (n not a multiple of 4, populations around the tile / window / single-CTA limits) that fixed-size tests
only sample."""
import os

import numpy as np
import pytest

import eqsolver_b200 as E
from conftest import assert_bit_equal
from oracle import oracle as O

pytestmark = pytest.mark.gpu

# EQS_SWEEP_SEED=k shifts every seed below: the same sweeps over other random shapes (bug hunting beyond the fixed set)
SWEEP = int(os.environ.get("EQS_SWEEP_SEED", "0"))

EXACT = [E.Objective.SPHERE, E.Objective.ROSENBROCK, E.Objective.USER, E.Objective.HEAVY]


def _pick_population(rng):
    edge = [1, 2, 31, 32, 33, 255, 256, 257, 2047, 2048, 2049, 4096, 4097]
    return int(rng.choice(edge)) if rng.random() < 0.5 else int(rng.integers(1, 7000))


def test_particle_swarm_random_shapes(ctx):
    rng = np.random.default_rng(20261019 + SWEEP)
    for case in range(400):
        obj = EXACT[int(rng.integers(0, len(EXACT)))]
        n = int(rng.integers(1, 41)) if obj != E.Objective.HEAVY else int(rng.integers(1, 13))
        N = _pick_population(rng)
        mode = E.PsoMode.SYNCHRONOUS if rng.random() < 0.5 else E.PsoMode.SEQUENTIAL_EXACT
        seed = int(rng.integers(0, 2 ** 40))
        lo, hi = -float(rng.uniform(0.5, 6.0)), float(rng.uniform(0.5, 12.0))
        x0 = rng.uniform(lo, hi, n)
        # sequential mode beyond the single-CTA kernel: force small windows half of the time (many rounds per pass)
        window = str(int(rng.integers(16, 3000))) if rng.random() < 0.5 else None
        if window is None:
            os.environ.pop("EQS_PSO_SEQ_WINDOW", None)
        else:
            os.environ["EQS_PSO_SEQ_WINDOW"] = window
        what = f"case {case}: {obj.name} n={n} N={N} {mode.name} seed={seed} window={window}"
        g = (E.ParticleSwarm.new(obj, [lo] * n, [hi] * n).with_particle_count(N).with_tol(0.0).with_seed(seed)
             .with_mode(mode).with_context(ctx))
        o = O.Pso(int(obj), [lo] * n, [hi] * n, particle_count=N, tol=0.0, mode=int(mode), seed=seed)
        run = g.start(x0)
        o.init(x0)
        for k in (1, 2):
            run.step(k)
            for _ in range(k):
                assert o.step() == 0
        for a, b, name in zip(run.state(), o.state(), ("x", "v", "pbest", "pbest cost")):
            assert_bit_equal(a, b, f"{what}: {name}")
        gb, gc, it, _, _ = run.gbest()
        assert_bit_equal(gb, o.gbest(), what + ": gbest")
        assert gc == o.gbest_cost() and it == o.iters() == 3, what
        assert_bit_equal(run.ring()[0], o.ring()[0], what + ": ring")
        assert run.ring()[1] == o.ring()[1], what
        run.close()
    os.environ.pop("EQS_PSO_SEQ_WINDOW", None)


def test_cross_entropy_random_shapes(ctx):
    rng = np.random.default_rng(7 + SWEEP)
    for case in range(150):
        obj = [E.Objective.SPHERE, E.Objective.ROSENBROCK, E.Objective.USER][int(rng.integers(0, 3))]
        n = int(rng.integers(1, 33))
        N = max(2, _pick_population(rng))
        k = int(rng.integers(2, min(N, 300) + 1))
        seed = int(rng.integers(0, 2 ** 40))
        div = E.MeanDivisor.ELITE_COUNT if rng.random() < 0.7 else E.MeanDivisor.REFERENCE_TEN
        x0, s0 = rng.uniform(-2, 2, n), rng.uniform(0.5, 2.0, n)
        what = f"case {case}: {obj.name} n={n} N={N} k={k} {div.name} seed={seed}"
        ce = (E.CrossEntropy.new(obj).with_sample_size(N).with_importance_selection_size(k).with_iter_max(4)
              .with_std_dev(s0).with_mean_divisor(div).with_seed(seed).with_tol(0.0).with_context(ctx))
        o = O.Ce(int(obj), x0, sample_size=N, elite_count=k, iter_max=4, std_dev=s0, mean_divisor=int(div), seed=seed,
                 tol=0.0)
        run = ce.start(x0)
        for t in range(2):
            run.step(1)
            assert o.step() == 0
            assert sorted(run.elites().tolist()) == sorted(o.elites().tolist()), what + f": elite set, pass {t}"
            mu, sg, _, _ = run.state()
            omu, osg = o.state()
            scale = float(np.abs(omu).max() + np.abs(osg).max() + 1.0)
            np.testing.assert_allclose(mu, omu, rtol=1e-11, atol=1e-12 * scale, err_msg=what)
            np.testing.assert_allclose(sg, osg, rtol=1e-11, atol=1e-12 * scale, err_msg=what)
        run.close()


def test_integrators_random_shapes(ctx):
    rng = np.random.default_rng(11 + SWEEP)
    for case in range(100):
        integrand = [E.Objective.SPHERE, E.Objective.ROSENBROCK, E.Objective.GAUSSIAN][int(rng.integers(0, 3))]
        n = int(rng.integers(1, 10))
        count = int(rng.choice([2, 3, 255, 256, 257, 1000, 65536, 65537])) if rng.random() < 0.5 else int(rng.integers(2, 90000))
        lower = rng.uniform(-2, 0, n)
        upper = lower + rng.uniform(0.1, 3.0, n)
        seed = int(rng.integers(0, 2 ** 40))
        what = f"case {case}: {integrand.name} n={n} count={count} seed={seed}"
        r = E.MonteCarlo.new(integrand).with_sample_count(count).with_context(ctx).integrate_with_seed(lower, upper, seed)
        st, mean, var = O.mc_integrate(int(integrand), lower, upper, count, seed)
        assert st == 0, what
        assert abs(r.mean - mean) <= 1e-12 * max(1.0, abs(mean)), what
        assert abs(r.variance - var) <= 1e-9 * max(1e-300, abs(var)), what
    for case in range(60):
        integrand = [E.Objective.SPHERE, E.Objective.GAUSSIAN, E.Objective.ROSENBROCK][int(rng.integers(0, 3))]
        n = int(rng.integers(1, 5))
        kw = dict(sample_count=int(rng.integers(10, 30000)), maximum_depth=int(rng.integers(0, 7)),
                  min_dim_count=int(rng.integers(1, 64)), alpha=float(rng.uniform(0.5, 3.0)),
                  dither=float(rng.uniform(0.0, 0.3)))
        lower = rng.uniform(-1, 0, n)
        upper = lower + rng.uniform(0.5, 2.0, n)
        seed = int(rng.integers(0, 2 ** 40))
        what = f"MISER case {case}: {integrand.name} n={n} {kw} seed={seed}"
        mi = (E.MISER.new(integrand).with_sample_count(kw["sample_count"]).with_maximum_depth(kw["maximum_depth"])
              .with_minimum_dimensional_sample_count(kw["min_dim_count"]).with_alpha(kw["alpha"]).with_dither(kw["dither"])
              .with_context(ctx))
        st, mean, var, evals = O.miser_integrate(int(integrand), lower, upper, seed=seed, **kw)
        if st != 0:   # e.g. 1-D Rosenbrock is constant: 0/0 points -> TypeConversionError, like the reference
            with pytest.raises({O.ERR_TYPE_CONVERSION: E.TypeConversionError, O.ERR_EXTERNAL: E.ExternalError,
                                O.ERR_INCORRECT_INPUT: E.IncorrectInput}[st]):
                mi.integrate_with_seed(lower, upper, seed)
            continue
        got = mi.integrate_with_seed(lower, upper, seed)
        assert mi.last_report.evals == evals, what
        assert abs(got.mean - mean) <= 1e-12 * max(1.0, abs(mean)), what
        assert abs(got.variance - var) <= 1e-9 * max(1e-300, abs(var)), what


def test_sharded_particle_swarm_random_shapes(ctx):
    """The multi-rank protocol (virtual ranks on one GPU, records moved by the harness) on random shapes, including
    worlds with more ranks than particles: every rank's gbest / ring equals the unsharded oracle bit for bit, and the
    concatenated shards are its population."""
    rng = np.random.default_rng(5 + SWEEP)
    for case in range(40):
        obj = EXACT[int(rng.integers(0, 3))]
        n = int(rng.integers(1, 20))
        world = int(rng.integers(2, 7))
        N = int(rng.choice([1, 2, world - 1, world, world + 1, 63, 64, 65, 257])) if rng.random() < 0.5 else int(rng.integers(1, 1500))
        N = max(N, 1)
        seed = int(rng.integers(0, 2 ** 40))
        lower, upper = [-5.0] * n, [10.0] * n
        x0 = rng.uniform(-5, 10, n)
        what = f"case {case}: {obj.name} n={n} N={N} world={world} seed={seed}"
        base = lambda: (E.ParticleSwarm.new(obj, lower, upper).with_particle_count(N).with_tol(0.0).with_seed(seed)
                        .with_mode(E.PsoMode.SYNCHRONOUS).with_context(ctx))
        o = O.Pso(int(obj), lower, upper, particle_count=N, tol=0.0, mode=O.SYNCHRONOUS, seed=seed)
        o.init(x0)
        ranks = [base().with_virtual_rank(r, world).start() for r in range(world)]
        assert sum(r.local for r in ranks) == N, what

        def gather():
            recs = np.stack([r.get_record() for r in ranks])
            for r in ranks:
                r.set_records(recs)

        for phase in (0, 1):
            for r in ranks:
                r.init_local(x0, phase)
            gather()
        for r in ranks:
            r.init_apply()
        for t in range(4):
            for r in ranks:
                g, c, it, _, _ = r.gbest()
                assert_bit_equal(g, o.gbest(), what + f": gbest pass {t}")
                assert c == o.gbest_cost() and it == o.iters(), what
                assert_bit_equal(r.ring()[0], o.ring()[0], what + ": ring")
                assert r.ring()[1] == o.ring()[1], what
            if t == 3:
                break
            assert o.step() == 0
            for r in ranks:
                r.step_local()
            gather()
            for r in ranks:
                r.step_apply()
        parts = [r.state() for r in ranks if r.local > 0]
        for j, name in enumerate(("x", "v", "pbest", "pbest cost")):
            assert_bit_equal(np.concatenate([p[j] for p in parts]), o.state()[j], what + ": " + name)
        for r in ranks:
            r.close()


def _same_numbers(a, b, what):
    """Bit equality, except that any NaN equals any NaN (x86 and the GPU differ in the payload of a generated NaN)."""
    a = np.ascontiguousarray(a, dtype=np.float64).ravel()
    b = np.ascontiguousarray(b, dtype=np.float64).ravel()
    assert a.shape == b.shape, what
    both_nan = np.isnan(a) & np.isnan(b)
    bad = np.flatnonzero((a.view(np.uint64) != b.view(np.uint64)) & ~both_nan)
    assert bad.size == 0, f"{what}: {bad.size} of {a.size} differ, first at {bad[0]}: {a[bad[0]]!r} vs {b[bad[0]]!r}"


def test_hostile_numerics_match_the_oracle(ctx):
    """NaN / infinite / huge coefficients, tolerances and start points: the GPU path must do what the reference's
    arithmetic does with them (NaN never wins a `<`, infinities propagate), never hang or crash."""
    rng = np.random.default_rng(123 + SWEEP)
    specials = [0.0, -0.0, -1.0, 2.5, 1e308, -1e308, np.inf, -np.inf, np.nan, 5e-324]
    for case in range(200):
        obj = [E.Objective.SPHERE, E.Objective.ROSENBROCK][int(rng.integers(0, 2))]
        n, N = int(rng.integers(1, 9)), int(rng.choice([1, 7, 100, 1500, 3000]))
        mode = E.PsoMode.SYNCHRONOUS if rng.random() < 0.5 else E.PsoMode.SEQUENTIAL_EXACT
        w, c1, c2, tol = (float(rng.choice(specials)) if rng.random() < 0.5 else float(rng.uniform(0, 1.5)) for _ in range(4))
        x0 = np.array([float(rng.choice(specials)) if rng.random() < 0.3 else float(rng.uniform(-3, 3)) for _ in range(n)])
        seed = int(rng.integers(0, 2 ** 32))
        what = f"case {case}: {obj.name} n={n} N={N} {mode.name} w={w} c1={c1} c2={c2} tol={tol} x0={x0.tolist()} seed={seed}"
        g = (E.ParticleSwarm.new(obj, [-2.0] * n, [3.0] * n).with_particle_count(N).with_iter_max(4).with_tol(tol)
             .with_inertia_weight(w).with_cognitive_coefficient(c1).with_social_coefficient(c2).with_seed(seed)
             .with_mode(mode).with_context(ctx))
        o = O.Pso(int(obj), [-2.0] * n, [3.0] * n, particle_count=N, iter_max=4, tol=tol, w=w, c1=c1, c2=c2,
                  mode=int(mode), seed=seed)
        run = g.start(x0)
        o.init(x0)
        for _ in range(3):
            run.step(1)
            o.step()
        for a, b, name in zip(run.state(), o.state(), ("x", "v", "pbest", "pbest cost")):
            _same_numbers(a, b, f"{what}: {name}")
        gb, gc, it, stalled, _ = run.gbest()
        _same_numbers(gb, o.gbest(), what + ": gbest")
        _same_numbers([gc], [o.gbest_cost()], what + ": gbest cost")
        assert it == o.iters(), what
        _same_numbers(run.ring()[0], o.ring()[0], what + ": ring")
        run.close()
        _same_numbers(g.solve(x0), o.solve(x0), what + ": solve")


def test_cross_entropy_hostile_numerics_match_the_oracle(ctx):
    """Zero / infinite / NaN standard deviations and start points: where the reference panics (NaN cost in the sort,
    non-finite sigma in Normal::new) the GPU path returns NotANumber, otherwise it follows the oracle."""
    rng = np.random.default_rng(321 + SWEEP)
    specials = [0.0, -0.0, -1.0, 1e308, np.inf, -np.inf, np.nan, 5e-324, 1e-300]
    for case in range(120):
        obj = [E.Objective.SPHERE, E.Objective.ROSENBROCK][int(rng.integers(0, 2))]
        n, N = int(rng.integers(1, 7)), int(rng.choice([2, 11, 100, 1024, 1500]))
        k = int(rng.integers(2, min(N, 40) + 1))
        s0 = np.array([float(rng.choice(specials)) if rng.random() < 0.3 else float(rng.uniform(0.2, 2)) for _ in range(n)])
        x0 = np.array([float(rng.choice(specials)) if rng.random() < 0.2 else float(rng.uniform(-3, 3)) for _ in range(n)])
        tol = float(rng.choice([0.0, 1e-6, np.nan, np.inf, -1.0]))
        seed = int(rng.integers(0, 2 ** 32))
        what = f"case {case}: {obj.name} n={n} N={N} k={k} sigma={s0.tolist()} x0={x0.tolist()} tol={tol} seed={seed}"
        ce = (E.CrossEntropy.new(obj).with_sample_size(N).with_importance_selection_size(k).with_iter_max(5)
              .with_std_dev(s0).with_mean_divisor(E.MeanDivisor.ELITE_COUNT).with_seed(seed).with_tol(tol).with_context(ctx))
        o = O.Ce(int(obj), x0, sample_size=N, elite_count=k, iter_max=5, std_dev=s0, mean_divisor=O.DIV_ELITE_COUNT,
                 seed=seed, tol=tol)
        try:
            want = o.solve()
        except FloatingPointError:          # the reference panics (NaN cost in the sort, non-finite sigma)
            with pytest.raises(E.NotANumber):
                ce.solve(x0)
            continue
        got = ce.solve(x0)
        # once sigma has collapsed every cost comparison is a rounding-level tie and the two sides may pick
        # different elites: the means then agree to the scale of the final sigma, not to 1e-10
        noise = float(np.nanmax(np.abs(o.state()[1]))) if np.isfinite(o.state()[1]).any() else 0.0
        with np.errstate(invalid="ignore"):   # inf - inf among the hostile values
            ok = ((np.isnan(got) & np.isnan(want)) | np.isclose(got, want, rtol=1e-10, atol=1e-10) | (got == want)
                  | (np.abs(got - want) <= noise))
        assert ok.all(), f"{what}: {got} vs {want} (final sigma {o.state()[1]})"


def test_miser_hostile_parameters_match_the_oracle(ctx):
    """alpha / dither from {0, negative, huge, inf, NaN}, degenerate boxes, tiny budgets: same status as the recursive
    oracle (ExternalError, TypeConversionError, IncorrectInput) or the same numbers."""
    rng = np.random.default_rng(777 + SWEEP)
    specials = [0.0, -0.0, -1.0, -0.5, 0.49, 0.5, 0.51, 1.0, 1e308, np.inf, -np.inf, np.nan]
    errors = {O.ERR_TYPE_CONVERSION: E.TypeConversionError, O.ERR_EXTERNAL: E.ExternalError,
              O.ERR_INCORRECT_INPUT: E.IncorrectInput}
    seen = set()
    for case in range(150):
        integrand = [E.Objective.SPHERE, E.Objective.GAUSSIAN, E.Objective.ROSENBROCK][int(rng.integers(0, 3))]
        n = int(rng.integers(1, 4))
        alpha = float(rng.choice(specials)) if rng.random() < 0.5 else float(rng.uniform(0.5, 3))
        dither = float(rng.choice(specials)) if rng.random() < 0.5 else float(rng.uniform(0, 0.3))
        kw = dict(sample_count=int(rng.choice([0, 1, 2, 50, 700, 5000])), maximum_depth=int(rng.integers(0, 5)),
                  min_dim_count=int(rng.choice([0, 1, 2, 32])), alpha=alpha, dither=dither)
        lower = rng.uniform(-1, 0, n)
        upper = lower + (0.0 if rng.random() < 0.15 else rng.uniform(0.5, 2.0, n))
        seed = int(rng.integers(0, 2 ** 32))
        what = f"case {case}: {integrand.name} n={n} {kw} box={lower.tolist()}..{np.broadcast_to(upper, (n,)).tolist()} seed={seed}"
        upper = np.broadcast_to(upper, (n,)).copy()
        mi = (E.MISER.new(integrand).with_sample_count(kw["sample_count"]).with_maximum_depth(kw["maximum_depth"])
              .with_minimum_dimensional_sample_count(kw["min_dim_count"]).with_alpha(alpha).with_dither(dither)
              .with_context(ctx))
        st, mean, var, evals = O.miser_integrate(int(integrand), lower, upper, seed=seed, **kw)
        seen.add(st)
        if st != 0:
            with pytest.raises(errors[st]):
                mi.integrate_with_seed(lower, upper, seed)
            continue
        got = mi.integrate_with_seed(lower, upper, seed)
        assert mi.last_report.evals == evals, what
        assert (np.isnan(got.mean) and np.isnan(mean)) or abs(got.mean - mean) <= 1e-12 * max(1.0, abs(mean)), what
        assert (np.isnan(got.variance) and np.isnan(var)) or abs(got.variance - var) <= 1e-9 * max(1e-300, abs(var)), what
    assert {0, O.ERR_EXTERNAL, O.ERR_TYPE_CONVERSION} <= seen   # the sweep did reach the error paths


def test_sharded_cross_entropy_random_shapes(ctx):
    """The CrossEntropy exchange protocol (virtual ranks on one GPU: histograms for the distributed select, moment
    sums in rank order) on random shapes incl. worlds larger than the sample count and ties at the k-th cost."""
    rng = np.random.default_rng(2024 + SWEEP)
    for case in range(30):
        obj = [E.Objective.SPHERE, E.Objective.ROSENBROCK][int(rng.integers(0, 2))]
        n = int(rng.integers(1, 12))
        world = int(rng.integers(2, 6))
        N = int(rng.choice([2, 3, world, world + 1, 64, 65, 1025])) if rng.random() < 0.5 else int(rng.integers(2, 3000))
        N = max(N, 2)
        k = int(rng.integers(2, min(N, 200) + 1))
        seed = int(rng.integers(0, 2 ** 40))
        x0, s0 = rng.uniform(-2, 2, n), rng.uniform(0.5, 2.0, n)
        what = f"case {case}: {obj.name} n={n} N={N} k={k} world={world} seed={seed}"
        base = lambda: (E.CrossEntropy.new(obj).with_sample_size(N).with_importance_selection_size(k).with_iter_max(4)
                        .with_std_dev(s0).with_mean_divisor(E.MeanDivisor.ELITE_COUNT).with_seed(seed).with_tol(0.0)
                        .with_context(ctx))
        o = O.Ce(int(obj), x0, sample_size=N, elite_count=k, iter_max=4, std_dev=s0, mean_divisor=O.DIV_ELITE_COUNT,
                 seed=seed, tol=0.0)
        ranks = [base().with_virtual_rank(r, world).start(x0) for r in range(world)]
        assert sum(r.local for r in ranks) == N, what
        for t in range(2):
            assert o.step() == 0
            for ph in range(E.CeRun.N_PHASES):
                for r in ranks:
                    r.phase(ph)
                if ranks[0].exchange_words(ph):
                    words = np.stack([r.get_exchange(ph) for r in ranks])
                    for r in ranks:
                        r.set_exchange(ph, words)
            el = np.sort(np.concatenate([r.elites() for r in ranks]))
            assert el.tolist() == np.sort(o.elites()).tolist(), what + f": elite set, pass {t}"
            omu, osg = o.state()
            scale = float(np.abs(omu).max() + np.abs(osg).max() + 1.0)
            for r in ranks:
                mu, sg, it, act = r.state()
                np.testing.assert_allclose(mu, omu, rtol=1e-11, atol=1e-12 * scale, err_msg=what)
                np.testing.assert_allclose(sg, osg, rtol=1e-11, atol=1e-12 * scale, err_msg=what)
        for r in ranks:
            r.close()


def test_particle_swarm_larger_random_shapes(ctx):
    """A few populations beyond one wave of the grid (75 776 particles): natural multi-window sequential rounds,
    several grid-stride iterations per thread in the synchronous kernel, ragged dimensions."""
    os.environ.pop("EQS_PSO_SEQ_WINDOW", None)
    rng = np.random.default_rng(99 + SWEEP)
    for case in range(8):
        obj = EXACT[int(rng.integers(0, 3))]
        n = int(rng.integers(1, 40))
        N = int(rng.integers(76_000, 400_000))
        mode = E.PsoMode.SYNCHRONOUS if case % 2 else E.PsoMode.SEQUENTIAL_EXACT
        seed = int(rng.integers(0, 2 ** 40))
        x0 = rng.uniform(-3, 3, n)
        what = f"case {case}: {obj.name} n={n} N={N} {mode.name} seed={seed}"
        g = (E.ParticleSwarm.new(obj, [-5.0] * n, [10.0] * n).with_particle_count(N).with_tol(0.0).with_seed(seed)
             .with_mode(mode).with_context(ctx))
        o = O.Pso(int(obj), [-5.0] * n, [10.0] * n, particle_count=N, tol=0.0, mode=int(mode), seed=seed)
        run = g.start(x0)
        o.init(x0)
        run.step(3)
        for _ in range(3):
            assert o.step() == 0
        for a, b, name in zip(run.state(), o.state(), ("x", "v", "pbest", "pbest cost")):
            assert_bit_equal(a, b, f"{what}: {name}")
        gb, gc, it, _, imp = run.gbest()
        assert_bit_equal(gb, o.gbest(), what + ": gbest")
        assert gc == o.gbest_cost() and it == 3, what
        assert_bit_equal(run.ring()[0], o.ring()[0], what + ": ring")
        run.close()
