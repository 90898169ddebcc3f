#!/usr/bin/env python3
"""Small-shape pass over every kernel family of the library, meant to run under compute-sanitizer:

    compute-sanitizer --tool memcheck  --error-exitcode 1 python tests/sanitizer_smoke.py
    compute-sanitizer --tool racecheck --error-exitcode 1 python tests/sanitizer_smoke.py

(one tool per gpurun call; B200_PROFILING.md).  The kernels use last-block tickets, a grid-wide barrier, role-swapped
buffers, shared-memory tables and histograms -- what memcheck (out-of-bounds, misaligned 256-bit accesses) and racecheck
(shared-memory hazards) exist for.  Shapes are small because the tools slow a kernel down by 10-100x, but chosen so that
every grid has several CTAs, ragged tails (n % 4 != 0, N % 32 != 0) and at least one gbest improvement / tie.
Every result is also compared with the CPU oracle, so a run is a parity check as well.  Summaries: profiles/r02_sanitizer.md.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import eqsolver_b200 as E  # noqa: E402
from oracle import oracle as O  # noqa: E402


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def main():
    ctx = E.Context(0)
    done = []

    # ParticleSwarm: synchronous step kernel (several CTAs, ragged n and N), init + ring scan
    for (obj, n, N) in ((E.Objective.ROSENBROCK, 7, 3001), (E.Objective.USER, 8, 2600), (E.Objective.RASTRIGIN, 5, 2500)):
        lo, hi, x0 = [-5.0] * n, [10.0] * n, np.full(n, 3.0)
        run = (E.ParticleSwarm.new(obj, lo, hi).with_particle_count(N).with_tol(0.0).with_seed(3)
               .with_mode(E.PsoMode.SYNCHRONOUS).with_context(ctx)).start(x0)
        o = O.Pso(int(obj), lo, hi, particle_count=N, tol=0.0, mode=O.SYNCHRONOUS, seed=3)
        o.init(x0)
        run.step(3)
        for _ in range(3):
            o.step()
        if obj != E.Objective.RASTRIGIN:
            assert np.array_equal(bits(run.gbest()[0]), bits(o.gbest())), "sync gbest"
            assert np.array_equal(bits(run.state()[0]), bits(o.state()[0])), "sync x"
        run.close()
        done.append(f"pso sync {obj.name}")

    # sequential-exact: the persistent cooperative kernel (grid barrier, slots), and the single-CTA kernels
    for N in (2600, 300):
        n = 6
        lo, hi, x0 = [-5.0] * n, [10.0] * n, np.full(n, 3.0)
        run = (E.ParticleSwarm.new(E.Objective.ROSENBROCK, lo, hi).with_particle_count(N).with_tol(0.0).with_seed(5)
               .with_mode(E.PsoMode.SEQUENTIAL_EXACT).with_context(ctx)).start(x0)
        o = O.Pso(O.ROSENBROCK, lo, hi, particle_count=N, tol=0.0, mode=O.SEQUENTIAL, seed=5)
        o.init(x0)
        run.step(3)
        for _ in range(3):
            o.step()
        assert np.array_equal(bits(run.gbest()[0]), bits(o.gbest())), "seq gbest"
        for a, b in zip(run.state(), o.state()):
            assert np.array_equal(bits(a), bits(b)), "seq state"
        run.close()
        done.append(f"pso sequential N={N}")

    # CrossEntropy: the 8-launch pass (select levels, candidate list, count / compact, moments), the single-CTA kernel,
    # ties at the k-th cost, and the phased pass over virtual ranks
    os.environ["EQS_CE_NO_SMALL"] = "1"
    for (N, k, n) in ((5001, 77, 6), (700, 700, 3), (2049, 2, 5)):
        x0, s0 = np.full(n, 3.0), np.full(n, 2.0)
        run = (E.CrossEntropy.new(E.Objective.ROSENBROCK).with_sample_size(N).with_importance_selection_size(k)
               .with_iter_max(4).with_std_dev(s0).with_mean_divisor(E.MeanDivisor.ELITE_COUNT).with_seed(11).with_tol(0.0)
               .with_context(ctx)).start(x0)
        o = O.Ce(O.ROSENBROCK, x0, sample_size=N, elite_count=k, iter_max=4, std_dev=s0, mean_divisor=O.DIV_ELITE_COUNT,
                 seed=11, tol=0.0)
        for _ in range(2):
            run.step(1)
            assert o.step() == 0
            assert sorted(run.elites().tolist()) == sorted(o.elites().tolist()), "ce elites"
        np.testing.assert_allclose(run.state()[0], o.state()[0], rtol=1e-11, atol=1e-11)
        run.close()
        done.append(f"ce fused N={N} k={k}")
    del os.environ["EQS_CE_NO_SMALL"]
    z = np.repeat(np.random.default_rng(3).standard_normal((1, 1500, 2)), 4, axis=1)   # every cost four times: ties
    run = (E.CrossEntropy.new(E.Objective.SPHERE).with_sample_size(6000).with_importance_selection_size(1001)
           .with_iter_max(3).with_mean_divisor(E.MeanDivisor.ELITE_COUNT).with_tol(0.0).with_context(ctx)).start(np.zeros(2))
    o = O.Ce(O.SPHERE, np.zeros(2), sample_size=6000, elite_count=1001, iter_max=3, mean_divisor=O.DIV_ELITE_COUNT,
             tol=0.0, replay_normals=z)
    run.step(1, z)
    o.step()
    assert sorted(run.elites().tolist()) == sorted(o.elites().tolist()), "ce ties"
    run.close()
    done.append("ce ties (replay)")
    n, N, k, world = 4, 3001, 40, 3
    x0, s0 = np.full(n, 1.0), np.full(n, 1.5)
    mk = lambda: (E.CrossEntropy.new(E.Objective.SPHERE).with_sample_size(N).with_importance_selection_size(k)
                  .with_iter_max(3).with_std_dev(s0).with_mean_divisor(E.MeanDivisor.ELITE_COUNT).with_seed(2).with_tol(0.0)
                  .with_context(ctx))
    single = mk().start(x0)
    ranks = [mk().with_virtual_rank(r, world).start(x0) for r in range(world)]
    single.step(1)
    for ph in range(E.CeRun.N_PHASES):
        for r in ranks:
            r.phase(ph)
        if ranks[0].exchange_words(ph):
            words = np.stack([r.get_exchange(ph) for r in ranks])
            for r in ranks:
                r.set_exchange(ph, words)
    el = np.sort(np.concatenate([r.elites() for r in ranks]))
    assert el.tolist() == np.sort(single.elites()).tolist(), "ce virtual ranks"
    done.append("ce phased, 3 virtual ranks")
    small = (E.CrossEntropy.new(E.Objective.RASTRIGIN).with_sample_size(100).with_std_dev(np.ones(5)).with_seed(5)
             .with_context(ctx)).solve(np.full(5, 0.4))
    assert np.isfinite(small).all()
    done.append("ce single-CTA solve")

    # MonteCarlo and MISER
    mc = E.MonteCarlo.new(E.Objective.GAUSSIAN).with_sample_count(20000).with_context(ctx).integrate_with_seed(0.0, 1.0, 3)
    st, m, v = O.mc_integrate(O.GAUSSIAN, [0.0], [1.0], 20000, 3)
    assert st == 0 and abs(mc.mean - m) <= 1e-12
    mi = E.MISER.new(E.Objective.SINCOS).with_context(ctx).integrate_with_seed([0.0, 0.0], [1.0, 1.0], 1729)
    st, m, v, _ = O.miser_integrate(O.SINCOS, [0.0, 0.0], [1.0, 1.0], seed=1729)
    assert st == 0 and abs(mi.mean - m) <= 1e-12
    done.append("monte carlo, miser")

    print("sanitizer_smoke OK: " + "; ".join(done))


if __name__ == "__main__":
    main()
