"""GPU: the CUDA ParticleSwarm path (through the C ABI) against the CPU oracle on identical draws."""
import os

import numpy as np
import pytest

import eqsolver_b200 as E
from conftest import assert_bit_equal, from_bits, golden
from oracle import oracle as O

pytestmark = pytest.mark.gpu

OBJ_POINTS = {
    E.Objective.SPHERE: 7, E.Objective.ROSENBROCK: 9, E.Objective.RASTRIGIN: 10, E.Objective.ACKLEY: 12,
    E.Objective.THREE_CIRCLES: 2, E.Objective.HEAVY: 11, E.Objective.USER: 6,
}


def test_objective_kats(ctx):
    for e in golden("objective_kats.json"):
        got = ctx.objective_eval(e["objective"], e["x"])[0]
        want = from_bits(e["bits"])
        if e["objective"] in (2, 3, 7, 8):  # cos/exp/sin: libdevice vs glibc, <= a few ulp of the summands
            assert abs(got - want) <= 1e-12 * max(1.0, abs(want)) + 1e-13, e["note"]
        else:
            assert_bit_equal([got], [want], e["note"])


def test_objectives_random_points_vs_oracle(ctx):
    rng = np.random.default_rng(0)
    for obj, n in OBJ_POINTS.items():
        x = rng.uniform(-3, 3, size=(257, n))
        got = ctx.objective_eval(obj, x)
        want = np.array([O.objective(int(obj), row) for row in x])
        if obj in (E.Objective.RASTRIGIN, E.Objective.ACKLEY):
            np.testing.assert_allclose(got, want, rtol=1e-13, atol=1e-12)
        else:
            assert_bit_equal(got, want, obj.name)


def test_philox_kats_and_contract(ctx):
    g = golden("philox_kats.json")
    for v in g["random123"]:
        assert list(ctx.philox_blocks([v["ctr"]], v["key"])[0]) == v["out"]
    rng = np.random.default_rng(1)
    ctr = rng.integers(0, 2 ** 32, size=(1000, 4), dtype=np.uint64).astype(np.uint32)
    key = np.array([123456789, 987654321], dtype=np.uint32)
    got = ctx.philox_blocks(ctr, key)
    for i in range(0, 1000, 37):
        assert list(got[i]) == O.philox4x32_10(ctr[i], key)


def test_draws_match_oracle_bit_exact(ctx):
    seed, n = 2 ** 40 + 1729, 9
    pos, vel = ctx.draws(0, seed, 0, 2 ** 33, 50, n)
    rp, rg = ctx.draws(1, seed, 17, 123, 50, n)
    for i in (0, 1, 49):
        p, v = O.pso_init_draws(seed, 2 ** 33 + i, n)
        assert_bit_equal(pos[i], p); assert_bit_equal(vel[i], v)
        a, b = O.pso_step_draws(seed, 17, 123 + i, n)
        assert_bit_equal(rp[i], a); assert_bit_equal(rg[i], b)
    (z,) = ctx.draws(2, seed, 3, 1000, 64, n)
    want = np.stack([O.ce_normals(seed, 3, 1000 + i, n) for i in range(64)])
    np.testing.assert_allclose(z, want, rtol=0, atol=4e-15)  # log/sincospi: libdevice vs glibc


def test_box_muller_normals_stay_within_ulps_of_the_oracle(ctx):
    """devmath.cuh's log / sqrt / sincos against glibc through the whole Box-Muller transform, on 64k normals: the root
    is correctly rounded and the log and the sine / cosine are within 1-2 ulp, so a normal is within a few ulp of 1."""
    seed, n, count = 977, 16, 4096
    (z,) = ctx.draws(2, seed, 5, 2 ** 35, count, n)
    want = np.stack([O.ce_normals(seed, 5, 2 ** 35 + i, n) for i in range(count)])
    assert np.isfinite(z).all()
    err = np.abs(z - want)
    assert (err <= 8.0 * np.maximum(np.spacing(np.abs(want)), np.spacing(1.0))).all(), err.max()  # a CPU model of the formulas sees <= 3
    assert (z == want).mean() > 0.5  # most of them agree to the last bit


def make(objective, n, N, mode, seed, lo=-5.0, hi=10.0, tol=0.0, iter_max=50):
    lower, upper = [lo] * n, [hi] * n
    g = (E.ParticleSwarm.new(objective, lower, upper).with_particle_count(N).with_iter_max(iter_max).with_tol(tol)
         .with_seed(seed).with_mode(mode))
    o = O.Pso(int(objective), lower, upper, particle_count=N, iter_max=iter_max, tol=tol, mode=int(mode), seed=seed)
    return g, o


def compare_state(run, orc, exact, what):
    gx, gv, gpb, gpbc = run.state()
    ox, ov, opb, opbc = orc.state()
    g, gc, it, stalled, _ = run.gbest()
    if exact:
        assert_bit_equal(gx, ox, what + " x"); assert_bit_equal(gv, ov, what + " v")
        assert_bit_equal(gpb, opb, what + " pbest"); assert_bit_equal(gpbc, opbc, what + " pbest cost")
        assert_bit_equal(g, orc.gbest(), what + " gbest"); assert_bit_equal([gc], [orc.gbest_cost()])
        assert_bit_equal(run.ring()[0], orc.ring()[0], what + " ring")
    else:  # north_star: positions and velocities within 1e-12 relative
        np.testing.assert_allclose(gx, ox, rtol=1e-12, atol=1e-12, err_msg=what)
        np.testing.assert_allclose(gv, ov, rtol=1e-12, atol=1e-12, err_msg=what)
        np.testing.assert_allclose(gpbc, opbc, rtol=1e-12, atol=1e-12, err_msg=what)
        np.testing.assert_allclose(g, orc.gbest(), rtol=1e-12, atol=1e-12, err_msg=what)
    assert run.ring()[1] == orc.ring()[1], what + " ring index"
    assert it == orc.iters()


@pytest.mark.parametrize("mode", [E.PsoMode.SYNCHRONOUS, E.PsoMode.SEQUENTIAL_EXACT])
def test_golden_trajectories(ctx, mode):
    for g in golden("pso_golden.json"):
        if g["mode"] != int(mode):
            continue
        ps = (E.ParticleSwarm.new(g["objective"], g["lower"], g["upper"]).with_particle_count(g["N"])
              .with_iter_max(g["iters"]).with_seed(g["seed"]).with_mode(mode).with_context(ctx))
        run = ps.start(g["x0"])
        tr = g["trace"]
        gb, gc, _, _, _ = run.gbest()
        assert_bit_equal(gb, [from_bits(b) for b in tr[0]["gbest"]]); assert_bit_equal([gc], [from_bits(tr[0]["gbest_cost"])])
        for t in range(1, len(tr)):
            run.step(1)
            gb, gc, it, _, _ = run.gbest()
            assert it == t
            assert_bit_equal(gb, [from_bits(b) for b in tr[t]["gbest"]], f"gbest after pass {t}")
            assert_bit_equal([gc], [from_bits(tr[t]["gbest_cost"])])
            assert run.ring()[1] == tr[t]["ring_i"]
        x, v, pb, pbc = run.state()
        assert_bit_equal(x, [[from_bits(b) for b in row] for row in g["final_x"]])
        assert_bit_equal(v, [[from_bits(b) for b in row] for row in g["final_v"]])
        assert_bit_equal(pbc, [from_bits(b) for b in g["final_pbest_cost"]])
        assert_bit_equal(run.ring()[0], [from_bits(b) for b in g["ring"]])


@pytest.mark.parametrize("objective,exact", [(E.Objective.SPHERE, True), (E.Objective.ROSENBROCK, True),
                                             (E.Objective.USER, True), (E.Objective.HEAVY, True),
                                             (E.Objective.RASTRIGIN, False), (E.Objective.ACKLEY, False)])
@pytest.mark.parametrize("mode", [E.PsoMode.SYNCHRONOUS, E.PsoMode.SEQUENTIAL_EXACT])
def test_philox_parity_per_iteration(ctx, objective, exact, mode):
    """Both sides generate the SAME Philox draws; state compared after init and after every pass."""
    n, N, iters = (7, 1500, 6) if mode == E.PsoMode.SEQUENTIAL_EXACT else (13, 4099, 10)
    lo, hi = (-5.12, 5.12) if objective in (E.Objective.RASTRIGIN,) else (-5.0, 10.0)
    g, o = make(objective, n, N, mode, seed=1729, lo=lo, hi=hi)
    x0 = np.full(n, 3.0)
    run = g.with_context(ctx).start(x0)
    o.init(x0)
    compare_state(run, o, exact, "init")
    for t in range(iters):
        run.step(1)
        assert o.step() == 0
        compare_state(run, o, exact, f"pass {t}")


@pytest.mark.parametrize("objective", [E.Objective.RASTRIGIN, E.Objective.ACKLEY])
@pytest.mark.parametrize("mode", [E.PsoMode.SYNCHRONOUS, E.PsoMode.SEQUENTIAL_EXACT])
@pytest.mark.parametrize("n,N", [(8, 3000), (13, 700)])
def test_cosine_arguments_out_of_range_in_some_particles(ctx, objective, mode, n, N):
    """The step kernels test the cosine's argument range once per particle (devmath.cuh: DEFER) and evaluate a particle
    that failed it a second time through the scalar path.  One coordinate with bounds of +-3e5 puts |2 pi x| beyond the
    polynomial's 8e5 in about half of the particles (and in gbest's neighbourhood); the others never leave the fast path.
    Both kinds must agree with the oracle's glibc cosine like any other Rastrigin / Ackley run."""
    lower, upper = [-5.0] * n, [5.0] * n
    lower[5], upper[5] = -3.0e5, 3.0e5  # in a full 4-coordinate group for n = 8 and n = 13
    if n == 13:
        lower[12], upper[12] = -2.0e5, 4.0e5  # and in the ragged last group
    g = (E.ParticleSwarm.new(objective, lower, upper).with_particle_count(N).with_iter_max(50).with_tol(0.0)
         .with_seed(99).with_mode(mode))
    o = O.Pso(int(objective), lower, upper, particle_count=N, iter_max=50, tol=0.0, mode=int(mode), seed=99)
    x0 = np.full(n, 1.0)
    run = g.with_context(ctx).start(x0)
    o.init(x0)
    compare_state(run, o, False, "init")
    x = run.state()[0]
    far = (np.abs(2 * np.pi * x) >= 8.0e5).any(axis=1)
    assert 0.2 < far.mean() < 0.9  # the population really is a mix of both paths
    for t in range(5):
        run.step(1)
        assert o.step() == 0
        compare_state(run, o, False, f"pass {t}")


@pytest.mark.parametrize("mode", [E.PsoMode.SYNCHRONOUS, E.PsoMode.SEQUENTIAL_EXACT])
def test_replay_mode_host_draws(ctx, mode):
    """Replay: identical HOST-generated draws (numpy Philox4x64, unrelated to the device generator) fed to both."""
    n, N, iters = 6, 777, 5
    rng = np.random.Generator(np.random.Philox(1729))
    r_init = rng.random((N, 2 * n))
    r_step = rng.random((iters, N, n, 2))
    lower, upper = [-5.0] * n, [10.0] * n
    g = (E.ParticleSwarm.new(E.Objective.ROSENBROCK, lower, upper).with_particle_count(N).with_iter_max(iters)
         .with_tol(0.0).with_mode(mode).with_replay_init(r_init).with_context(ctx))
    o = O.Pso(O.ROSENBROCK, lower, upper, particle_count=N, iter_max=iters, tol=0.0, mode=int(mode),
              replay_init=r_init, replay_step=r_step)
    x0 = np.zeros(n)
    run = g.start(x0)
    o.init(x0)
    compare_state(run, o, True, "init")
    # two passes in one call, then one at a time
    run.step(2, r_step[:2]); o.step(); o.step()
    compare_state(run, o, True, "passes 0-1")
    for t in range(2, iters):
        run.step(1, r_step[t:t + 1]); o.step()
        compare_state(run, o, True, f"pass {t}")


def test_ragged_and_tiny_populations(ctx):
    for N in (1, 2, 3, 31, 255, 256, 257, 513):
        g, o = make(E.Objective.ROSENBROCK, 3, N, E.PsoMode.SYNCHRONOUS, seed=N)
        x0 = np.array([1.5, -0.5, 2.0])
        run = g.with_context(ctx).start(x0); o.init(x0)
        run.step(3); [o.step() for _ in range(3)]
        compare_state(run, o, True, f"N={N}")
    # n = 1 (Rosenbrock degenerates to 0), n not a multiple of the unroll
    for n in (1, 2, 5, 33):
        g, o = make(E.Objective.SPHERE, n, 300, E.PsoMode.SYNCHRONOUS, seed=n)
        x0 = np.full(n, 2.0)
        run = g.with_context(ctx).start(x0); o.init(x0)
        run.step(4); [o.step() for _ in range(4)]
        compare_state(run, o, True, f"n={n}")


def test_empty_population_and_zero_iterations(ctx):
    ps = E.ParticleSwarm.new(E.Objective.SPHERE, [-1.0] * 3, [1.0] * 3).with_particle_count(0).with_context(ctx)
    x0 = np.array([0.25, 0.5, -0.75])
    assert_bit_equal(ps.solve(x0), x0)  # no particle ever beats f(x0): gbest stays x0 (particle_swarm.rs:357)
    ps = E.ParticleSwarm.new(E.Objective.SPHERE, [-1.0] * 3, [1.0] * 3).with_iter_max(0).with_seed(5).with_context(ctx)
    o = O.Pso(O.SPHERE, [-1.0] * 3, [1.0] * 3, iter_max=0, seed=5)
    assert_bit_equal(ps.solve(x0), o.solve(x0))


def test_stall_exit_matches_oracle(ctx):
    """The stall test (particle_swarm.rs:395-397,433-441) fires on the same pass on both sides."""
    n, N = 2, 64
    for mode in (E.PsoMode.SEQUENTIAL_EXACT, E.PsoMode.SYNCHRONOUS):
        g, o = make(E.Objective.SPHERE, n, N, mode, seed=9, lo=-1.0, hi=1.0, tol=1e-3, iter_max=400)
        x0 = np.array([0.9, -0.9])
        got = g.with_context(ctx).solve(x0)
        want = o.solve(x0)
        assert_bit_equal(got, want)
        assert g.last_report.iters_run == o.iters()
        assert o.iters() < 400, "the case must actually stall"
        assert g.last_report.stalled == 1
        assert g.last_report.evals == 1 + N * (1 + o.iters())


def test_solve_known_optima(ctx):
    # reference style: loose tolerance on the known answer (tests/integrators.rs:87-116)
    x = (E.ParticleSwarm.new(E.Objective.RASTRIGIN, [-5.12] * 2, [5.12] * 2).with_particle_count(1000)
         .with_iter_max(1000).with_tol(0.0).with_mode(E.PsoMode.SYNCHRONOUS).with_seed(1).with_context(ctx)
         .solve([5.0, 5.0]))
    assert np.abs(x).max() < 1e-6
    x = (E.ParticleSwarm.new(E.Objective.THREE_CIRCLES, [1.0, 1.0], [7.0, 7.0]).with_particle_count(500)
         .with_iter_max(300).with_tol(0.0).with_seed(2).with_context(ctx).solve([4.5, 2.5]))
    assert np.allclose(x, [4.217265312839526, 2.317879970005811], atol=1e-6)  # tests/multi.rs:70
    x = (E.ParticleSwarm.new(E.Objective.ROSENBROCK, [-5.0] * 4, [10.0] * 4).with_particle_count(20000)
         .with_iter_max(600).with_tol(0.0).with_mode(E.PsoMode.SYNCHRONOUS).with_seed(3).with_context(ctx)
         .solve([0.0] * 4))
    assert np.allclose(x, 1.0, atol=1e-4)


def test_virtual_ranks_match_single_rank(ctx):
    """Multi-rank protocol on one GPU: 3 virtual ranks, records moved by the harness, equal one rank bit for bit."""
    n, N, iters, world = 5, 1001, 6, 3
    lower, upper, x0 = [-5.0] * n, [10.0] * n, np.full(n, 3.0)
    base = lambda: (E.ParticleSwarm.new(E.Objective.ROSENBROCK, lower, upper).with_particle_count(N).with_tol(0.0)
                    .with_seed(77).with_mode(E.PsoMode.SYNCHRONOUS).with_context(ctx))
    single = base().start(x0)
    ranks = [base().with_virtual_rank(r, world).start() for r in range(world)]
    assert [r.first for r in ranks] == [0, 334, 668] and sum(r.local for r in ranks) == N

    def gather():
        recs = np.stack([r.get_record() for r in ranks])
        for r in ranks:
            r.set_records(recs)

    for r in ranks:
        r.init_local(x0, 0)
    gather()
    for r in ranks:
        r.init_local(x0, 1)
    gather()
    for r in ranks:
        r.init_apply()
    for t in range(iters + 1):
        sg, sc, sit, _, _ = single.gbest()
        sring = single.ring()
        for r in ranks:
            g, c, it, _, _ = r.gbest()
            assert_bit_equal(g, sg, f"gbest t={t}"); assert_bit_equal([c], [sc]); assert it == sit
            assert_bit_equal(r.ring()[0], sring[0]); assert r.ring()[1] == sring[1]
        if t == iters:
            break
        single.step(1)
        for r in ranks:
            r.step_local()
        gather()
        for r in ranks:
            r.step_apply()
    x = np.concatenate([r.state()[0] for r in ranks])
    assert_bit_equal(x, single.state()[0])


def test_full_size_c2_properties(ctx):
    """BASELINE config 2 at full size (Rosenbrock n=32, 2^20 particles): size-independent properties."""
    n, N = 32, 1 << 20
    ps = (E.ParticleSwarm.new(E.Objective.ROSENBROCK, [-5.0] * n, [10.0] * n).with_particle_count(N).with_tol(0.0)
          .with_seed(1729).with_mode(E.PsoMode.SYNCHRONOUS).with_context(ctx))
    run = ps.start(np.full(n, 8.0))  # f(x0) ~ 9.7e6: the swarm beats it at once
    _, c_prev, _, _, _ = run.gbest()
    pbc_prev = run.state()[3]
    for _ in range(3):
        run.step(1)
        x, v, pb, pbc = run.state()
        g, c, it, _, _ = run.gbest()
        assert c <= c_prev and np.all(pbc <= pbc_prev)            # gbest / pbest costs never increase
        assert c == pbc.min()                                      # gbest cost = min pbest cost
        i = int(np.argmin(pbc))
        assert_bit_equal(pb[i], g)                                 # gbest position = that particle's pbest
        idx = np.random.default_rng(it).integers(0, N, 64)
        for j in idx:                                              # pbest cost is f(pbest) -- checksum of the state
            assert pbc[j] == O.objective(O.ROSENBROCK, pb[j])
        c_prev, pbc_prev = c, pbc
    # and a sampled replay against the oracle's arithmetic for pass 3: recompute 32 particles from the state
    x0s, v0s, pb0s, _ = x, v, pb, pbc
    g_before = g.copy()
    run.step(1)
    x1, v1, _, _ = run.state()
    for j in np.random.default_rng(5).integers(0, N, 32):
        rp, rg = O.pso_step_draws(1729, 3, int(j), n)
        vv = 0.5 * v0s[j] + 1.0 * rp * (pb0s[j] - x0s[j]) + 1.0 * rg * (g_before - x0s[j])
        assert_bit_equal(v1[j], vv); assert_bit_equal(x1[j], x0s[j] + vv)


def test_virtual_ranks_stall_exit_matches_oracle(ctx):
    """Multi-rank stall: once the stall test fires, further passes (step + exchange + apply) change nothing."""
    n, N, world = 2, 128, 2
    lower, upper, x0 = [-1.0] * n, [1.0] * n, np.array([0.9, -0.9])
    base = lambda: (E.ParticleSwarm.new(E.Objective.SPHERE, lower, upper).with_particle_count(N).with_tol(1e-3)
                    .with_iter_max(400).with_seed(9).with_mode(E.PsoMode.SYNCHRONOUS).with_context(ctx))
    o = O.Pso(O.SPHERE, lower, upper, particle_count=N, tol=1e-3, iter_max=400, mode=O.SYNCHRONOUS, seed=9)
    want = o.solve(x0)
    assert o.iters() < 400, "the case must stall"
    ranks = [base().with_virtual_rank(r, world).start() for r in range(world)]

    def gather():
        recs = np.stack([r.get_record() for r in ranks])
        for r in ranks:
            r.set_records(recs)

    for ph in (0, 1):
        for r in ranks:
            r.init_local(x0, ph)
        gather()
    for r in ranks:
        r.init_apply()
    for _ in range(o.iters() + 25):  # 25 passes beyond the stall
        for r in ranks:
            r.step_local()
        gather()
        for r in ranks:
            r.step_apply()
    for r in ranks:
        g, c, it, stalled, _ = r.gbest()
        assert it == o.iters() and stalled
        assert_bit_equal(g, want)


@pytest.mark.parametrize("mode", [E.PsoMode.SYNCHRONOUS, E.PsoMode.SEQUENTIAL_EXACT])
def test_small_kernel_equals_multi_launch_path(ctx, mode, monkeypatch):
    """K8 (one CTA, all passes in one launch) and the multi-launch path produce identical states."""
    n, N, iters = 6, 700, 9
    lower, upper, x0 = [-5.0] * n, [10.0] * n, np.full(n, 3.0)
    mk = lambda: (E.ParticleSwarm.new(E.Objective.ROSENBROCK, lower, upper).with_particle_count(N).with_tol(0.0)
                  .with_seed(21).with_mode(mode).with_context(ctx))
    small = mk().start(x0)
    small.step(iters)
    monkeypatch.setenv("EQS_PSO_NO_SMALL", "1")
    big = mk().start(x0)
    big.step(iters)
    monkeypatch.delenv("EQS_PSO_NO_SMALL")
    for a, b in zip(small.state(), big.state()):
        assert_bit_equal(a, b)
    assert_bit_equal(small.gbest()[0], big.gbest()[0]); assert small.gbest()[1:4] == big.gbest()[1:4]
    assert_bit_equal(small.ring()[0], big.ring()[0])
    assert small.last_step_ms()[1] == 1 and big.last_step_ms()[1] >= iters   # launches


def test_sequential_passes_are_one_persistent_kernel(ctx):
    """Above the single-CTA limit, on one rank, ALL rounds of ALL passes of a step() call run inside one cooperative
    launch (pso_seqp_kernel: grid barriers instead of kernel boundaries) behind a one-thread arming kernel."""
    n, N = 8, 50000
    g, o = make(E.Objective.ROSENBROCK, n, N, E.PsoMode.SEQUENTIAL_EXACT, seed=5)
    x0 = np.full(n, 2.0)
    run = g.with_context(ctx).start(x0)
    o.init(x0)
    for k in (7, 1, 12):
        run.step(k)
        for _ in range(k):
            assert o.step() == 0
        assert run.last_step_ms()[1] == 2
        compare_state(run, o, True, f"after {k} passes")
    assert run.gbest()[4] > 20        # many repair rounds happened


def test_sequential_host_driven_path_large_population(ctx):
    """Populations above the single-CTA limit run speculate-and-repair with host-driven rounds."""
    n, N, iters = 5, 2600, 4
    g, o = make(E.Objective.ROSENBROCK, n, N, E.PsoMode.SEQUENTIAL_EXACT, seed=33)
    x0 = np.full(n, 3.0)
    run = g.with_context(ctx).start(x0); o.init(x0)
    for t in range(iters):
        run.step(1); o.step()
        compare_state(run, o, True, f"pass {t}")


def test_degenerate_inputs_match_oracle(ctx):
    # NaN start: f(x0) = NaN, no `cost < NaN` is ever true (:376, :420) -> gbest stays x0, the ring is never pushed
    n, N = 3, 500
    for mode in (E.PsoMode.SYNCHRONOUS, E.PsoMode.SEQUENTIAL_EXACT):
        g, o = make(E.Objective.SPHERE, n, N, mode, seed=4)
        x0 = np.array([np.nan, 1.0, 2.0])
        run = g.with_context(ctx).start(x0); o.init(x0)
        run.step(3); [o.step() for _ in range(3)]
        gb, gc, it, _, _ = run.gbest()
        assert np.isnan(gc) and np.isnan(o.gbest_cost()) and it == o.iters() == 3
        assert np.array_equal(gb.view(np.uint64), o.gbest().view(np.uint64))
        assert_bit_equal(run.state()[0], o.state()[0]); assert_bit_equal(run.ring()[0], o.ring()[0])
    # lower == upper is legal for Uniform::new_inclusive: every particle starts on the same point with zero velocity
    lower = upper = [1.5, -2.0]
    ps = (E.ParticleSwarm.new(E.Objective.ROSENBROCK, lower, upper).with_particle_count(200).with_tol(0.0).with_seed(1)
          .with_mode(E.PsoMode.SYNCHRONOUS).with_context(ctx))
    o = O.Pso(O.ROSENBROCK, lower, upper, particle_count=200, tol=0.0, mode=O.SYNCHRONOUS, seed=1)
    x0 = np.array([0.0, 0.0])
    run = ps.start(x0); o.init(x0)
    run.step(4); [o.step() for _ in range(4)]
    compare_state(run, o, True, "degenerate bounds")
    # zero coefficients: particles keep half their velocity each pass
    ps = (E.ParticleSwarm.new(E.Objective.SPHERE, [-1.0] * 2, [1.0] * 2).with_particle_count(100).with_tol(0.0)
          .with_cognitive_coefficient(0.0).with_social_coefficient(0.0).with_inertia_weight(0.25).with_seed(2)
          .with_mode(E.PsoMode.SEQUENTIAL_EXACT).with_context(ctx))
    o = O.Pso(O.SPHERE, [-1.0] * 2, [1.0] * 2, particle_count=100, tol=0.0, c1=0.0, c2=0.0, w=0.25, seed=2)
    run = ps.start(x0); o.init(x0)
    run.step(3); [o.step() for _ in range(3)]
    compare_state(run, o, True, "zero coefficients")


def test_large_dimension_and_limits(ctx):
    for n, N, mode in ((1000, 300, E.PsoMode.SYNCHRONOUS), (257, 3000, E.PsoMode.SYNCHRONOUS),
                       (130, 260, E.PsoMode.SEQUENTIAL_EXACT)):
        g, o = make(E.Objective.ROSENBROCK, n, N, mode, seed=n, lo=-2.0, hi=2.0)
        x0 = np.full(n, 0.5)
        run = g.with_context(ctx).start(x0); o.init(x0)
        run.step(2); [o.step() for _ in range(2)]
        compare_state(run, o, True, f"n={n}")
    # the gbest tile lives in shared memory: n is capped
    with pytest.raises(E.IncorrectInput):
        E.ParticleSwarm.new(E.Objective.SPHERE, [-1.0] * 13000, [1.0] * 13000).with_context(ctx).solve([0.0] * 13000)
    ps = (E.ParticleSwarm.new(E.Objective.SPHERE, [-1.0] * 12800, [1.0] * 12800).with_particle_count(64).with_iter_max(2)
          .with_mode(E.PsoMode.SYNCHRONOUS).with_context(ctx))
    assert ps.solve([0.5] * 12800).shape == (12800,)


def test_device_cos_accuracy(ctx):
    """eqs_cos (objectives.cuh) against glibc on the arguments the path produces: absolute error <= 2.8e-16 (1.25 ulp of
    1.0 -- one even polynomial on [-pi/2, pi/2]; every consumer adds the cosine to sums of magnitude >= 1, so an absolute
    bound is the one that matters), exact special cases."""
    rng = np.random.default_rng(7)
    xs = np.concatenate([rng.uniform(-700, 700, 400000), rng.uniform(-8e5, 8e5, 200000), rng.uniform(-1e-3, 1e-3, 50000),
                         2 * np.pi * rng.uniform(-5.12, 5.12, 200000), np.array([0.0, -0.0, np.pi / 2, np.pi, 1e6, -3e7, 1e15, 1e300])])
    got = ctx.device_cos(xs)
    want = np.cos(xs)
    ulp = np.spacing(np.abs(want))
    err = np.abs(got - want) / np.maximum(ulp, np.finfo(float).tiny)
    ok = (err <= 2.0) | (np.abs(got - want) <= 2.8e-16)
    assert ok.all(), (xs[~ok][:5], got[~ok][:5], want[~ok][:5], err.max())
    sp = ctx.device_cos(np.array([np.inf, -np.inf, np.nan]))
    assert np.isnan(sp).all()
    assert ctx.device_cos(np.array([0.0]))[0] == 1.0


def test_seeded_and_replay_solves_are_public_api(ctx):
    """solve_with_seed / solve_with_draws (the optimisers' twin of integrate_with_rng)."""
    n, N, iters = 5, 777, 9
    lower, upper, x0 = [-5.0] * n, [10.0] * n, np.full(n, 3.0)
    ps = E.ParticleSwarm.new(E.Objective.ROSENBROCK, lower, upper).with_particle_count(N).with_iter_max(iters) \
        .with_tol(0.0).with_context(ctx)
    a = ps.solve_with_seed(x0, 42)
    assert np.array_equal(a, ps.solve_with_seed(x0, 42)) and not np.array_equal(a, ps.solve_with_seed(x0, 43))
    o = O.Pso(O.ROSENBROCK, lower, upper, particle_count=N, iter_max=iters, tol=0.0, mode=O.SEQUENTIAL, seed=42)
    assert np.array_equal(a.view(np.uint64), o.solve(x0).view(np.uint64))
    rng = np.random.Generator(np.random.Philox(7))
    init, steps = rng.random((N, 2, n)), rng.random((iters, N, n, 2))
    got = ps.solve_with_draws(x0, init, steps)
    od = O.Pso(O.ROSENBROCK, lower, upper, particle_count=N, iter_max=iters, tol=0.0, mode=O.SEQUENTIAL,
               replay_init=init, replay_step=steps)
    assert np.array_equal(got.view(np.uint64), od.solve(x0).view(np.uint64))
    with pytest.raises(E.IncorrectInput):
        ps.solve_with_draws(x0, init, steps.reshape(-1)[:-1])


@pytest.mark.parametrize("window", [None, "64", "1000", "100000"])
@pytest.mark.parametrize("objective,exact", [(E.Objective.ROSENBROCK, True), (E.Objective.HEAVY, True),
                                             (E.Objective.RASTRIGIN, False)])
def test_sequential_windowed_rounds_match_the_oracle(ctx, monkeypatch, objective, exact, window):
    """Sequential mode beyond the single-CTA kernel: device-driven windowed speculate-and-repair (pso_seqw_kernel).
    Any window size must give the reference's sequential semantics: whole state after every pass, several passes
    per step() call, buffers swapping roles between passes."""
    if window is None:
        monkeypatch.delenv("EQS_PSO_SEQ_WINDOW", raising=False)
    else:
        monkeypatch.setenv("EQS_PSO_SEQ_WINDOW", window)
    n, N = (6, 2500) if objective == E.Objective.HEAVY else (9, 5003)
    lo, hi = (-5.12, 5.12) if objective == E.Objective.RASTRIGIN else (-5.0, 10.0)
    g, o = make(objective, n, N, E.PsoMode.SEQUENTIAL_EXACT, seed=77, lo=lo, hi=hi)
    x0 = np.full(n, 3.0)
    run = g.with_context(ctx).start(x0)
    o.init(x0)
    compare_state(run, o, exact, "init")
    for t, k in enumerate([1, 1, 2, 3, 1]):
        run.step(k)
        for _ in range(k):
            assert o.step() == 0
        compare_state(run, o, exact, f"after step call {t}")
        assert run.gbest()[4] > 0     # gbest did improve: the repair rounds were exercised


def test_sequential_windowed_replay_stall_and_solve(ctx, monkeypatch):
    monkeypatch.setenv("EQS_PSO_SEQ_WINDOW", "777")
    n, N, iters = 4, 3001, 5
    lower, upper, x0 = [-5.0] * n, [10.0] * n, np.full(n, 1.5)
    rng = np.random.Generator(np.random.Philox(5))
    init, steps = rng.random((N, 2, n)), rng.random((iters, N, n, 2))
    ps = (E.ParticleSwarm.new(E.Objective.ROSENBROCK, lower, upper).with_particle_count(N).with_iter_max(iters)
          .with_tol(0.0).with_context(ctx))
    od = O.Pso(O.ROSENBROCK, lower, upper, particle_count=N, iter_max=iters, tol=0.0, mode=O.SEQUENTIAL,
               replay_init=init, replay_step=steps)
    assert_bit_equal(ps.solve_with_draws(x0, init, steps), od.solve(x0))
    # whole solve with the stall test active: same pass count, same answer
    monkeypatch.delenv("EQS_PSO_SEQ_WINDOW")
    ps = (E.ParticleSwarm.new(E.Objective.SPHERE, [-1.0] * 3, [1.0] * 3).with_particle_count(4000).with_iter_max(400)
          .with_tol(1e-3).with_seed(9).with_context(ctx))
    o = O.Pso(O.SPHERE, [-1.0] * 3, [1.0] * 3, particle_count=4000, iter_max=400, tol=1e-3, mode=O.SEQUENTIAL, seed=9)
    got = ps.solve(np.full(3, 0.5))
    assert_bit_equal(got, o.solve(np.full(3, 0.5)))
    assert ps.last_report.iters_run == o.iters() and 0 < o.iters() < 400 and ps.last_report.stalled == 1
    # the host-driven rounds (what ranks > 1 use) stay available and agree
    monkeypatch.setenv("EQS_PSO_SEQ_HOST", "1")
    ps2 = (E.ParticleSwarm.new(E.Objective.SPHERE, [-1.0] * 3, [1.0] * 3).with_particle_count(4000).with_iter_max(400)
           .with_tol(1e-3).with_seed(9).with_context(ctx))
    assert_bit_equal(ps2.solve(np.full(3, 0.5)), got)


def test_concurrent_solves_from_threads(ctx):
    """the reference's solve(&self) is re-entrant, so the C ABI must be callable concurrently on distinct
    contexts, and on distinct solver handles of one context (its stream serialises the launches)."""
    import threading
    jobs = [(E.Objective.ROSENBROCK, 6, 3000 + 17 * j, 11 + j) for j in range(6)]

    def solve(c, job):
        obj, n, N, seed = job
        return (E.ParticleSwarm.new(obj, [-5.0] * n, [10.0] * n).with_particle_count(N).with_iter_max(12).with_tol(0.0)
                .with_seed(seed).with_mode(E.PsoMode.SYNCHRONOUS if seed % 2 else E.PsoMode.SEQUENTIAL_EXACT)
                .with_context(c).solve(np.full(n, 2.0)))

    want = [solve(ctx, j) for j in jobs]
    for shared in (False, True):
        got, errs = [None] * len(jobs), []

        def work(i):
            try:
                got[i] = solve(ctx if shared else E.Context(0), jobs[i])
            except Exception as e:  # noqa: BLE001
                errs.append(e)
        th = [threading.Thread(target=work, args=(i,)) for i in range(len(jobs))]
        [t.start() for t in th]
        [t.join() for t in th]
        assert not errs, errs
        for a, b in zip(got, want):
            assert_bit_equal(a, b, f"shared context = {shared}")
    # MonteCarlo and MISER from threads as well
    res = [None] * 4

    def integ(i):
        c = E.Context(0)
        res[i] = (E.MonteCarlo.new(E.Objective.GAUSSIAN).with_sample_count(100000).with_context(c)
                  .integrate_with_seed(0.0, 1.0, 5),
                  E.MISER.new(E.Objective.SINCOS).with_context(c).integrate_with_seed([0.0, 0.0], [1.0, 1.0], 5))
    th = [threading.Thread(target=integ, args=(i,)) for i in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert all(r == res[0] and r is not None for r in res)


def test_full_size_c4_properties(ctx):
    """BASELINE config 4 at its per-GPU size (Ackley n=256, 2^21 particles = 12.9 GB of state): size-independent
    properties of the synchronous pass."""
    n, N, lo, hi, seed = 256, 1 << 21, -32.768, 32.768, 1729
    ps = (E.ParticleSwarm.new(E.Objective.ACKLEY, [lo] * n, [hi] * n).with_particle_count(N).with_tol(0.0)
          .with_seed(seed).with_mode(E.PsoMode.SYNCHRONOUS).with_context(ctx))
    run = ps.start(np.full(n, 30.0))
    g, c_prev, _, _, _ = run.gbest()
    pbc_prev = run.pbest_costs()
    f0 = O.objective(O.ACKLEY, np.full(n, 30.0))
    assert c_prev == min(pbc_prev.min(), f0)
    # A sample of particles is carried through the same passes in the ORACLE's arithmetic, from nothing but the draw
    # contract and the gbest positions the device reports (the 13 GB of state stay on the device): init :361-389, update
    # :406-413 in rustc's evaluation order, objective, pbest test :417-419.  Checked against the device's pbest costs.
    sample = np.random.default_rng(4).integers(0, N, 24)
    mine = {}
    for j in sample:
        pu, vu = O.pso_init_draws(seed, int(j), n)
        x = pu * (hi - lo) + lo
        dist = abs(hi - lo)
        v = vu * (dist + dist) + (-dist)
        mine[int(j)] = [x, v, x.copy(), O.objective(O.ACKLEY, x)]
    for j, (x, v, pb, pc) in mine.items():
        assert abs(pbc_prev[j] - pc) <= 1e-12 * abs(pc), f"init cost of particle {j}"
    for t in range(3):
        g_before = g.copy()
        run.step(1)
        g, c, it, stalled, _ = run.gbest()
        pbc = run.pbest_costs()
        assert it == t + 1 and not stalled
        assert c <= c_prev and np.all(pbc <= pbc_prev)              # gbest / pbest costs never increase
        assert c <= pbc.min() and (c == pbc.min() or c == f0)       # gbest cost = the best cost seen
        assert abs(O.objective(O.ACKLEY, g) - c) <= 1e-12 * c       # the gbest cost is f(gbest position)
        assert np.all(np.isfinite(pbc))
        for j, st in mine.items():
            x, v, pb, pc = st
            rp, rg = O.pso_step_draws(seed, t, j, n)
            v = 0.5 * v + 1.0 * rp * (pb - x) + 1.0 * rg * (g_before - x)
            x = x + v
            cost = O.objective(O.ACKLEY, x)
            if cost < pc:
                pb, pc = x.copy(), cost
            st[:] = [x, v, pb, pc]
            assert abs(pbc[j] - pc) <= 1e-12 * abs(pc), f"pbest cost of particle {j} after pass {t}"
        c_prev, pbc_prev = c, pbc
    run.close()


def test_user_objective_hook(ctx, tmp_path):
    """a user objective is a CUDA header named at build time (EQS_USER_OBJECTIVE_HEADER).  Build a
    library variant whose api.cu carries a custom streaming functor and evaluate it through eqs_objective_eval."""
    import ctypes as C
    import subprocess
    from eqsolver_b200 import _ffi, build as kbuild
    hdr = tmp_path / "mine.cuh"
    hdr.write_text('''#pragma once
namespace eqs {
struct UserObjective {  // sum (d + 1) * |x_d|^3, streaming form
    static constexpr bool streaming = true;
    __device__ static void begin(ObjAcc &s, int) { s.a = 0.0; }
    __device__ static void feed(ObjAcc &s, double x, int d) { s.a += (double)(d + 1) * (fabs(x) * x * x); }
    __device__ static double end(const ObjAcc &s, int) { return s.a; }
};
}
''')
    nvcc = kbuild._nvcc()
    obj = tmp_path / "api.o"
    subprocess.check_call([nvcc, *kbuild.NVCC_FLAGS, f'-DEQS_USER_OBJECTIVE_HEADER="{hdr}"', "-c",
                           os.path.join(kbuild.CSRC, "api.cu"), "-o", str(obj)])
    others = [os.path.join(kbuild.OBJ_DIR, s.replace(".cu", ".o")) for s in kbuild.SOURCES if s != "api.cu"]
    for s, o in zip([s for s in kbuild.SOURCES if s != "api.cu"], others):
        if not os.path.exists(o):   # a tree without build/: compile the stock object
            subprocess.check_call([nvcc, *kbuild.NVCC_FLAGS, "-c", os.path.join(kbuild.CSRC, s), "-o", o])
    lib = tmp_path / "libcustom.so"
    subprocess.check_call([nvcc, "-shared", "-o", str(lib), str(obj), *others, "-gencode",
                           "arch=compute_100a,code=sm_100a", "-ldl"])
    L = C.CDLL(str(lib))
    L.eqs_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    L.eqs_objective_eval.argtypes = _ffi.SYMBOLS["eqs_objective_eval"][1]
    h = C.c_void_p()
    assert L.eqs_ctx_create(0, C.byref(h)) == 0
    x = np.random.default_rng(3).uniform(-2, 2, size=(33, 5))
    out = np.empty(33)
    assert L.eqs_objective_eval(h, int(E.Objective.USER), x.ctypes.data_as(_ffi.dp), 5, 33, out.ctypes.data_as(_ffi.dp)) == 0
    want = np.zeros(33)
    for d in range(5):
        want = want + (d + 1.0) * (np.abs(x[:, d]) * x[:, d] * x[:, d])
    assert_bit_equal(out, want)
    L.eqs_ctx_destroy.argtypes = [C.c_void_p]
    L.eqs_ctx_destroy(h)


def test_whole_c4_population_on_one_gpu(ctx):
    """Sized for 180 GB: BASELINE config 4's WHOLE population (2^24 particles x 256 dimensions, 103 GB of state) on one
    GPU -- 64-bit indexing everywhere, same invariants as the per-GPU shard."""
    import torch
    free, _ = torch.cuda.mem_get_info(0)
    if free < 125 * 2 ** 30:
        pytest.skip("needs ~105 GB of free device memory")
    n, N = 256, 1 << 24
    x0 = np.full(n, 30.0)
    ps = (E.ParticleSwarm.new(E.Objective.ACKLEY, [-32.768] * n, [32.768] * n).with_particle_count(N).with_tol(0.0)
          .with_seed(1729).with_mode(E.PsoMode.SYNCHRONOUS).with_context(ctx))
    run = ps.start(x0)
    f0 = O.objective(O.ACKLEY, x0)
    pbc_prev = run.pbest_costs()
    assert pbc_prev.size == N and np.isfinite(pbc_prev).all()
    # spot-check the far end of the index space against the oracle's arithmetic (draws keyed by the 64-bit index)
    for i in (0, N // 2 + 12345, N - 1):
        pos_u, _ = O.pso_init_draws(1729, i, n)
        assert abs(pbc_prev[i] - O.objective(O.ACKLEY, pos_u * 65.536 + -32.768)) <= 1e-12 * pbc_prev[i]
    run.step(3)
    g, c, it, stalled, _ = run.gbest()
    pbc = run.pbest_costs()
    assert it == 3 and not stalled and np.all(pbc <= pbc_prev)
    assert c <= pbc.min() and (c == pbc.min() or c == f0)
    assert abs(O.objective(O.ACKLEY, g) - c) <= 1e-12 * c
    run.close()


def test_out_of_device_memory_is_an_error_not_a_crash(ctx):
    """A population that cannot fit (6.6 TB of state) comes back as ExternalError with the CUDA message, and the
    context stays usable."""
    n = 32
    big = (E.ParticleSwarm.new(E.Objective.SPHERE, [-1.0] * n, [1.0] * n).with_particle_count(1 << 33)
           .with_mode(E.PsoMode.SYNCHRONOUS).with_context(ctx))
    with pytest.raises(E.ExternalError, match="(?i)memory"):
        big.solve(np.zeros(n))
    with pytest.raises(E.ExternalError, match="(?i)memory"):
        (E.CrossEntropy.new(E.Objective.SPHERE).with_sample_size(1 << 38).with_importance_selection_size(1 << 36)
         .with_context(ctx).solve(np.zeros(n)))
    ok = (E.ParticleSwarm.new(E.Objective.SPHERE, [-1.0] * n, [1.0] * n).with_particle_count(500).with_seed(1)
          .with_context(ctx).solve(np.zeros(n)))
    assert ok.shape == (n,) and np.isfinite(ok).all()


def test_stepping_before_init_is_refused(ctx):
    run = (E.ParticleSwarm.new(E.Objective.SPHERE, [-1.0] * 3, [1.0] * 3).with_particle_count(5000)
           .with_context(ctx).start())          # created, not initialised
    for call in (lambda: run.step(1), run.gbest, run.state):
        with pytest.raises(E.IncorrectInput, match="initialised"):
            call()
    run.init(np.zeros(3))
    run.step(2)
    assert run.gbest()[2] == 2
    run.close()


def test_null_data_pointers_with_live_handles(ctx):
    """Live context / solver handles but NULL data pointers: a status, never a fault."""
    import ctypes as C
    from eqsolver_b200 import _ffi
    L = _ffi.lib()
    ok_or_bad = {_ffi.OK, _ffi.ERR_INCORRECT_INPUT}
    ps = E.ParticleSwarm.new(E.Objective.SPHERE, [-1.0] * 3, [1.0] * 3).with_particle_count(3000).with_context(ctx)
    run = ps.start(np.zeros(3))
    h = run._h
    assert L.eqs_pso_init(h, None) == _ffi.ERR_INCORRECT_INPUT
    assert L.eqs_pso_read_gbest(h, None, None, None, None, None) in ok_or_bad
    assert L.eqs_pso_read_state(h, None, None, None, None) in ok_or_bad
    assert L.eqs_pso_read_ring(h, None, None) in ok_or_bad
    assert L.eqs_pso_last_step_ms(h, None, None) in ok_or_bad | {_ffi.ERR_EXTERNAL}   # no step timed yet
    assert L.eqs_pso_local_range(h, None, None) in ok_or_bad
    assert L.eqs_pso_get_record(h, None) == _ffi.ERR_INCORRECT_INPUT
    assert L.eqs_pso_set_records(h, None) == _ffi.ERR_INCORRECT_INPUT
    assert L.eqs_pso_step(h, -5, None) in ok_or_bad and L.eqs_pso_step(h, 0, None) == _ffi.OK
    run.step(1)
    assert run.gbest()[2] == 1
    run.close()
    ce = (E.CrossEntropy.new(E.Objective.SPHERE).with_sample_size(2000).with_importance_selection_size(20)
          .with_context(ctx).start(np.zeros(3)))
    hc = ce._h
    assert L.eqs_ce_read_state(hc, None, None, None, None) in ok_or_bad
    assert L.eqs_ce_read_elites(hc, None, None) in ok_or_bad
    assert L.eqs_ce_read_costs(hc, None) == _ffi.ERR_INCORRECT_INPUT
    assert L.eqs_ce_get_exchange(hc, 1, None) == _ffi.ERR_INCORRECT_INPUT
    assert L.eqs_ce_set_exchange(hc, 1, None) == _ffi.ERR_INCORRECT_INPUT
    assert L.eqs_ce_phase(hc, 99, None) == _ffi.ERR_INCORRECT_INPUT and L.eqs_ce_phase(hc, -1, None) == _ffi.ERR_INCORRECT_INPUT
    assert L.eqs_ce_step(hc, -3, None) in ok_or_bad
    ce.step(1)
    ce.close()
    c = ctx._h
    assert L.eqs_objective_eval(c, 0, None, 3, 5, None) == _ffi.ERR_INCORRECT_INPUT
    assert L.eqs_device_cos(c, None, 5, None) == _ffi.ERR_INCORRECT_INPUT
    assert L.eqs_philox_blocks(c, None, None, 4, None) == _ffi.ERR_INCORRECT_INPUT
    assert L.eqs_draws(c, 0, 1, 0, 0, 4, 3, None, None) == _ffi.ERR_INCORRECT_INPUT
    assert L.eqs_pso_solve(c, None, None, None, None) == _ffi.ERR_INCORRECT_INPUT
    assert L.eqs_ce_solve(c, None, None, None, None) == _ffi.ERR_INCORRECT_INPUT
    assert L.eqs_mc_integrate(c, None, None, None, None) == _ffi.ERR_INCORRECT_INPUT
    assert L.eqs_miser_integrate(c, None, None, None, None) == _ffi.ERR_INCORRECT_INPUT
    assert L.eqs_measure_copy(c, 1 << 20, None) == _ffi.ERR_INCORRECT_INPUT
    assert L.eqs_ctx_comm_init(c, 0, 2, None) != _ffi.OK


def test_dynamic_shared_memory_just_below_48k(ctx):
    """Regression (found by the seed sweep): 47-48 KB of dynamic shared memory plus the kernels' static arrays exceeds
    the 48 KB a launch gets without the opt-in attribute."""
    n = 6100                                    # gbest tile 48 800 B
    lower, upper, x0 = [-2.0] * n, [2.0] * n, np.full(n, 0.5)
    for mode in (E.PsoMode.SYNCHRONOUS, E.PsoMode.SEQUENTIAL_EXACT):
        for N in (64, 3000):                    # single-CTA kernel / grid kernels
            g, o = make(E.Objective.ROSENBROCK, n, N, mode, seed=3, lo=-2.0, hi=2.0)
            run = g.with_context(ctx).start(x0); o.init(x0)
            run.step(2); [o.step() for _ in range(2)]
            compare_state(run, o, True, f"n={n} N={N} {mode.name}")
            run.close()
    ce = (E.CrossEntropy.new(E.Objective.ROSENBROCK).with_sample_size(646).with_importance_selection_size(286)
          .with_iter_max(3).with_std_dev(np.full(19, 1.5)).with_mean_divisor(E.MeanDivisor.ELITE_COUNT).with_seed(5)
          .with_tol(0.0).with_context(ctx))        # single-CTA CE kernel: (2*19 + 646 + 286*19) * 8 = 48 944 B
    oc = O.Ce(O.ROSENBROCK, np.zeros(19), sample_size=646, elite_count=286, iter_max=3, std_dev=np.full(19, 1.5),
              mean_divisor=O.DIV_ELITE_COUNT, seed=5, tol=0.0)
    np.testing.assert_allclose(ce.solve(np.zeros(19)), oc.solve(), rtol=1e-10, atol=1e-12)
    r = E.MonteCarlo.new(E.Objective.SPHERE).with_sample_count(5000).with_context(ctx).integrate_with_seed([0.0] * 3050, [1.0] * 3050, 1)
    assert abs(r.mean - 3050 / 3.0) < 5.0       # 2 * 3050 * 8 = 48 800 B of box data in shared memory


@pytest.mark.parametrize("N,host_driven", [(500, False), (3000, False), (3000, True)])
def test_sequential_mode_nan_pbest_cost_never_updates(ctx, N, host_driven):
    """particle_swarm.rs:417-425 nests the gbest test inside `cost < best_cost`: a particle whose FIRST cost is NaN keeps a
    NaN pbest cost, so it can never improve its pbest or move gbest, however good its later positions are.  SERIES(x) =
    sum (-1)^i x^i is NaN for x > 2.03 (inf - inf), so on [-1, 3]^2 four particles in ten start like that.  All three
    sequential paths (single-CTA, windowed device-driven, host-driven rounds), bit for bit against the oracle."""
    n, iters = 2, 6
    lower, upper, x0 = [-1.0] * n, [3.0] * n, np.full(n, 0.25)
    if host_driven:
        os.environ["EQS_PSO_SEQ_HOST"] = "1"
    try:
        run = (E.ParticleSwarm.new(E.Objective.SERIES, lower, upper).with_particle_count(N).with_tol(0.0).with_seed(77)
               .with_mode(E.PsoMode.SEQUENTIAL_EXACT).with_context(ctx)).start(x0)
    finally:
        os.environ.pop("EQS_PSO_SEQ_HOST", None)
    o = O.Pso(O.SERIES, lower, upper, particle_count=N, tol=0.0, mode=O.SEQUENTIAL, seed=77)
    o.init(x0)
    pbc0 = run.pbest_costs()
    assert 0.2 * N < np.isnan(pbc0).sum() < 0.7 * N          # the case under test is there
    for t in range(iters):
        run.step(1)
        o.step()
        g, c, it, _, imp = run.gbest()
        assert_bit_equal(g, o.gbest(), f"gbest pass {t}")
        assert c == o.gbest_cost()
        for a, b, nm in zip(run.state(), o.state(), ("x", "v", "pbest", "pbest cost")):
            if nm == "pbest cost":   # a NaN's sign and payload are not part of the contract (x86: inf - inf = -NaN)
                assert np.array_equal(np.isnan(a), np.isnan(b)), f"{nm} pass {t}"
                a, b = np.where(np.isnan(a), 0.0, a), np.where(np.isnan(b), 0.0, b)
            assert_bit_equal(a, b, f"{nm} pass {t}")
    assert np.array_equal(np.isnan(run.pbest_costs()), np.isnan(pbc0))   # NaN pbest costs stayed NaN, nobody else got one
    run.close()
