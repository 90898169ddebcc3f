"""CPU: the C oracle (oracle/eqs_oracle.c) against the committed golden fixtures (tests/golden/make_golden.py)."""
import numpy as np

from conftest import assert_bit_equal, from_bits, golden
from oracle import oracle as O


def test_objective_kats_bit_exact():
    for e in golden("objective_kats.json"):
        got = O.objective(e["objective"], e["x"])
        assert_bit_equal([got], [from_bits(e["bits"])], e["note"])


def test_survey_objective_anchors():
    # SURVEY.md §4 item 3 (computed there with glibc libm): the reference's own inputs
    assert O.objective(O.RASTRIGIN, [80., -81., 82., -83., 84., -85., 86., -87., 88., -89.]) == 71485.0
    assert O.objective(O.RASTRIGIN, [0.4] * 30) == 547.5050983124845
    assert O.objective(O.RASTRIGIN, [80.] * 16) == 102400.0
    assert O.objective(O.ROSENBROCK, [-1.2, 1.0]) == 24.199999999999996
    assert O.objective(O.ACKLEY, [0.0] * 256) == 4.440892098500626e-16
    assert O.objective(O.ACKLEY, [1.0] * 256) == 3.6253849384403627
    # summation order of nalgebra's norm is third-party (unverified): tolerance, not bits
    assert abs(O.objective(O.HEAVY, [0.3] * 100) - 52.351074487540366) < 1e-13
    # reference tests/multi.rs:70: the three-circles least-squares point; it must be a local minimum
    sol = np.array([4.217265312839526, 2.317879970005811])
    f0 = O.objective(O.THREE_CIRCLES, sol)
    for dx in (1e-4, -1e-4):
        assert O.objective(O.THREE_CIRCLES, sol + [dx, 0]) > f0
        assert O.objective(O.THREE_CIRCLES, sol + [0, dx]) > f0


def test_philox_random123_kats():
    g = golden("philox_kats.json")
    for v in g["random123"]:
        assert O.philox4x32_10(v["ctr"], v["key"]) == v["out"]


def test_philox_draw_contract():
    M = 0xFFFFFFFF
    for v in golden("philox_kats.json")["contract"]:
        ctr = [v["gidx"] & M, (v["gidx"] >> 32) & M, ((v["stream"] << 28) | v["j"]) & M, v["t"] & M]
        key = [v["seed"] & M, (v["seed"] >> 32) & M]
        w = O.philox4x32_10(ctr, key)
        assert w == v["words"]
        assert_bit_equal([O.u01(w[0], w[1]), O.u01(w[2], w[3])], [from_bits(v["u_lo"]), from_bits(v["u_hi"])])
    # the oracle's draw helpers follow the same contract
    pos, vel = O.pso_init_draws(1729, 0, 1)
    v0 = golden("philox_kats.json")["contract"][0]
    assert_bit_equal([pos[0], vel[0]], [from_bits(v0["u_lo"]), from_bits(v0["u_hi"])])
    rp, rg = O.pso_step_draws(1729, 7, 123456789, 32)
    v1 = golden("philox_kats.json")["contract"][1]
    assert_bit_equal([rp[31], rg[31]], [from_bits(v1["u_lo"]), from_bits(v1["u_hi"])])


def test_normals_are_standard():
    z = np.concatenate([O.ce_normals(5, 0, i, 64) for i in range(2000)])
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1.0) < 0.02
    assert abs(np.mean(z ** 3)) < 0.05 and abs(np.mean(z ** 4) - 3.0) < 0.1


def test_pso_trajectories_bit_exact():
    for g in golden("pso_golden.json"):
        o = O.Pso(g["objective"], g["lower"], g["upper"], particle_count=g["N"], iter_max=g["iters"], mode=g["mode"],
                  seed=g["seed"])
        o.init(g["x0"])
        tr = g["trace"]
        assert_bit_equal(o.gbest(), [from_bits(b) for b in tr[0]["gbest"]], "gbest after init")
        assert_bit_equal([o.gbest_cost()], [from_bits(tr[0]["gbest_cost"])])
        assert o.ring()[1] == tr[0]["ring_i"]
        for t in range(1, len(tr)):
            assert o.step() == 0
            assert_bit_equal(o.gbest(), [from_bits(b) for b in tr[t]["gbest"]], f"gbest after pass {t}")
            assert_bit_equal([o.gbest_cost()], [from_bits(tr[t]["gbest_cost"])])
            assert o.ring()[1] == tr[t]["ring_i"], f"ring index after pass {t}"
        x, v, pb, pbc = o.state()
        assert_bit_equal(x, [[from_bits(b) for b in row] for row in g["final_x"]], "positions")
        assert_bit_equal(v, [[from_bits(b) for b in row] for row in g["final_v"]], "velocities")
        assert_bit_equal(pbc, [from_bits(b) for b in g["final_pbest_cost"]], "pbest costs")
        assert_bit_equal(o.ring()[0], [from_bits(b) for b in g["ring"]], "stall ring")


def test_ce_trajectories_bit_exact():
    for g in golden("ce_golden.json"):
        z = np.array([[[from_bits(b) for b in row] for row in it] for it in g["normals"]])
        o = O.Ce(g["objective"], g["x0"], sample_size=g["N"], elite_count=g["k"], iter_max=g["iter_max"],
                 std_dev=g["sigma0"], replay_normals=z,
                 mean_divisor=O.DIV_REFERENCE_TEN if g["divisor_ten"] else O.DIV_ELITE_COUNT)
        for t, tr in enumerate(g["trace"]):
            assert o.step() == 0
            mu, sg = o.state()
            assert_bit_equal(mu, [from_bits(b) for b in tr["mu"]], f"mu after pass {t}")
            assert_bit_equal(sg, [from_bits(b) for b in tr["sigma"]], f"sigma after pass {t}")
            assert list(o.elites()) == tr["elites"]
        assert o.step() == 1  # iter reached iter_max (cross_entropy.rs:255)


def test_pso_sharded_protocol_equals_unsharded():
    """The exchange protocol on 3 virtual shards reproduces the unsharded synchronous oracle."""
    n, N, iters, world = 5, 97, 8, 3
    lo, hi, x0 = [-5.0] * n, [10.0] * n, [3.0] * n
    ref = O.Pso(O.ROSENBROCK, lo, hi, particle_count=N, iter_max=iters, mode=O.SYNCHRONOUS, seed=3, tol=0.0)
    ref.init(x0)
    shards = []
    for r in range(world):
        base, rem = divmod(N, world)
        first = r * base + min(r, rem)
        shards.append(O.Pso(O.ROSENBROCK, lo, hi, particle_count=N, iter_max=iters, mode=O.SYNCHRONOUS, seed=3,
                            tol=0.0, first_index=first, local_count=base + (1 if r < rem else 0)))
    recs = np.stack([s.init_local(x0) for s in shards])
    f0 = O.objective(O.ROSENBROCK, x0)
    pushes, tails = [], []
    for r, s in enumerate(shards):
        seed_cost = min([f0] + [recs[q, 0] for q in range(r) if recs[q, 1] >= 0])
        k, tail = s.init_ring_local(seed_cost)
        pushes.append(k); tails.append(tail)
    for s in shards:
        s.init_apply(recs, pushes, np.stack(tails))
        assert_bit_equal(s.gbest(), ref.gbest()); assert_bit_equal(s.ring()[0], ref.ring()[0])
        assert s.ring()[1] == ref.ring()[1]
    for _ in range(iters):
        ref.step()
        out = [s.step_local() for s in shards]
        recs = np.stack([o[1] for o in out])
        for s in shards:
            s.step_apply(recs)
            assert_bit_equal(s.gbest(), ref.gbest()); assert_bit_equal([s.gbest_cost()], [ref.gbest_cost()])
    x = np.concatenate([s.state()[0] for s in shards])
    assert_bit_equal(x, ref.state()[0])


def test_pso_validate_like_new():
    import pytest
    with pytest.raises(ValueError):
        O.Pso(O.SPHERE, [1.0, 0.0], [0.0, 1.0])      # lower > upper -> Uniform::new_inclusive error
    with pytest.raises(ValueError):
        O.Pso(O.SPHERE, [0.0], [float("inf")])


def test_known_optima():
    x = O.Pso(O.SPHERE, [-5] * 4, [5] * 4, particle_count=400, iter_max=200, seed=1, tol=0.0).solve([4.0] * 4)
    assert O.objective(O.SPHERE, x) < 1e-12
    x = O.Pso(O.THREE_CIRCLES, [1, 1], [7, 7], particle_count=200, iter_max=200, seed=2, tol=0.0).solve([4.5, 2.5])
    assert np.allclose(x, [4.217265312839526, 2.317879970005811], atol=1e-5)
    mu = O.Ce(O.SPHERE, [3.0] * 4, sample_size=400, elite_count=40, iter_max=80, std_dev=[2.0] * 4,
              mean_divisor=O.DIV_ELITE_COUNT, seed=3).solve()
    assert O.objective(O.SPHERE, mu) < 1e-8


def test_mc_oracle_against_pure_python_welford():
    """orc_mc_integrate (lib.rs:172-202, monte_carlo.rs:143-212) against a pure-Python restatement, bit for bit."""
    import math
    n, count, seed = 3, 257, 5
    lower, upper = [-1.0, 0.0, 2.0], [1.0, 0.5, 3.5]
    mean = ssq = 0.0
    for i in range(count):
        u = O.mc_draws(seed, i, n)
        x = [u[d] * (upper[d] - lower[d]) + lower[d] for d in range(n)]
        s = 0.0
        for d in range(n - 1):
            a = x[d + 1] - x[d] * x[d]; b = 1.0 - x[d]
            s += 100.0 * (a * a) + b * b
        delta = s - mean
        mean = mean + delta / float(i + 1)
        ssq = ssq + delta * (s - mean)
    vol = 1.0
    for d in range(n):
        vol *= upper[d] - lower[d]
    st, m, v = O.mc_integrate(O.ROSENBROCK, lower, upper, count, seed)
    assert st == 0
    assert_bit_equal([m, v], [mean * vol, ssq / (count - 1) * vol * vol])
    assert O.mc_integrate(O.SPHERE, [1.0], [0.0], 10, 1)[0] == O.ERR_EXTERNAL
    assert O.mc_integrate(O.SPHERE, [0.0], [1.0], 1, 1)[0] == O.ERR_INCORRECT_INPUT
    # the reference's own tests (tests/integrators.rs:87-106): 1000 samples, tolerance 0.1
    assert abs(O.mc_integrate(O.GAUSSIAN, [0.0], [1.0], 1000, 1729)[1] - 0.746824132812427025) <= 0.1
    assert abs(O.mc_integrate(O.SINCOS, [0.0, 0.0], [1.0, 1.0], 1000, 1729)[1] - 0.9248464799519) <= 0.1


def test_all_cores_pass_is_bit_identical_to_the_scalar_pass():
    """orc_pso_step_mt (threads over particles, SURVEY 8d's all-cores baseline) == orc_pso_step, bit for bit."""
    n, N = 7, 1031
    lower, upper, x0 = [-5.0] * n, [10.0] * n, np.full(n, 2.0)
    a = O.Pso(O.ROSENBROCK, lower, upper, particle_count=N, tol=0.0, mode=O.SYNCHRONOUS, seed=11)
    b = O.Pso(O.ROSENBROCK, lower, upper, particle_count=N, tol=0.0, mode=O.SYNCHRONOUS, seed=11)
    a.init(x0); b.init(x0)
    for t in range(6):
        assert a.step() == 0 and b.step_mt(1 + t) == 0
        for u, w in zip(a.state(), b.state()):
            assert_bit_equal(u, w)
        assert_bit_equal(a.gbest(), b.gbest())
        assert a.gbest_cost() == b.gbest_cost() and a.iters() == b.iters()
    c = O.Pso(O.ROSENBROCK, lower, upper, particle_count=N, tol=0.0, mode=O.SEQUENTIAL, seed=11)
    c.init(x0)
    assert c.step_mt(4) == -1      # the reference's sequential gbest semantics cannot be threaded


def test_miser_oracle_against_pure_python_recursion():
    """orc_miser_integrate (miser.rs:301-443) against a pure-Python restatement of the recursion, bit for bit, plus
    the reference's own acceptance test (tests/integrators.rs:108-116: |mean - 0.924846| <= 0.001 at the defaults)."""
    import math
    seed, n, min_dim, alpha, dither = 1729, 2, 32, 2.0, 0.05
    key = [seed & 0xFFFFFFFF, seed >> 32]
    spent = [0]

    def draw(node, gidx, j):
        w = O.philox4x32_10([gidx & 0xFFFFFFFF, gidx >> 32, (4 << 28) | j, node], key)
        return O.u01(int(w[0]), int(w[1])), O.u01(int(w[2]), int(w[3]))

    def batch(lo, hi, count, node, gbase, bessel):
        mean = ssq = 0.0
        for i in range(count):
            u0, u1 = draw(node, gbase | i, 0)
            x0, x1 = u0 * (hi[0] - lo[0]) + lo[0], u1 * (hi[1] - lo[1]) + lo[1]
            s = math.sin(math.cos(x1) + x0)
            d = s - mean
            mean = mean + d / float(i + 1)
            ssq = ssq + d * (s - mean)
        spent[0] += count
        return mean, ssq / float(count - 1 if bessel else count)

    def recurse(lo, hi, node, depth, count):
        if depth == 0 or count < min_dim * n:
            count = max(count, min_dim * n)
            vol = 1.0
            for d in range(n):
                vol = vol * (hi[d] - lo[d])
            m, v = batch(lo, hi, count, node, (2 * n + 1) << 40, False)
            return m * vol, v * vol * vol
        count = max(count, 2)
        best = (1.7976931348623157e308,) * 3
        bd = min(n - 1, int(draw(node, (2 * n) << 40, 0)[0] * n))
        bmid = lo[bd] + (hi[bd] - lo[bd]) * 0.5
        for d in range(n):
            r = draw(node, ((2 * n) << 40) | (1 + d), 0)[0] * (dither - (-dither)) + (-dither)
            mid = lo[d] + (hi[d] - lo[d]) * (0.5 + r)
            h2 = list(hi); h2[d] = mid
            l2 = list(lo); l2[d] = mid
            vl = batch(lo, h2, count, node, (2 * d) << 40, True)[1]
            vu = batch(l2, hi, count, node, (2 * d + 1) << 40, True)[1]
            if vl + vu < best[0]:
                best, bd, bmid = (vl + vu, vl, vu), d, mid
        beta = 1.0 / (1.0 + alpha)
        la, ua = best[1] ** beta, best[2] ** beta
        nl, nu = int(float(count) * (la / (la + ua))), int(float(count) * (ua / (la + ua)))
        h2 = list(hi); h2[bd] = bmid
        l2 = list(lo); l2[bd] = bmid
        a = recurse(lo, h2, 2 * node, depth - 1, nl)
        b = recurse(l2, hi, 2 * node + 1, depth - 1, nu)
        return a[0] + b[0], a[1] + b[1]

    want = recurse([0.0, 0.0], [1.0, 1.0], 1, 3, 400)
    st, mean, var, evals = O.miser_integrate(O.SINCOS, [0.0, 0.0], [1.0, 1.0], sample_count=400, maximum_depth=3, seed=seed)
    assert st == 0 and evals == spent[0]
    assert_bit_equal([mean, var], list(want))
    st, mean, var, _ = O.miser_integrate(O.SINCOS, [0.0, 0.0], [1.0, 1.0], seed=1729)
    assert st == 0 and abs(mean - 0.9248464799519) <= 0.001
    assert O.miser_integrate(O.SPHERE, [0.0, 0.0], [1.0, -1.0])[0] == O.ERR_EXTERNAL
    assert O.miser_integrate(O.SPHERE, [0.0, 0.0], [0.0, 0.0])[0] == O.ERR_TYPE_CONVERSION
    assert O.miser_integrate(O.SPHERE, [0.0, 0.0], [1.0, 1.0], min_dim_count=0, sample_count=1, maximum_depth=0)[0] \
        == O.ERR_INCORRECT_INPUT
