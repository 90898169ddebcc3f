#!/usr/bin/env python3
"""Generates the committed golden fixtures in this directory.  Run: python tests/golden/make_golden.py

The reference (AzeezDa/eqsolver v0.4.1) cannot run in this image (no cargo/rustc) and has no golden vectors for
ParticleSwarm / CrossEntropy (its RNG is unseedable, its tests assert nothing on this path), so
these fixtures come from an INDEPENDENT pure-Python restatement written straight from the Rust source:
  * objective known-answer values, computed with Python floats (IEEE f64, libm cos/exp/sqrt) in the operation
    order of examples/global_optimisation.rs:9-17, benchmarks/multi.rs:106-115,137,
    examples/global_optimiser_as_solver.rs:22-42 and the textbook Sphere/Rosenbrock/Ackley;
  * the three Random123 Philox4x32-10 known-answer vectors plus blocks from a pure-Python Philox;
  * short ParticleSwarm (sequential = reference semantics, and synchronous) and CrossEntropy trajectories that
    follow particle_swarm.rs:356-441 and cross_entropy.rs:245-306 line by line, on Sphere / Rosenbrock
    (only + - * / sqrt: bit-reproducible everywhere).
Python never contracts a*b+c, like rustc.  The C oracle (oracle/eqs_oracle.c) must reproduce every number here
bit for bit; the CUDA path is then checked against the oracle.
"""
import json
import math
import os
import struct

HERE = os.path.dirname(os.path.abspath(__file__))
M32 = 0xFFFFFFFF


def bits(x: float) -> str:
    return "0x%016x" % struct.unpack("<Q", struct.pack("<d", x))[0]


# ---------------------------------------------------------------- objectives
def sphere(x):
    t = 0.0
    for w in x:
        t += w * w
    return t


def rosenbrock(x):
    t = 0.0
    for d in range(len(x) - 1):
        a = x[d + 1] - x[d] * x[d]
        b = 1.0 - x[d]
        t += 100.0 * (a * a) + b * b
    return t


def rastrigin(x):
    t = 10.0 * float(len(x))
    for w in x:
        t += w * w - 10.0 * math.cos(2.0 * math.pi * w)
    return t


def ackley(x):
    n = float(len(x))
    s2 = 0.0
    sc = 0.0
    for w in x:
        s2 += w * w
        sc += math.cos(2.0 * math.pi * w)
    a = -20.0 * math.exp(-0.2 * math.sqrt(s2 / n))
    b = math.exp(sc / n)
    return a - b + 20.0 + math.e


def three_circles(x):
    cs = [(3., 5., 3.), (1., 0., 4.), (6., 2., 2.)]
    acc = 0.0
    for cx, cy, r in cs:
        e = ((x[0] - cx) * (x[0] - cx) + (x[1] - cy) * (x[1] - cy)) - r * r
        acc += e * e
    return math.sqrt(acc)


def heavy(x):
    acc = 0.0
    for i in range(len(x)):
        o = 0.0
        for j in range(i, len(x)):
            o += x[j] * x[j]
        acc += o * o
    return math.sqrt(acc)


def zakharov(x):
    s1 = 0.0
    s2 = 0.0
    for d, w in enumerate(x):
        s1 += w * w
        s2 += (0.5 * float(d + 1)) * w
    q = s2 * s2
    return s1 + q + q * q


def gaussian(x):  # tests/integrators.rs:11-13, squares summed over the dimensions
    s2 = 0.0
    for w in x:
        s2 += w * w
    return math.exp(-s2)


def sincos(x):  # tests/integrators.rs:15-17
    return math.sin(math.cos(x[1]) + x[0])


def powi(a, b):  # f64::powi with a run-time exponent (compiler-rt __powidf2), b >= 0
    r = 1.0
    while True:
        if b & 1:
            r *= a
        b //= 2
        if b == 0:
            return r
        a *= a


def series(x):  # benchmarks/integrators.rs:69-75
    total = 0.0
    for w in x:
        s = 0.0
        for i in range(1000):
            s += powi(-1.0, i) * powi(w, i)
        total += s
    return total


OBJECTIVES = {0: sphere, 1: rosenbrock, 2: rastrigin, 3: ackley, 4: three_circles, 5: heavy, 6: zakharov,
              7: gaussian, 8: sincos, 9: series}


# ---------------------------------------------------------------- Philox4x32-10
def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & M32, p1 & M32, ((p0 >> 32) ^ c3 ^ k1) & M32, p0 & M32
        k0 = (k0 + 0x9E3779B9) & M32
        k1 = (k1 + 0xBB67AE85) & M32
    return [c0, c1, c2, c3]


def u01(w_lo, w_hi):
    return float(((w_hi & 0xFFFFF) << 32) | w_lo) * 2.0 ** -52


def draw(seed, stream, t, gidx, j):
    return philox4x32_10([gidx & M32, (gidx >> 32) & M32, ((stream << 28) | j) & M32, t & M32],
                         [seed & M32, (seed >> 32) & M32])


# ---------------------------------------------------------------- ParticleSwarm, particle_swarm.rs:356-441
def pso_trajectory(obj_id, n, N, iters, lower, upper, x0, seed, mode, w=0.5, c1=1.0, c2=1.0, tol=1e-6):
    f = OBJECTIVES[obj_id]
    G = list(x0)
    Gc = f(G)
    ring = [math.inf] * 50
    ring_i = 0
    X, V, PB, PBC = [], [], [], []
    for i in range(N):
        pos, vel = [], []
        us = [draw(seed, 0, 0, i, d) for d in range(n)]
        for d in range(n):
            pos.append(u01(us[d][0], us[d][1]) * (upper[d] - lower[d]) + lower[d])
        for d in range(n):
            dist = abs(upper[d] - lower[d])
            vel.append(u01(us[d][2], us[d][3]) * (dist + dist) + (-dist))
        cost = f(pos)
        if cost < Gc:
            ring[ring_i] = Gc
            ring_i = (ring_i + 1) % 50
            G = list(pos)
            Gc = cost
        X.append(pos); V.append(vel); PB.append(list(pos)); PBC.append(cost)
    trace = [dict(gbest=list(G), gbest_cost=Gc, ring_i=ring_i)]
    for t in range(iters):
        if all(abs(r - Gc) < tol for r in ring):
            break
        best = (math.inf, -1)
        for i in range(N):
            for d in range(n):
                wds = draw(seed, 1, t, i, d)
                rp, rg = u01(wds[0], wds[1]), u01(wds[2], wds[3])
                V[i][d] = w * V[i][d] + c1 * rp * (PB[i][d] - X[i][d]) + c2 * rg * (G[d] - X[i][d])
            for d in range(n):
                X[i][d] += V[i][d]
            cost = f(X[i])
            if cost < PBC[i]:
                PB[i] = list(X[i])
                PBC[i] = cost
                if mode == 0 and PBC[i] < Gc:
                    ring[ring_i] = Gc
                    ring_i = (ring_i + 1) % 50
                    G = list(X[i])
                    Gc = PBC[i]
            if cost < best[0]:
                best = (cost, i)
        if mode == 1 and best[0] < Gc:
            ring[ring_i] = Gc
            ring_i = (ring_i + 1) % 50
            G = list(X[best[1]])
            Gc = best[0]
        trace.append(dict(gbest=list(G), gbest_cost=Gc, ring_i=ring_i))
    return dict(objective=obj_id, n=n, N=N, iters=iters, lower=lower, upper=upper, x0=x0, seed=seed, mode=mode,
                trace=[dict(gbest=[bits(v) for v in s["gbest"]], gbest_cost=bits(s["gbest_cost"]), ring_i=s["ring_i"])
                       for s in trace],
                final_x=[[bits(v) for v in row] for row in X], final_v=[[bits(v) for v in row] for row in V],
                final_pbest_cost=[bits(v) for v in PBC],
                ring=[("inf" if math.isinf(r) else bits(r)) for r in ring])


# ---------------------------------------------------------------- CrossEntropy, cross_entropy.rs:245-306
def ce_trajectory(obj_id, n, N, k, iter_max, x0, sigma0, normals, divisor_ten, tol=1e-6):
    f = OBJECTIVES[obj_id]
    mus = list(x0)
    sigmas = list(sigma0)
    it = 1
    trace = []
    while max(abs(s) for s in sigmas) > tol and it < iter_max:
        pairs = []
        for s in range(N):
            z = normals[it - 1][s]
            x = [mus[d] + sigmas[d] * z[d] for d in range(n)]
            pairs.append((x, f(x), s))
        pairs.sort(key=lambda p: (p[1], p[2]))
        new_mu, new_sig = [], []
        for d in range(n):
            mu = 0.0
            for j in range(k):
                mu += pairs[j][0][d]
            mu /= 10.0 if divisor_ten else float(k)
            sg = 0.0
            for j in range(k):
                e = pairs[j][0][d] - mu
                sg += e * e
            sg /= float(k - 1)
            new_mu.append(mu)
            new_sig.append(math.sqrt(sg))
        mus, sigmas = new_mu, new_sig
        trace.append(dict(mu=[bits(v) for v in mus], sigma=[bits(v) for v in sigmas],
                          elites=[p[2] for p in pairs[:k]]))
        it += 1
    return trace


def main():
    # ---- objective KATs (points from SURVEY.md §4: reference example/benchmark/doctest inputs)
    kat_points = [
        (2, [80., -81., 82., -83., 84., -85., 86., -87., 88., -89.], "examples/global_optimisation.rs:19 guess"),
        (2, [0.4] * 30, "benchmarks/multi.rs:176 init"),
        (2, [80.] * 16, "particle_swarm.rs doctest guess"),
        (2, [5.12, -5.12], "rastrigin corner"),
        (2, [0.0] * 8, "rastrigin optimum"),
        (1, [0.0] * 32, "rosenbrock origin"),
        (1, [-1.2, 1.0], "rosenbrock classic start"),
        (1, [1.0] * 32, "rosenbrock optimum"),
        (3, [0.0] * 256, "ackley optimum (residual 2^-51)"),
        (3, [1.0] * 256, "ackley ones"),
        (0, [1.0, 2.0, 3.0], "sphere"),
        (4, [4.217265312839526, 2.317879970005811], "three circles optimum, tests/multi.rs:70"),
        (4, [4.5, 2.5], "three circles guess, examples/global_optimiser_as_solver.rs:42"),
        (5, [0.3] * 100, "benchmarks/multi.rs:129 init (Heavy)"),
        (6, [1.0, 1.0, 1.0], "zakharov (shipped user hook example)"),
        (6, [0.25, -0.5, 0.75, 1.5], "zakharov"),
        (7, [0.5], "gaussian, tests/integrators.rs:11"),
        (7, [0.25, -1.5, 0.75], "gaussian, 3-D"),
        (8, [0.3, 0.9], "sin(cos y + x), tests/integrators.rs:15"),
        (9, [0.25], "alternating power series, benchmarks/integrators.rs:69 (1/(1+x) = 0.8)"),
        (9, [0.5, 0.125], "alternating power series, 2-D sum"),
    ]
    obj = [dict(objective=o, x=x, value=repr(OBJECTIVES[o](x)), bits=bits(OBJECTIVES[o](x)), note=note)
           for o, x, note in kat_points]
    json.dump(obj, open(os.path.join(HERE, "objective_kats.json"), "w"), indent=1)

    # ---- Philox: Random123 kat_vectors (philox4x32 10 rounds) + the draw contract
    r123 = [
        dict(ctr=[0, 0, 0, 0], key=[0, 0], out=[0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        dict(ctr=[M32] * 4, key=[M32] * 2, out=[0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        dict(ctr=[0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], key=[0xa4093822, 0x299f31d0],
             out=[0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for v in r123:
        assert philox4x32_10(v["ctr"], v["key"]) == v["out"], "pure-Python Philox disagrees with Random123"
    contract = []
    for (seed, stream, t, gidx, j) in [(1729, 0, 0, 0, 0), (1729, 1, 7, 123456789, 31), (2 ** 63 + 5, 2, 3, 2 ** 33 + 1, 511),
                                       (0, 1, 0, 1, 0), (42, 1, 99, 65535, 255)]:
        w = draw(seed, stream, t, gidx, j)
        contract.append(dict(seed=seed, stream=stream, t=t, gidx=gidx, j=j, words=w,
                             u_lo=bits(u01(w[0], w[1])), u_hi=bits(u01(w[2], w[3]))))
    json.dump(dict(random123=r123, contract=contract), open(os.path.join(HERE, "philox_kats.json"), "w"), indent=1)

    # ---- ParticleSwarm trajectories
    pso = [
        pso_trajectory(1, 4, 12, 6, [-5.0] * 4, [10.0] * 4, [0.0] * 4, 1729, 0),
        pso_trajectory(1, 4, 12, 6, [-5.0] * 4, [10.0] * 4, [0.0] * 4, 1729, 1),
        pso_trajectory(0, 3, 9, 5, [-1.0, -2.0, -3.0], [1.0, 2.5, 3.0], [0.5, 0.5, 0.5], 7, 0),
        pso_trajectory(0, 3, 9, 5, [-1.0, -2.0, -3.0], [1.0, 2.5, 3.0], [0.5, 0.5, 0.5], 7, 1),
    ]
    json.dump(pso, open(os.path.join(HERE, "pso_golden.json"), "w"), indent=0)

    # ---- CrossEntropy trajectories (replay normals built from Philox uniforms by a fixed rational map: the
    #      fixture only needs SOME reproducible z stream, its distribution is irrelevant)
    def z_stream(seed, iters, N, n):
        out = []
        for t in range(iters):
            rows = []
            for s in range(N):
                row = []
                for d in range(n):
                    w = draw(seed, 2, t, s, d)
                    row.append((u01(w[0], w[1]) - 0.5) * 4.0 + (u01(w[2], w[3]) - 0.5))
                rows.append(row)
            out.append(rows)
        return out

    ce = []
    for (o, n, N, k, itmax, x0, s0, ten, seed) in [(0, 3, 40, 10, 6, [1.0, -2.0, 0.5], [1.0, 1.0, 1.0], True, 11),
                                                    (1, 4, 64, 8, 5, [0.0] * 4, [2.0] * 4, False, 12),
                                                    (1, 4, 64, 8, 4, [0.0] * 4, [2.0] * 4, True, 12)]:
        zs = z_stream(seed, itmax - 1, N, n)
        ce.append(dict(objective=o, n=n, N=N, k=k, iter_max=itmax, x0=x0, sigma0=s0, divisor_ten=ten,
                       normals=[[[bits(v) for v in row] for row in it] for it in zs],
                       trace=ce_trajectory(o, n, N, k, itmax, x0, s0, zs, ten)))
    json.dump(ce, open(os.path.join(HERE, "ce_golden.json"), "w"), indent=0)
    print("wrote objective_kats.json philox_kats.json pso_golden.json ce_golden.json")


if __name__ == "__main__":
    main()
