#!/usr/bin/env python3
"""The reference's own benchmark groups for this path, reference benchmarks/multi.rs:102-183:

    "30D Rastrigin":  PSO (init 0.4, bounds +-1) and CE (init 0.4, std_dev 1)        multi.rs:151-183
    "100D Heavy"   :  PSO (init 0.3, bounds [0.3, 0.7]) and CE (init 0.3, std_dev 0.3)  multi.rs:102-149

Like criterion's `bh.iter(|| Solver::new(..).solve(black_box(guess)).unwrap())` (bench_pso! :70-84, bench_ce! :86-100)
every sample CONSTRUCTS the solver and runs a WHOLE DEFAULT solve (100 particles / samples, <= 50 passes, tol 1e-6).
Reported: median wall time per solve through the public builder call (host buffers in and out), for the GPU library
and for the CPU oracle (C restatement of the reference, 1 thread).  This is the latency regime: one SM's worth of
work, so the GPU number is launch + synchronisation cost, not bandwidth.

    python benchmarks/multi.py [--samples 200] [--json]
"""
import argparse
import json
import os
import statistics
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timeit(fn, samples, warmup=5):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(samples):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return statistics.median(ts), min(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=200)
    ap.add_argument("--json", action="store_true")
    a = ap.parse_args()
    import eqsolver_b200 as E
    # the CPU restatement is the reported baseline timed BESIDE the GPU arm, like bench.py's cpu_baseline leg: the GPU arm
    # never calls it, and nothing in the eqsolver_b200 package imports oracle/
    from oracle import oracle as O
    ctx = E.Context(0)

    groups = [
        ("30D Rastrigin", E.Objective.RASTRIGIN, O.RASTRIGIN, 30, 0.4, -1.0, 1.0, 1.0),
        ("100D Heavy", E.Objective.HEAVY, O.HEAVY, 100, 0.3, 0.3, 0.7, 0.3),
    ]
    rows = []
    for name, gobj, oobj, n, init, lo, hi, sd in groups:
        guess = np.full(n, init)
        lower, upper, std = np.full(n, lo), np.full(n, hi), np.full(n, sd)
        seed = [0]

        def gpu_pso(mode):
            seed[0] += 1
            return (E.ParticleSwarm.new(gobj, lower, upper).with_seed(seed[0]).with_mode(mode).with_context(ctx)
                    .solve(guess))

        def cpu_pso():
            seed[0] += 1
            return O.Pso(oobj, lower, upper, seed=seed[0], mode=O.SEQUENTIAL).solve(guess)

        def gpu_ce():
            seed[0] += 1
            return E.CrossEntropy.new(gobj).with_std_dev(std).with_seed(seed[0]).with_context(ctx).solve(guess)

        def cpu_ce():
            seed[0] += 1
            return O.Ce(oobj, guess, std_dev=std, seed=seed[0]).solve()

        for label, fn in (("PSO  gpu sequential-exact (reference semantics)", lambda: gpu_pso(E.PsoMode.SEQUENTIAL_EXACT)),
                          ("PSO  gpu synchronous", lambda: gpu_pso(E.PsoMode.SYNCHRONOUS)),
                          ("PSO  cpu oracle (1 thread)", cpu_pso),
                          ("CE   gpu", gpu_ce),
                          ("CE   cpu oracle (1 thread)", cpu_ce)):
            med, best = timeit(fn, a.samples)
            rows.append({"group": name, "bench": label, "median_us": med * 1e6, "min_us": best * 1e6})
    if a.json:
        print(json.dumps(rows))
    else:
        for r in rows:
            print(f"{r['group']:14s} {r['bench']:50s} median {r['median_us']:10.1f} us   min {r['min_us']:10.1f} us")


if __name__ == "__main__":
    main()
