#!/usr/bin/env python3
"""MonteCarlo throughput on one GPU (samples/s through the public API, host buffers in, two doubles out).

    python benchmarks/mc_bench.py [--count 2**28] [--reps 5]

The reference benchmark (benches/integrators.rs) integrates its two test integrands at the default 1000 samples; the
same two are timed here at 1000 (latency) and at --count (throughput), beside the CPU oracle on a bounded sample.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--count", type=int, default=1 << 28)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    import eqsolver_b200 as E
    # the CPU restatement is the reported baseline timed BESIDE the GPU arm, like bench.py's cpu_baseline leg: the GPU arm
    # never calls it, and nothing in the eqsolver_b200 package imports oracle/
    from oracle import oracle as O
    ctx = E.Context(0)
    for name, integrand, n in [("gaussian_1d", E.Objective.GAUSSIAN, 1), ("sincos_2d", E.Objective.SINCOS, 2),
                               ("rastrigin_64d", E.Objective.RASTRIGIN, 64)]:
        lower, upper = [0.0] * n, [1.0] * n
        out = {"integrand": name, "n": n}
        for count in (1000, a.count // max(1, n // 2)):
            mc = E.MonteCarlo.new(integrand).with_sample_count(count).with_context(ctx)
            mc.integrate_with_seed(lower, upper, 1)
            best, kern = 1e30, 1e30
            for r in range(a.reps):
                t0 = time.perf_counter()
                res = mc.integrate_with_seed(lower, upper, 2 + r)
                best = min(best, time.perf_counter() - t0)
                kern = min(kern, mc.last_report.loop_ms)
            out[f"count_{count}"] = {"wall_ms": best * 1e3, "kernel_ms": kern, "samples_per_s": count / best,
                                     "mean": res.mean, "std_err": (res.variance / count) ** 0.5}
        cpu_count = 2_000_000 // n
        t0 = time.perf_counter()
        O.mc_integrate(int(integrand), lower, upper, cpu_count, 1)
        out["cpu_oracle_samples_per_s"] = cpu_count / (time.perf_counter() - t0)
        print(json.dumps(out))

    # the reference's own integrator benchmark group, benchmarks/integrators.rs:67-78 "1D Heavy Integrators":
    # MonteCarlo::new(f).integrate(0., 0.5) at the default 1000 samples, f = sum_{i<1000} (-1)^i x^i
    def best_of(fn, reps=a.reps):
        fn()
        best = 1e30
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            best = min(best, time.perf_counter() - t0)
        return best
    mc = E.MonteCarlo.new(E.Objective.SERIES).with_context(ctx)
    group = {"group": "1D Heavy Integrators (benchmarks/integrators.rs:67-78)", "exact": 0.4054651081081644}
    group["Monte Carlo"] = {"gpu_ms": 1e3 * best_of(lambda: mc.integrate(0.0, 0.5)),
                            "cpu_oracle_ms": 1e3 * best_of(lambda: O.mc_integrate(O.SERIES, [0.0], [0.5], 1000, 1), 2),
                            "mean": mc.integrate(0.0, 0.5)}
    mi = E.MISER.new(E.Objective.SERIES).with_context(ctx)
    group["MISER"] = {"gpu_ms": 1e3 * best_of(lambda: mi.integrate(0.0, 0.5)),
                      "cpu_oracle_ms": 1e3 * best_of(lambda: O.miser_integrate(O.SERIES, [0.0], [0.5]), 2),
                      "mean": mi.integrate(0.0, 0.5), "evals": mi.last_report.evals, "levels": mi.last_report.iters_run}
    print(json.dumps(group))
    # MISER at the reference test's shape (tests/integrators.rs:108-116) and at GPU-sized point counts
    for count in (1000, 1 << 20, 1 << 24):
        mi = E.MISER.new(E.Objective.SINCOS).with_sample_count(count).with_context(ctx)
        t = best_of(lambda: mi.integrate_with_seed([0.0, 0.0], [1.0, 1.0], 3))
        r = mi.integrate_with_seed([0.0, 0.0], [1.0, 1.0], 3)
        row = {"miser": "sincos_2d", "sample_count": count, "wall_ms": 1e3 * t, "kernel_ms": mi.last_report.loop_ms,
               "evals": mi.last_report.evals, "evals_per_s": mi.last_report.evals / t, "error": r.mean - 0.9248464799519,
               "sigma": r.variance ** 0.5}
        if count <= (1 << 20):
            cc = min(count, 1 << 16)
            t0 = time.perf_counter()
            ev = O.miser_integrate(O.SINCOS, [0.0, 0.0], [1.0, 1.0], sample_count=cc, seed=3)[3]
            row["cpu_oracle_evals_per_s"] = ev / (time.perf_counter() - t0)
        print(json.dumps(row))


if __name__ == "__main__":
    main()
