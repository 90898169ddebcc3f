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


if __name__ == "__main__":
    main()
