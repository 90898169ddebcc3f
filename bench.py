#!/usr/bin/env python3
"""bench.py -- particle-evals/s (f64) of the ParticleSwarm population step on B200, one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): ParticleSwarm, Rosenbrock n=32,
2^20 particles PER GPU (weak scaling: contiguous index blocks per rank, one gbest exchange per pass), f64,
synchronous gbest, device Philox draws, tol = 0 so the stall test never ends the run early.  A "step" is one pass
of the outer loop (particle_swarm.rs:394-428) over the whole population = 2^20 * N particle-evaluations.

  value     evals/s with the population resident in HBM, timed on the device (CUDA events on the library's stream
            around exactly K passes, barrier + synchronize on both sides, max over ranks)
  e2e       evals/s through the public builder call `ParticleSwarm.new(..).solve(x0)` with HOST buffers: allocation,
            H2D of x0/bounds, init, K passes, D2H of the result all inside the timed region (host clock, max over ranks)
  roofline  the step kernel against measured HBM copy bandwidth: 40n+16 algorithmic bytes per particle-evaluation
  cpu_baseline  the C restatement of the reference's algorithm (oracle/, 1 thread -- the reference is single-threaded)
            on a bounded sample of the same workload; a reported baseline, not the target.
--impl reference times that CPU restatement on the same config/metric (the Rust reference cannot be built here).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (objective, n, particles per GPU, lower, upper, x0, description)
    "c2": ("ROSENBROCK", 32, 1 << 20, -5.0, 10.0, 0.0,
           "C2: ParticleSwarm Rosenbrock n=32, 2^20 particles per GPU, f64, synchronous gbest, Philox draws"),
    "c4": ("ACKLEY", 256, 1 << 21, -32.768, 32.768, 20.0,
           "C4: ParticleSwarm Ackley n=256, 2^21 particles per GPU (16M over 8), f64, synchronous gbest"),
    "c1": ("RASTRIGIN", 2, 1000, -5.12, 5.12, 5.0,
           "C1: ParticleSwarm Rastrigin n=2, 1000 particles (latency-bound; examples/global_optimisation.rs style)"),
}
SEED = 1729  # tests/integrators.rs:21


def algorithmic_bytes_per_eval(n: int) -> int:
    """B_pso(n) = 40n + 16: read x, v, pbest; write x, v; read + write pbest cost."""
    return 40 * n + 16


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi sampled every 100 ms while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap,clocks.mem,temperature.gpu")

    def __init__(self, device: int):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.monotonic(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0: float, t1: float, window: str):
        sm, mx, reasons, pw, mem, temp = [], [], set(), [], [], []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if ts < t0 or ts > t1 + 0.15:
                continue
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
            try:
                pw.append(float(parts[2])); mem.append(float(parts[7])); temp.append(float(parts[8]))
            except (ValueError, IndexError):
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "window": window}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "window": window,
                "power_w": float(np.median(pw)) if pw else None, "mem_mhz": float(np.median(mem)) if mem else None,
                "temp_c": float(np.median(temp)) if temp else None}


def cpu_baseline_port(workload: str, budget_s: float):
    """The oracle (C restatement, reference semantics = sequential gbest, 1 thread) on a bounded sample."""
    from oracle import oracle as O
    obj, n, _, lo, hi, x0v, _ = WORKLOADS[workload]
    N = 1 << 16 if WORKLOADS[workload][2] >= (1 << 16) else WORKLOADS[workload][2]
    o = O.Pso(getattr(O, obj), [lo] * n, [hi] * n, particle_count=N, iter_max=10 ** 9, tol=0.0, mode=O.SEQUENTIAL,
              seed=SEED)
    o.init(np.full(n, x0v))
    t1 = o.time_steps(1)                       # calibration pass (also warms the caches)
    iters = max(2, int(budget_s / max(t1, 1e-6)))
    t = o.time_steps(iters)
    # additionally the synchronous pass threaded over the particles on every host core (the fair "per box"
    # figure; the reference itself is single-threaded and sequential, so this one is NOT the reference's algorithm)
    cores = os.cpu_count() or 1
    om = O.Pso(getattr(O, obj), [lo] * n, [hi] * n, particle_count=N, iter_max=10 ** 9, tol=0.0, mode=O.SYNCHRONOUS,
               seed=SEED)
    om.init(np.full(n, x0v))
    tm1 = om.time_steps_mt(1, cores)
    iters_m = max(2, int(0.5 * budget_s / max(tm1, 1e-6)))
    tm = om.time_steps_mt(iters_m, cores)
    return {"value": N * iters / t, "unit": "particle-evals/s", "cores": 1, "kind": "port",
            "all_cores": {"value": N * iters_m / tm, "unit": "particle-evals/s", "cores": cores, "kind": "port",
                          "sample": f"{iters_m} synchronous-gbest passes over {N} particles, {cores} threads over the "
                                    f"particles (oracle orc_pso_step_mt, bit-identical to the 1-thread synchronous pass)"},
            "sample": f"{iters} sequential-gbest passes over {N} particles (n={n}, {obj.lower()}), "
                      f"gcc -O2 -ffp-contract=off restatement of particle_swarm.rs:399-427, host Philox draws, "
                      f"{os.cpu_count()} host cores visible, 1 used (the reference is single-threaded)"}


def also_measured(ctx, E):
    """Informational side measurements at N=1, outside the timed region of the headline (a few seconds): the other
    kernels of the path on their BASELINE configurations, device-resident, CUDA events on the library stream."""
    out = {}
    # C2 in the reference's sequential-gbest semantics (pso_seqw_kernel)
    n, N = 32, 1 << 20
    run = (E.ParticleSwarm.new(E.Objective.ROSENBROCK, [-5.0] * n, [10.0] * n).with_particle_count(N).with_tol(0.0)
           .with_seed(SEED).with_mode(E.PsoMode.SEQUENTIAL_EXACT).with_context(ctx)).start(np.zeros(n))
    run.step(3); run.sync(); run.step(40); run.sync()
    ms, launches = run.last_step_ms()
    out["c2_sequential_exact"] = {"value": N * 40 / (ms * 1e-3), "unit": "particle-evals/s", "ms_per_pass": ms / 40,
                                  "launches_per_pass": launches / 40, "passes": 40}
    run.close()
    # C3: CrossEntropy Rastrigin n=64, 2^22 samples per pass, 1 % elites, divisor k
    n, N = 64, 1 << 22
    ce = (E.CrossEntropy.new(E.Objective.RASTRIGIN).with_sample_size(N).with_importance_selection_size(N // 100)
          .with_iter_max(40).with_std_dev(np.full(n, 2.0)).with_mean_divisor(E.MeanDivisor.ELITE_COUNT).with_seed(SEED)
          .with_tol(0.0).with_context(ctx))
    crun = ce.start(np.full(n, 3.0))
    crun.step(3); crun.sync(); crun.step(20); crun.sync()
    ms, launches = crun.last_step_ms()
    out["c3_cross_entropy"] = {"value": N * 20 / (ms * 1e-3), "unit": "sample-evals/s", "ms_per_pass": ms / 20,
                               "launches_per_pass": launches / 20, "passes": 20,
                               "bound": "FP64 pipe 62 % busy, issue side (DESIGN.md section 4, profiles/r01i_ncu_ce_sample_c3.csv, profiles/r01i_issue_probe.log)"}
    crun.close()
    # MonteCarlo, the reference's 1-D test integrand at a GPU-sized point count
    count = 1 << 28
    mc = E.MonteCarlo.new(E.Objective.GAUSSIAN).with_sample_count(count).with_context(ctx)
    mc.integrate_with_seed(0.0, 1.0, 1)
    r = mc.integrate_with_seed(0.0, 1.0, 2)
    out["monte_carlo_1d"] = {"value": count / (mc.last_report.loop_ms * 1e-3), "unit": "samples/s", "samples": count,
                             "mean": r.mean, "exact": 0.746824132812427}
    return out


def run_reference(args, json_out):
    """--impl reference: the reference's CPU algorithm (oracle port; cargo/rustc are absent) on host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle as O
    obj, n, per_gpu, lo, hi, x0v, desc = WORKLOADS[args.workload]
    # bounded sample: size each step so that warmup + steps fit in ~90 s at ~1.2e6 evals/s
    budget_evals = 1.2e6 * 90.0
    N = int(min(per_gpu, max(1024, budget_evals / max(1, args.steps + args.warmup))))
    o = O.Pso(getattr(O, obj), [lo] * n, [hi] * n, particle_count=N, iter_max=10 ** 9, tol=0.0, mode=O.SEQUENTIAL,
              seed=SEED)
    o.init(np.full(n, x0v))
    o.time_steps(args.warmup)
    t = o.time_steps(args.steps)
    value = N * args.steps / t
    sample = (f"each step = one sequential-gbest pass over {N} particles of the workload (bounded sample of "
              f"{per_gpu} per GPU), 1 thread: the reference's solve() is single-threaded and its in-loop gbest update "
              f"is sequential, so one thread is all it can use")
    # beside it, what every host core gives when the pass is made synchronous and threaded over the particles
    # (NOT the reference's algorithm; SURVEY 8d's fair per-box figure)
    cores = os.cpu_count() or 1
    om = O.Pso(getattr(O, obj), [lo] * n, [hi] * n, particle_count=N, iter_max=10 ** 9, tol=0.0, mode=O.SYNCHRONOUS,
               seed=SEED)
    om.init(np.full(n, x0v))
    om.time_steps_mt(1, cores)
    steps_m = max(2, min(args.steps, 20))
    all_cores = {"value": N * steps_m / om.time_steps_mt(steps_m, cores), "unit": "particle-evals/s", "cores": cores,
                 "kind": "port", "sample": f"{steps_m} synchronous passes over the same {N} particles, {cores} threads"}
    line = {"impl": "reference", "metric": "particle-evals/s", "value": value, "unit": "particle-evals/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "n": n, "particles_per_gpu": per_gpu, "mode": "sequential (reference)"},
            "cpu_baseline": {"value": value, "unit": "particle-evals/s", "cores": 1, "kind": "port", "sample": sample,
                             "all_cores": all_cores},
            "e2e": {"value": value, "unit": "particle-evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "C restatement of eqsolver's algorithm (oracle/eqs_oracle.c), not the Rust crate: no cargo/rustc here"}
    json_out.write(json.dumps(line) + "\n")
    json_out.flush()
    return 0


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries underneath may print there too (NCCL writes its version line to
    fd 1 when NCCL_DEBUG is set in the environment): from here on fd 1 is stderr, and the returned file is the real
    stdout, used for the JSON line only."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the informational side measurements (N=1 only)")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for cpu_baseline")
    args = ap.parse_args()
    json_out = claim_stdout()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args, json_out)

    import torch
    import torch.distributed as dist
    import eqsolver_b200 as E
    from eqsolver_b200 import distributed as ED

    rank, world, local = ED.env_rank_world()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch N>1 with torch.distributed.run (one process per GPU)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group(backend="cpu:gloo,cuda:nccl", device_id=torch.device("cuda", local))
    ctx = ED.init_context(local)

    obj, n, per_gpu, lo, hi, x0v, desc = WORKLOADS[args.workload]
    N = per_gpu * world
    x0 = np.full(n, x0v)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def builder(iter_max):
        return (E.ParticleSwarm.new(getattr(E.Objective, obj), [lo] * n, [hi] * n).with_particle_count(N)
                .with_iter_max(iter_max).with_tol(0.0).with_seed(SEED).with_mode(E.PsoMode.SYNCHRONOUS)
                .with_context(ctx))

    # ---- device-resident throughput: exactly K passes between CUDA events on the library stream -------------
    run = builder(args.steps + args.warmup).start(x0)
    run.step(args.warmup)
    run.sync()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    t0 = time.monotonic()
    run.step(args.steps)
    run.sync()
    barrier()
    t1 = time.monotonic()
    ms, launches = run.last_step_ms()
    window = "timed region"
    # collective decision (the extra passes contain the per-pass exchange): every rank must take the same branch
    if max_over_ranks(1.0 if (t1 - t0) < 0.6 else 0.0) > 0.5:
        # the region is shorter than a few nvidia-smi periods: keep the same kernel running (untimed) for the sampler
        extra = int(max_over_ranks(max(1.0, 1.2 / max(ms * 1e-3 / args.steps, 1e-6))))
        run.step(extra)
        run.sync()
        t1 = time.monotonic()
        window = "timed region + ~1.2 s of the same step kernel (region too short to sample)"
    if world > 1:
        dist.barrier()
    clocks = None
    if sampler is not None:
        sampler.stop()
        clocks = sampler.summary(t0, t1, window)
    ms = max_over_ranks(ms)
    launches_total = int(sum_over_ranks(float(launches)))
    g, gcost, iters_run, stalled, _ = run.gbest()
    assert iters_run >= args.steps + args.warmup and not stalled, "passes were skipped: the timed region is invalid"
    run.close()
    value = N * args.steps / (ms * 1e-3)

    # ---- end to end through the public builder call, host buffers in and out ---------------------------------
    e2e_iters = args.steps
    ps = builder(e2e_iters)
    ps.solve(x0)  # warm-up call: first-use allocation paths
    barrier()
    te0 = time.perf_counter()
    out = ps.solve(x0)
    te1 = time.perf_counter()
    e2e_s = max_over_ranks(te1 - te0)
    rep = ps.last_report
    assert rep.iters_run == e2e_iters
    e2e_value = float(rep.evals) / e2e_s
    h2d = 3 * n * 8                    # x0, lower, upper
    d2h = n * 8 + 512                  # gbest position + the control block

    # ---- roofline of the step kernel --------------------------------------------------------------------------
    peak, peak_src = measured_peaks()
    bytes_per_launch = algorithmic_bytes_per_eval(n) * per_gpu
    launch_s = (ms * 1e-3) / args.steps          # one step kernel per pass per GPU (+ the exchange when N > 1)
    achieved = bytes_per_launch / launch_s / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "pso_step_traffic.json")) as f:
            traffic = json.load(f).get(args.workload, {}).get("dram_bytes_per_launch")
    except Exception:
        pass

    # the same access pattern with the arithmetic removed, measured now on this GPU (informational second ceiling)
    pattern = ctx.measure_pattern(n, per_gpu) if n % 8 == 0 else None

    line = {
        "metric": "particle-evals/s", "value": value, "unit": "particle-evals/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "n": n, "particles_per_gpu": per_gpu, "particles_total": N,
                   "mode": "synchronous gbest (one min-loc + exchange per pass)", "seed": SEED,
                   "l2": f"state {5 * n * 8 * per_gpu / 1e9:.2f} GB per GPU per pass exceeds the 126 MB L2 (no flush needed)",
                   "gbest_cost_after_run": gcost},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "particle-evals/s", "h2d_bytes_per_step": h2d / e2e_iters,
                "d2h_bytes_per_step": d2h / e2e_iters, "h2d_bytes_per_solve": h2d, "d2h_bytes_per_solve": d2h,
                "passes_per_solve": e2e_iters, "seconds_per_solve": e2e_s,
                "device_init_ms": rep.init_ms, "device_loop_ms": rep.loop_ms,
                "host_overhead_ms": 1e3 * e2e_s - rep.init_ms - rep.loop_ms,
                "call": "ParticleSwarm.new(obj, lower, upper).with_particle_count(N).with_iter_max(K).solve(x0)"},
        "gpu_launches": launches_total,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "kernel": "pso_step_kernel",
                     "algorithmic_bytes_per_eval": algorithmic_bytes_per_eval(n),
                     "frac_of_nominal_8000": achieved / 8000.0,
                     "pattern_ceiling": None if pattern is None else {
                         "value": pattern, "unit": "GB/s", "frac": achieved / pattern,
                         "what": "eqs_measure_pattern: this kernel's loads and stores (3 buffers read, 2 written in "
                                 "place, 256-bit, same layout and grid) with no arithmetic, best of 11 launches"}},
    }
    if world == 1 and not args.no_also:
        line["also"] = also_measured(ctx, E)
    # the ranks part before rank 0 formats its line: nothing below is collective, so a failure there cannot leave the
    # other ranks waiting in a barrier
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_port(args.workload, args.cpu_budget)
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
    return 0


if __name__ == "__main__":
    sys.exit(main())
