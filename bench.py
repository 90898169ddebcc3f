#!/usr/bin/env python3
"""bench.py -- particle-evals/s (f64) of the ParticleSwarm / CrossEntropy population step on B200, one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c1|c2|c3|c4|c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Default workload c2 (BASELINE.json configs[1], the configuration the metric is quoted on): ParticleSwarm, Rosenbrock
n=32, 2^20 particles PER GPU (weak scaling: contiguous index blocks per rank, one gbest exchange per pass), f64,
synchronous gbest, device Philox draws, tol = 0 so the stall test never ends the run early.  A "step" is one pass of the
outer loop (particle_swarm.rs:394-428 / cross_entropy.rs:256-302) over the whole population.

  value     evals/s with the population resident in HBM, timed on the device (CUDA events on the library's stream
            around exactly K passes, barrier + synchronize on both sides, max over ranks)
  e2e       evals/s through the public builder call `ParticleSwarm.new(..).solve(x0)` with HOST buffers: allocation,
            H2D of x0/bounds, init, K passes, D2H of the result all inside the timed region (host clock, max over
            ranks); only the K passes are counted as evaluations, the init pass is paid but not counted
  roofline  the dominant kernel against its measured ceiling: HBM copy bandwidth for the ParticleSwarm step (40n+16
            algorithmic bytes per particle-evaluation), the DFMA issue rate measured in the same run for the
            CrossEntropy sample kernel
  parity_check (N > 1)  the sharded result of a short run against the same run UNSHARDED on rank 0's GPU: gbest, ring
            and pbest costs bit-equal (ParticleSwarm); elite set equal and mu / sigma within 1e-12 (CrossEntropy).
            A mismatch makes every rank exit non-zero.
  also      the other BASELINE configurations at this N with their own roofline objects (per-GPU shapes, weak scaling)
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

SEED = 1729  # tests/integrators.rs:21
# where the few tensors of the cross-rank bookkeeping live (timings, flags, elite lists); tests/test_bench_dry_run.py runs
# the host logic over gloo on a box without a GPU and sets "cpu"
DEVICE = os.environ.get("EQS_BENCH_TENSOR_DEVICE", "cuda")

# ParticleSwarm workloads: (objective, n, particles per GPU, lower, upper, x0, description)
PSO_WORKLOADS = {
    "c2": ("ROSENBROCK", 32, 1 << 20, -5.0, 10.0, 0.0,
           "C2: ParticleSwarm Rosenbrock n=32, 2^20 particles per GPU, f64, synchronous gbest, Philox draws"),
    "c4": ("ACKLEY", 256, 1 << 21, -32.768, 32.768, 20.0,
           "C4: ParticleSwarm Ackley n=256, 2^21 particles per GPU (16M over 8), f64, synchronous gbest"),
    "c1": ("RASTRIGIN", 2, 1000, -5.12, 5.12, 5.0,
           "C1: ParticleSwarm Rastrigin n=2, 1000 particles (latency-bound; examples/global_optimisation.rs style)"),
}
# CrossEntropy workloads: (objective, n, samples per GPU, mu0, sigma0, elite fraction, description); mean divisor = k
# (the reference's literal 10 diverges for k != 10)
CE_WORKLOADS = {
    "c3": ("RASTRIGIN", 64, 1 << 22, 3.0, 2.0, 0.01,
           "C3: CrossEntropy Rastrigin n=64, 2^22 samples per pass per GPU, 1 % elites (radix select), f64"),
    "c5": ("ROSENBROCK", 1024, 1 << 23, 0.0, 2.0, 0.01,
           "C5: CrossEntropy Rosenbrock n=1024, 2^23 samples per pass per GPU (64M over 8), 1 % elites, f64"),
}
WORKLOADS = {**PSO_WORKLOADS, **CE_WORKLOADS}


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


def profile_json(name: str) -> dict:
    try:
        with open(os.path.join(ROOT, "profiles", name)) as f:
            return json.load(f)
    except Exception:
        return {}


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


# ---- CPU arm: the oracle port on host cores (the only place besides tests/ and smoke() that touches oracle/) ---------

def cpu_pso(workload: str, N: int, mode: int):
    from oracle import oracle as O
    obj, n, _, lo, hi, x0v, _ = PSO_WORKLOADS[workload]
    o = O.Pso(getattr(O, obj), [lo] * n, [hi] * n, particle_count=N, iter_max=10 ** 9, tol=0.0, mode=mode, seed=SEED)
    o.init(np.full(n, x0v))
    return o


def cpu_ce(workload: str, N: int):
    from oracle import oracle as O
    obj, n, _, mu0, s0, frac, _ = CE_WORKLOADS[workload]
    return O.Ce(getattr(O, obj), np.full(n, mu0), sample_size=N, elite_count=max(2, int(N * frac)), iter_max=10 ** 9,
                tol=0.0, std_dev=np.full(n, s0), mean_divisor=O.DIV_ELITE_COUNT, seed=SEED)


def cpu_baseline_port(workload: str, budget_s: float):
    """The oracle (C restatement, reference semantics, 1 thread) on a bounded sample of the workload."""
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    if workload in CE_WORKLOADS:
        obj, n, per_gpu, *_ = CE_WORKLOADS[workload]
        N = max(1000, min(per_gpu, int(2.0e8 / (40 * n))))  # a pass of about a second
        o = cpu_ce(workload, N)
        t1 = o.time_steps(1)
        iters = max(2, int(budget_s / max(t1, 1e-6)))
        t = o.time_steps(iters)
        return {"value": N * iters / t, "unit": "sample-evals/s", "cores": 1, "kind": "port",
                "sample": f"{iters} passes of {N} samples (n={n}, {obj.lower()}, 1 % elites by a full sort), gcc -O2 "
                          f"-ffp-contract=off restatement of cross_entropy.rs:256-302, host Philox + Box-Muller draws, "
                          f"{cores} host cores visible, 1 used (the reference is single-threaded)"}
    obj, n, per_gpu, *_ = PSO_WORKLOADS[workload]
    N = 1 << 16 if per_gpu >= (1 << 16) else per_gpu
    o = cpu_pso(workload, N, O.SEQUENTIAL)
    t1 = o.time_steps(1)                       # calibration pass (also warms the caches)
    iters = max(2, int(budget_s / max(t1, 1e-6)))
    t = o.time_steps(iters)
    # additionally the synchronous pass threaded over the particles on every host core (the fair "per box"
    # figure; the reference itself is single-threaded and sequential, so this one is NOT the reference's algorithm)
    om = cpu_pso(workload, N, O.SYNCHRONOUS)
    tm1 = om.time_steps_mt(1, cores)
    iters_m = max(2, int(0.5 * budget_s / max(tm1, 1e-6)))
    tm = om.time_steps_mt(iters_m, cores)
    return {"value": N * iters / t, "unit": "particle-evals/s", "cores": 1, "kind": "port",
            "all_cores": {"value": N * iters_m / tm, "unit": "particle-evals/s", "cores": cores, "kind": "port",
                          "sample": f"{iters_m} synchronous-gbest passes over {N} particles, {cores} threads over the "
                                    f"particles (oracle orc_pso_step_mt, bit-identical to the 1-thread synchronous pass)"},
            "sample": f"{iters} sequential-gbest passes over {N} particles (n={n}, {obj.lower()}), "
                      f"gcc -O2 -ffp-contract=off restatement of particle_swarm.rs:399-427, host Philox draws, "
                      f"{cores} host cores visible, 1 used (the reference is single-threaded)"}


def run_reference(args, json_out):
    """--impl reference: the reference's CPU algorithm (oracle port; cargo/rustc are absent) on host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    total = max(1, args.steps + args.warmup)
    if args.workload in CE_WORKLOADS:
        obj, n, per_gpu, _, _, _, desc = CE_WORKLOADS[args.workload]
        unit = "sample-evals/s"
        N = int(min(per_gpu, max(1000, 5.0e6 * 90.0 / (n * total))))   # ~5e6 coordinates/s, ~90 s in all
        o = cpu_ce(args.workload, N)
        o.time_steps(args.warmup)
        t = o.time_steps(args.steps)
        value = N * args.steps / t
        sample = (f"each step = one CrossEntropy pass of {N} samples (bounded sample of {per_gpu} per GPU), 1 thread: the "
                  f"reference's solve() is single-threaded")
        all_cores = None
        mode = "reference (full sort, sequential moments)"
    else:
        obj, n, per_gpu, _, _, _, desc = PSO_WORKLOADS[args.workload]
        unit = "particle-evals/s"
        # bounded sample: size each step so that warmup + steps fit in ~90 s at ~1.2e6 evals/s (n = 32)
        N = int(min(per_gpu, max(1024, 1.2e6 * 90.0 * 32 / (n * total))))
        o = cpu_pso(args.workload, N, O.SEQUENTIAL)
        o.time_steps(args.warmup)
        t = o.time_steps(args.steps)
        value = N * args.steps / t
        sample = (f"each step = one sequential-gbest pass over {N} particles of the workload (bounded sample of "
                  f"{per_gpu} per GPU), 1 thread: the reference's solve() is single-threaded and its in-loop gbest update "
                  f"is sequential, so one thread is all it can use")
        # beside it, what every host core gives when the pass is made synchronous and threaded over the particles
        # (NOT the reference's algorithm; SURVEY 8d's fair per-box figure)
        om = cpu_pso(args.workload, N, O.SYNCHRONOUS)
        om.time_steps_mt(1, cores)
        steps_m = max(2, min(args.steps, 20))
        all_cores = {"value": N * steps_m / om.time_steps_mt(steps_m, cores), "unit": unit, "cores": cores,
                     "kind": "port", "sample": f"{steps_m} synchronous passes over the same {N} particles, {cores} threads"}
        mode = "sequential (reference)"
    cpu = {"value": value, "unit": unit, "cores": 1, "kind": "port", "sample": sample}
    if all_cores:
        cpu["all_cores"] = all_cores
    line = {"impl": "reference", "metric": unit, "value": value, "unit": unit,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "n": n, "units_per_gpu": per_gpu, "mode": mode},
            "cpu_baseline": cpu,
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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


# ---- product arm ----------------------------------------------------------------------------------------------------

class Bench:
    """Rank-local state of one bench.py process: contexts, collectives over torch.distributed, builders."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import eqsolver_b200 as E
        from eqsolver_b200 import distributed as ED
        self.torch, self.dist, self.E = torch, dist, E
        self.rank, self.world, self.local = ED.env_rank_world()
        if self.world != args.gpus and self.world > 1:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}")
        if args.gpus > 1 and self.world == 1:
            raise SystemExit("launch N>1 with torch.distributed.run (one process per GPU)")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group(backend="cpu:gloo,cuda:nccl", device_id=torch.device("cuda", self.local))
        self.ctx = ED.init_context(self.local)   # joined to the library communicator when world > 1
        self._ctx1 = None

    @property
    def ctx1(self):
        """A second context on this GPU that is NOT part of the communicator: unsharded runs (parity_check, C1)."""
        if self.world == 1:
            return self.ctx
        if self._ctx1 is None:
            self._ctx1 = self.E.Context(self.local)
        return self._ctx1

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, v: float, op: str) -> float:
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device=DEVICE)
        self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op))
        return float(t.item())

    def gather_i64(self, a: np.ndarray, cap: int) -> np.ndarray:
        """All ranks' int64 arrays (each at most `cap` long) concatenated in rank order."""
        if self.world == 1:
            return a
        t = self.torch.full((cap + 1,), -1, dtype=self.torch.int64, device=DEVICE)
        t[0] = a.size
        if a.size:
            t[1:1 + a.size] = self.torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64)).to(DEVICE)
        parts = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(parts, t)
        out = []
        for p in parts:
            p = p.cpu().numpy()
            out.append(p[1:1 + int(p[0])])
        return np.concatenate(out)

    def agree(self, ok: bool) -> bool:
        """True only when every rank says so (a parity failure must take every rank out, not leave some in a barrier)."""
        return self.reduce(1.0 if ok else 0.0, "MIN") > 0.5

    # builders of the public API, on the sharded context unless told otherwise
    def pso(self, workload: str, N: int, iter_max: int, ctx=None, mode=None):
        E = self.E
        obj, n, _, lo, hi, _, _ = PSO_WORKLOADS[workload]
        return (E.ParticleSwarm.new(getattr(E.Objective, obj), [lo] * n, [hi] * n).with_particle_count(N)
                .with_iter_max(iter_max).with_tol(0.0).with_seed(SEED)
                .with_mode(E.PsoMode.SYNCHRONOUS if mode is None else mode).with_context(ctx or self.ctx))

    def ce(self, workload: str, N: int, passes: int, ctx=None):
        E = self.E
        obj, n, _, _, s0, frac, _ = CE_WORKLOADS[workload]
        return (E.CrossEntropy.new(getattr(E.Objective, obj)).with_sample_size(N)
                .with_importance_selection_size(max(2, int(N * frac))).with_iter_max(passes + 1)   # iter starts at 1, :253
                .with_std_dev(np.full(n, s0)).with_mean_divisor(E.MeanDivisor.ELITE_COUNT).with_seed(SEED).with_tol(0.0)
                .with_context(ctx or self.ctx))

    def start(self, workload: str, passes: int, ctx=None, world=None):
        world = self.world if world is None else world
        if workload in CE_WORKLOADS:
            _, n, per_gpu, mu0, *_ = CE_WORKLOADS[workload]
            return self.ce(workload, per_gpu * world, passes, ctx).start(np.full(n, mu0))
        _, n, per_gpu, _, _, x0v, _ = PSO_WORKLOADS[workload]
        return self.pso(workload, per_gpu * world, passes, ctx).start(np.full(n, x0v))

    def timed(self, run, warmup: int, steps: int, before_barrier=None):
        """W untimed passes, then exactly K passes between CUDA events on the library's stream.  The last warm-up pass is
        enqueued right behind the barrier with NO host work or host sync between it and the timed passes: at N > 1 it ends
        with a completed gbest exchange, so the start event of every rank is recorded within one exchange latency of
        the others' and no rank's first timed kernel waits in the mailbox for a peer that is still on the host."""
        run.step(max(warmup - 1, 0))
        run.sync()
        if before_barrier:
            before_barrier()
        self.barrier()
        t0 = time.monotonic()
        run.step(1)
        run.step(steps)
        run.sync()
        self.barrier()
        t1 = time.monotonic()
        ms, launches = run.last_step_ms()
        self.last_ms_by_rank = self.gather_f64(ms)      # every rank's own device time of the same K passes
        return self.reduce(ms, "MAX"), launches, t0, t1

    def gather_f64(self, v: float):
        if self.world == 1:
            return [v]
        t = self.torch.tensor([v], dtype=self.torch.float64, device=DEVICE)
        parts = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(parts, t)
        return [float(p.item()) for p in parts]


def pso_roofline(workload: str, n: int, per_gpu: int, ms_per_pass: float, pattern=None):
    peak, peak_src = measured_peaks()
    achieved = algorithmic_bytes_per_eval(n) * per_gpu / (ms_per_pass * 1e-3) / 1e9
    traffic = profile_json("pso_step_traffic.json").get(workload, {})
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic.get("dram_bytes_per_launch"),
            "traffic_source": ("static: " + traffic["source"]) if traffic.get("source") else None,
            "peak_source": peak_src, "kernel": "pso_step_kernel",
            "algorithmic_bytes_per_eval": algorithmic_bytes_per_eval(n), "evals_per_launch": per_gpu,
            "frac_of_nominal_8000": achieved / 8000.0,
            "pattern_ceiling": None if pattern is None else {
                "value": pattern, "unit": "GB/s", "frac": achieved / pattern,
                "what": "eqs_measure_pattern: this kernel's loads and stores (3 buffers read, 2 written in place, "
                        "256-bit, same layout and grid) with no arithmetic, best of 11 launches"}}


def ce_roofline(workload: str, per_gpu: int, ms_per_pass: float, fp64_peak_tflops: float):
    """The sample kernel against the FP64 pipe: FP64 instructions per sample (static SASS count of this build, cross-checked
    by the committed ncu capture) x samples per second, as DFMA-equivalent TFLOP/s (2 flop per lane), over the DFMA rate
    eqs_measure_fp64 reached in THIS run.  The whole pass is charged to the sample kernel (it is 95 % of it), so `frac`
    is a lower bound for the kernel alone.  `issue_frac` is the same rate against the ISSUE side of the same machine: a
    scheduler issues one warp instruction per cycle and the FP64 pipe takes one from it every second cycle, so a trip of F
    FP64 and O other instructions needs at least 2F + O scheduler cycles -- that, not the FP64 pipe alone, is what this
    kernel runs into, and why the work of both rounds went into the instruction count."""
    obj, n, *_ = CE_WORKLOADS[workload]
    prof = profile_json("ce_sample_fp64.json").get(workload, {})
    per_sample, other = prof.get("fp64_instr_per_sample"), prof.get("other_instr_per_sample")
    rate = per_gpu / (ms_per_pass * 1e-3)
    achieved = None if per_sample is None else rate * per_sample * 2.0 / 1e12
    frac = None if achieved is None or not fp64_peak_tflops else achieved / fp64_peak_tflops
    return {"bound": "fp64", "achieved": achieved, "peak": fp64_peak_tflops, "unit": "TFLOP/s", "frac": frac,
            "issue_frac": None if frac is None or other is None else frac * (2.0 * per_sample + other) / (2.0 * per_sample),
            "traffic": prof.get("dram_bytes_per_launch"),
            "traffic_source": ("static: " + prof["source"]) if prof.get("source") else None,
            "peak_source": "measured in this run (eqs_measure_fp64: independent DFMA chains on every SM)",
            "kernel": "ce_sample_kernel", "fp64_instr_per_sample": per_sample, "other_instr_per_sample": other,
            "samples_per_launch": per_gpu, "algorithmic_hbm_bytes_per_sample": 8,
            "note": "FP64-pipe / issue bound, not HBM: the sample matrix never leaves the registers"}


def bits_checksum(a: np.ndarray) -> np.ndarray:
    """Three 64-bit checksums of the BIT PATTERNS of a float64 array: wrapping sum, xor, position-weighted wrapping sum."""
    u = np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)
    w = np.arange(1, 2 * u.size, 2, dtype=np.uint64)
    with np.errstate(over="ignore"):
        c = np.array([np.sum(u, dtype=np.uint64), np.bitwise_xor.reduce(u) if u.size else 0, np.sum(u * w, dtype=np.uint64)],
                     dtype=np.uint64)
    return c.view(np.int64)


def parity_check(b: Bench, workload: str):
    """Sharded result == the same run unsharded on rank 0's GPU.  Collective: every rank calls it."""
    E = b.E
    t_start = time.monotonic()
    if workload in CE_WORKLOADS:
        obj, n, per_gpu, mu0, s0, frac, _ = CE_WORKLOADS[workload]
        per = min(per_gpu, max(1 << 14, (1 << 31) // n))   # keeps the regenerated elite matrix of the unsharded run small
        N, passes = per * b.world, 3
        k = max(2, int(N * frac))
        run = b.ce(workload, N, passes).start(np.full(n, mu0))
        run.step(passes)
        run.sync()
        mu, sg, it, _ = run.state()
        el = b.gather_i64(run.elites(), k)
        run.close()
        ok, detail = True, {}
        if b.rank == 0:
            ref = b.ce(workload, N, passes, ctx=b.ctx1).start(np.full(n, mu0))
            ref.step(passes)
            ref.sync()
            rmu, rsg, rit, _ = ref.state()
            rel = ref.elites()
            ref.close()
            scale = float(np.abs(rmu).max() + np.abs(rsg).max())
            dmu = float(np.abs(mu - rmu).max() / scale)
            dsg = float(np.abs(sg - rsg).max() / scale)
            same_set = el.size == rel.size and bool(np.array_equal(np.sort(el), np.sort(rel)))
            ok = same_set and dmu <= 1e-12 and dsg <= 1e-12 and it == rit == passes
            detail = {"elite_set_equal": same_set, "elites": int(rel.size), "mu_max_rel_diff": dmu,
                      "sigma_max_rel_diff": dsg, "tolerance": 1e-12}
        what = (f"CrossEntropy {obj.lower()} n={n}, {N} samples over {b.world} GPUs, {passes} passes: elite set of the last "
                f"pass and mu / sigma against the unsharded run on one GPU")
    else:
        obj, n, per_gpu, lo, hi, x0v, _ = PSO_WORKLOADS[workload]
        per = min(per_gpu, max(1024, int(40e9 / (b.world * 3 * 8 * n))))   # the unsharded population must fit one GPU
        N, passes = per * b.world, 8
        x0 = np.full(n, x0v)
        run = b.pso(workload, N, passes).start(x0)
        run.step(passes)
        run.sync()
        g, gc, it, _, imp = run.gbest()
        ring, ring_i = run.ring()
        sums = b.gather_i64(bits_checksum(run.pbest_costs()), 3)   # [world * 3], rank order
        run.close()
        ok, detail = True, {}
        if b.rank == 0:
            ref = b.pso(workload, N, passes, ctx=b.ctx1).start(x0)
            ref.step(passes)
            ref.sync()
            rg, rgc, rit, _, rimp = ref.gbest()
            rring, rring_i = ref.ring()
            pbc = ref.pbest_costs()
            ref.close()
            shards = [E.shard_range(N, b.world, r) for r in range(b.world)]
            want = np.concatenate([bits_checksum(pbc[f:f + c]) for f, c in shards])
            bits = lambda a: np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)
            checks = {"gbest_bit_equal": bool(np.array_equal(bits(g), bits(rg))) and gc == rgc,
                      "ring_bit_equal": bool(np.array_equal(bits(ring), bits(rring))) and ring_i == rring_i,
                      "improvements_equal": imp == rimp, "passes_equal": it == rit == passes,
                      "pbest_cost_checksums_equal": bool(np.array_equal(sums, want))}
            ok = all(checks.values())
            detail = dict(checks, gbest_cost=gc)
        what = (f"ParticleSwarm {obj.lower()} n={n}, {N} particles over {b.world} GPUs, {passes} synchronous passes on the "
                f"fused peer-memory exchange: gbest, stall ring, improvement count and per-rank pbest-cost checksums, bit "
                f"for bit against the unsharded run on one GPU")
    ok = b.agree(ok)
    return {"ok": ok, "what": what, "seconds": time.monotonic() - t_start, **detail}


def parity_check_sequential(b: Bench):
    """The reference's own semantics (sequential-exact gbest, particle_swarm.rs:399-427) sharded over the ranks -- tiles of
    256 particles dealt round-robin, device-driven windowed rounds with the first-improving index exchanged over the peer
    mailboxes -- against the same run on one GPU.  Collective: every rank calls it."""
    E = b.E
    t_start = time.monotonic()
    obj, n, per_gpu, lo, hi, x0v, _ = PSO_WORKLOADS["c2"]
    N, passes = (1 << 17) * b.world + 77, 6          # a ragged last tile
    x0 = np.full(n, x0v)
    run = b.pso("c2", N, passes, mode=E.PsoMode.SEQUENTIAL_EXACT).start(x0)
    run.step(passes)
    run.sync()
    g, gc, it, _, imp = run.gbest()
    ring, ring_i = run.ring()
    idx = run.local_indices()
    sums = b.gather_i64(bits_checksum(run.pbest_costs()), 3)
    counts = b.gather_i64(np.array([idx.size, int(idx.sum()) if idx.size else 0], dtype=np.int64), 2)
    run.close()
    ok, detail = True, {}
    if b.rank == 0:
        ref = b.pso("c2", N, passes, ctx=b.ctx1, mode=E.PsoMode.SEQUENTIAL_EXACT).start(x0)
        ref.step(passes)
        ref.sync()
        rg, rgc, rit, _, rimp = ref.gbest()
        rring, rring_i = ref.ring()
        pbc = ref.pbest_costs()
        ref.close()
        tile = np.arange(N) // 256
        owners = [np.nonzero(tile % b.world == r)[0] for r in range(b.world)]   # global tile T lives on rank T mod world
        want = np.concatenate([bits_checksum(pbc[o]) for o in owners])
        want_counts = np.concatenate([[o.size, int(o.sum())] for o in owners]).astype(np.int64)
        bits = lambda a: np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)
        checks = {"gbest_bit_equal": bool(np.array_equal(bits(g), bits(rg))) and gc == rgc,
                  "ring_bit_equal": bool(np.array_equal(bits(ring), bits(rring))) and ring_i == rring_i,
                  "improvements_equal": imp == rimp, "passes_equal": it == rit == passes,
                  "index_sets_equal": bool(np.array_equal(counts, want_counts)),
                  "pbest_cost_checksums_equal": bool(np.array_equal(sums, want))}
        ok = all(checks.values())
        detail = dict(checks, gbest_cost=gc, gbest_improvements=int(imp))
    ok = b.agree(ok)
    what = (f"ParticleSwarm {obj.lower()} n={n}, {N} particles over {b.world} GPUs, {passes} passes in the reference's "
            f"sequential-exact semantics (tile-cyclic distribution, device-driven rounds): gbest, stall ring, improvement "
            f"count and per-rank pbest-cost checksums, bit for bit against the same run on one GPU")
    return {"ok": ok, "what": what, "seconds": time.monotonic() - t_start, **detail}


def also_measured(b: Bench, headline: str, fp64_peak: float):
    """The other BASELINE configurations at this N (per-GPU shapes, weak scaling), outside the headline's timed region:
    device-resident, CUDA events on the library stream, max over ranks.  Collective: every rank calls it."""
    E, out = b.E, {}
    for wl, warm, passes in (("c4", 3, 12), ("c3", 3, 20), ("c5", 2, 4)):
        if wl == headline:
            continue
        run = b.start(wl, warm + passes)
        ms, launches, _, _ = b.timed(run, warm, passes)
        run.close()
        if wl in CE_WORKLOADS:
            obj, n, per_gpu, *_ = CE_WORKLOADS[wl]
            out[wl] = {"value": per_gpu * b.world * passes / (ms * 1e-3), "unit": "sample-evals/s",
                       "ms_per_pass": ms / passes, "launches_per_pass": launches / passes, "passes": passes,
                       "samples_per_gpu": per_gpu, "n": n, "roofline": ce_roofline(wl, per_gpu, ms / passes, fp64_peak)}
        else:
            obj, n, per_gpu, *_ = PSO_WORKLOADS[wl]
            out[wl] = {"value": per_gpu * b.world * passes / (ms * 1e-3), "unit": "particle-evals/s",
                       "ms_per_pass": ms / passes, "launches_per_pass": launches / passes, "passes": passes,
                       "particles_per_gpu": per_gpu, "n": n, "roofline": pso_roofline(wl, n, per_gpu, ms / passes)}
    if b.world > 1:
        # Every GPU of the job ALONE on the per-GPU shape, all at the same time, no exchange (each rank on its own
        # unsharded context): what separates "N x one GPU" from the sharded figures above is partly the exchange and partly
        # that the ranks wait for the slowest GPU of the box in every pass -- this shows the latter.
        for wl, warm, passes in (("c2", 5, 40), ("c3", 3, 20)):
            run = b.start(wl, warm + passes, ctx=b.ctx1, world=1)
            run.step(warm)
            run.sync()
            b.barrier()
            run.step(passes)
            run.sync()
            ms, _ = run.last_step_ms()
            run.close()
            per = b.gather_f64(ms / passes)
            per_gpu = (CE_WORKLOADS[wl] if wl in CE_WORKLOADS else PSO_WORKLOADS[wl])[2]
            out[wl + "_solo_by_rank"] = {"ms_per_pass": per, "slowest_over_fastest": max(per) / min(per),
                                         "n_times_slowest": per_gpu * b.world / (max(per) * 1e-3),
                                         "unit": "sample-evals/s" if wl in CE_WORKLOADS else "particle-evals/s",
                                         "what": "each rank runs the per-GPU shape unsharded on its own GPU, all ranks at "
                                                 "once; n_times_slowest = what a sharded pass could reach with a free exchange"}
        # C2 in the reference's sequential-gbest semantics over all ranks (weak scaling: 2^20 particles per GPU)
        n, per_gpu, passes = 32, 1 << 20, 30
        run = b.pso("c2", per_gpu * b.world, 10 ** 6, mode=E.PsoMode.SEQUENTIAL_EXACT).start(np.zeros(n))
        ms, launches, _, _ = b.timed(run, 3, passes)
        imp = run.gbest()[4]
        run.close()
        out["c2_sequential_exact_sharded"] = {
            "value": per_gpu * b.world * passes / (ms * 1e-3), "unit": "particle-evals/s", "ms_per_pass": ms / passes,
            "launches_per_pass": launches / passes, "passes": passes, "particles_per_gpu": per_gpu,
            "gbest_improvements": int(imp), "roofline": pso_roofline("c2", n, per_gpu, ms / passes),
            "how": "tiles of 256 particles dealt round-robin over the ranks; one persistent kernel per rank, the first "
                   "improving index of every round exchanged over the NVLink mailboxes"}
    if b.rank == 0:
        # single-GPU items, on rank 0's unsharded context
        if headline != "c1":   # C1: 1000 particles, 2-D -- one CTA runs the passes back to back (latency-bound, no roofline)
            obj, n, N, lo, hi, x0v, _ = PSO_WORKLOADS["c1"]
            run = b.pso("c1", N, 10 ** 6, ctx=b.ctx1).start(np.full(n, x0v))
            run.step(100); run.sync(); run.step(1000); run.sync()
            ms, launches = run.last_step_ms()
            out["c1"] = {"value": N * 1000 / (ms * 1e-3), "unit": "particle-evals/s", "us_per_pass": ms, "passes": 1000,
                         "launches": launches, "bound": "latency (state fits one SM; no roofline claim)"}
            run.close()
        # C2 in the reference's sequential-gbest semantics (pso_seqp_kernel)
        n, N = 32, 1 << 20
        run = b.pso("c2", N, 10 ** 6, ctx=b.ctx1, mode=E.PsoMode.SEQUENTIAL_EXACT).start(np.zeros(n))
        run.step(3); run.sync(); run.step(40); run.sync()
        ms, launches = run.last_step_ms()
        out["c2_sequential_exact"] = {"value": N * 40 / (ms * 1e-3), "unit": "particle-evals/s", "ms_per_pass": ms / 40,
                                      "launches_per_pass": launches / 40, "passes": 40,
                                      "roofline": pso_roofline("c2", n, N, ms / 40)}
        run.close()
        if b.world == 1:
            # MonteCarlo, the reference's 1-D test integrand at a GPU-sized point count
            count = 1 << 28
            mc = E.MonteCarlo.new(E.Objective.GAUSSIAN).with_sample_count(count).with_context(b.ctx1)
            mc.integrate_with_seed(0.0, 1.0, 1)
            r = mc.integrate_with_seed(0.0, 1.0, 2)
            out["monte_carlo_1d"] = {"value": count / (mc.last_report.loop_ms * 1e-3), "unit": "samples/s",
                                     "samples": count, "mean": r.mean, "exact": 0.746824132812427}
    if b.world > 1:
        b.dist.barrier()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the side measurements of the other configurations")
    ap.add_argument("--no-parity", action="store_true", help="skip the sharded-vs-unsharded check (N > 1)")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for cpu_baseline")
    args = ap.parse_args()
    json_out = claim_stdout()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args, json_out)

    b = Bench(args)
    rank, world, local = b.rank, b.world, b.local
    wl = args.workload
    is_ce = wl in CE_WORKLOADS
    if is_ce:
        obj, n, per_gpu, mu0, s0, frac, desc = CE_WORKLOADS[wl]
        unit, x0 = "sample-evals/s", np.full(n, mu0)
        if wl == "c5" and args.steps > 200:
            args.steps, args.warmup = 20, 3          # a pass is tens of milliseconds: the default K would take minutes
    else:
        obj, n, per_gpu, lo, hi, x0v, desc = PSO_WORKLOADS[wl]
        unit, x0 = "particle-evals/s", np.full(n, x0v)
        if wl == "c4" and args.steps > 400:
            args.steps, args.warmup = 100, 5
    N = per_gpu * world

    # ---- device-resident throughput: exactly K passes between CUDA events on the library stream -------------
    sampler = []
    run = b.start(wl, args.steps + args.warmup + 10 ** 6)
    # rank 0 forks nvidia-smi BEFORE the barrier: nothing but the enqueue of the passes happens behind it
    ms, launches, t0, t1 = b.timed(run, args.warmup, args.steps,
                                   before_barrier=(lambda: sampler.append(ClockSampler(local))) if rank == 0 else None)
    ms_by_rank = [m / args.steps for m in b.last_ms_by_rank]
    window = "timed region"
    # collective decision (the extra passes contain the per-pass exchange): every rank must take the same branch
    if b.reduce(1.0 if (t1 - t0) < 0.6 else 0.0, "MAX") > 0.5:
        # the region is shorter than a few nvidia-smi periods: keep the same kernel running (untimed) for the sampler
        extra = int(b.reduce(max(1.0, 1.2 / max(ms * 1e-3 / args.steps, 1e-6)), "MAX"))
        run.step(extra)
        run.sync()
        t1 = time.monotonic()
        window = "timed region + ~1.2 s of the same step kernel (region too short to sample)"
    if world > 1:
        b.dist.barrier()
    clocks = None
    if sampler:
        sampler[0].stop()
        clocks = sampler[0].summary(t0, t1, window)
    launches_total = int(b.reduce(float(launches), "SUM"))
    if is_ce:
        _, _, iters_run, active = run.state()
        assert iters_run >= args.steps + args.warmup and active, "passes were skipped: the timed region is invalid"
        end_state = {}
    else:
        g, gcost, iters_run, stalled, _ = run.gbest()
        assert iters_run >= args.steps + args.warmup and not stalled, "passes were skipped: the timed region is invalid"
        end_state = {"gbest_cost_after_run": gcost}
    run.close()
    value = N * args.steps / (ms * 1e-3)

    # ---- end to end through the public builder call, host buffers in and out ---------------------------------
    e2e_iters = args.steps
    solver = b.ce(wl, N, e2e_iters) if is_ce else b.pso(wl, N, e2e_iters)
    solver.solve(x0)  # warm-up call: first-use allocation paths
    b.barrier()
    te0 = time.perf_counter()
    solver.solve(x0)
    te1 = time.perf_counter()
    e2e_s = b.reduce(te1 - te0, "MAX")
    rep = solver.last_report
    assert rep.iters_run == e2e_iters
    e2e_value = float(N) * e2e_iters / e2e_s   # the K passes only: the init pass inside the call is paid, not counted
    h2d = (2 if is_ce else 3) * n * 8          # x0 + sigma0, or x0 + lower + upper
    d2h = n * 8 + 512                          # the result vector + the control block
    call = ("CrossEntropy.new(obj).with_sample_size(N).with_importance_selection_size(k).with_std_dev(s).with_iter_max(K+1).solve(x0)"
            if is_ce else "ParticleSwarm.new(obj, lower, upper).with_particle_count(N).with_iter_max(K).solve(x0)")

    # ---- rooflines ------------------------------------------------------------------------------------------------
    fp64_peak = b.ctx.measure_fp64()           # DFMA TFLOP/s, measured now on this GPU
    if is_ce:
        roofline = ce_roofline(wl, per_gpu, ms / args.steps, fp64_peak)
    else:
        # the same access pattern with the arithmetic removed, measured now on this GPU (informational second ceiling)
        pattern = b.ctx.measure_pattern(n, per_gpu) if n % 8 == 0 else None
        roofline = pso_roofline(wl, n, per_gpu, ms / args.steps, pattern)
        if pattern is not None:
            # the same pattern SUSTAINED for a second (csrc/api.cu: EQS_PATTERN_SECONDS): what the power-capped board
            # delivers for this traffic without any arithmetic -- the ceiling a long run of the step kernel is under
            os.environ["EQS_PATTERN_SECONDS"] = "1.0"
            try:
                sustained = b.ctx.measure_pattern(n, per_gpu)
            finally:
                os.environ.pop("EQS_PATTERN_SECONDS", None)
            roofline["pattern_ceiling"]["sustained"] = {"value": sustained, "unit": "GB/s", "seconds": 1.0,
                                                        "frac": roofline["achieved"] / sustained}

    line = {
        "metric": unit, "value": value, "unit": unit, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "ms_per_step_by_rank": ms_by_rank,   # each rank's own CUDA-event time; ms_per_step is their maximum
        "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict({"workload": desc, "n": n, "units_per_gpu": per_gpu, "units_total": N,
                        "particles_per_gpu": per_gpu, "particles_total": N,
                        "mode": ("radix-select elites, two-pass moments, mean divisor k" if is_ce else
                                 "synchronous gbest (one min-loc + exchange per pass)"), "seed": SEED,
                        "l2": (f"costs {8 * per_gpu / 1e6:.0f} MB per pass, the sample matrix is never materialised "
                               f"(compute-bound: nothing to flush)" if is_ce else
                               f"state {5 * n * 8 * per_gpu / 1e9:.2f} GB per GPU per pass exceeds the 126 MB L2 (no flush needed)"),
                        "warmup_note": "the last warm-up pass is enqueued behind the barrier, directly in front of the "
                                       "timed passes, so the ranks start aligned on the device"}, **end_state),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": h2d / e2e_iters,
                "d2h_bytes_per_step": d2h / e2e_iters, "h2d_bytes_per_solve": h2d, "d2h_bytes_per_solve": d2h,
                "passes_per_solve": e2e_iters, "seconds_per_solve": e2e_s,
                "evals_counted": "passes only (N x K); the init pass runs inside the timed call but is not counted",
                "value_with_init_evals": float(rep.evals) / e2e_s,
                "device_init_ms": rep.init_ms, "device_loop_ms": rep.loop_ms,
                "host_overhead_ms": 1e3 * e2e_s - rep.init_ms - rep.loop_ms, "call": call},
        "gpu_launches": launches_total,
        "roofline": roofline,
        "fp64_peak": {"value": fp64_peak, "unit": "TFLOP/s", "how": "eqs_measure_fp64, this run, this GPU"},
    }
    failed = False
    if world > 1 and not args.no_parity:
        pc = {"headline": parity_check(b, wl)}
        other = "c2" if is_ce else "c3"      # the other half of the metric: a CrossEntropy (or ParticleSwarm) pass
        pc[other] = parity_check(b, other)
        pc["c2_sequential_exact"] = parity_check_sequential(b)
        line["parity_check"] = all(v["ok"] for v in pc.values())
        line["parity_detail"] = pc
        failed = not line["parity_check"]
    if not args.no_also and not failed:
        line["also"] = also_measured(b, wl, fp64_peak)
    # the ranks part before rank 0 formats its line: nothing below is collective, so a failure there cannot leave the
    # other ranks waiting in a barrier
    if world > 1:
        b.dist.barrier()
        b.dist.destroy_process_group()
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_port(wl, args.cpu_budget)
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
    if failed:
        sys.stderr.write("parity_check FAILED: the sharded result differs from the unsharded run\n")
        return 3
    return 0


if __name__ == "__main__":
    sys.exit(main())
