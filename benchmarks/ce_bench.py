#!/usr/bin/env python3
"""CrossEntropy throughput (sample-evals/s) on one GPU: BASELINE config C3 by default.

    python benchmarks/ce_bench.py [--dim 64] [--samples 4194304] [--elite-frac 0.01] [--objective RASTRIGIN] [--iters 20]
    # C5 (CrossEntropy Rosenbrock n=1024, 64M samples per pass over 8 GPUs; --samples is PER GPU, weak scaling):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
        benchmarks/ce_bench.py --dim 1024 --samples 8388608 --objective ROSENBROCK --iters 5

Also prints the measured machine ceilings (DFMA issue rate, copy bandwidth) the CE roofline is quoted against.
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", dest="n", type=int, default=64)  # not --n: torchrun's own parser would claim it
    ap.add_argument("--samples", type=int, default=1 << 22)
    ap.add_argument("--elite-frac", type=float, default=0.01)
    ap.add_argument("--objective", default="RASTRIGIN")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-peaks", action="store_true")
    a = ap.parse_args()
    import eqsolver_b200 as E
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = 0
    if world > 1:  # one process per GPU; histogram + moment exchanges over NCCL (csrc/comm.cu)
        import torch
        import torch.distributed as dist
        from eqsolver_b200 import distributed as ED
        rank, world, local = ED.env_rank_world()
        torch.cuda.set_device(local)
        dist.init_process_group(backend="cpu:gloo,cuda:nccl", device_id=torch.device("cuda", local))
        ctx = ED.init_context(local)
    else:
        ctx = E.Context(0)
    a.samples *= world
    k = max(2, int(a.samples * a.elite_frac))
    ce = (E.CrossEntropy.new(getattr(E.Objective, a.objective)).with_sample_size(a.samples)
          .with_importance_selection_size(k).with_iter_max(a.iters + a.warmup + 1).with_std_dev(np.full(a.n, 2.0))
          .with_mean_divisor(E.MeanDivisor.ELITE_COUNT).with_seed(1729).with_tol(0.0).with_context(ctx))
    run = ce.start(np.full(a.n, 3.0))
    run.step(a.warmup)
    run.sync()
    if world > 1:
        dist.barrier()
    run.step(a.iters)
    run.sync()
    ms, launches = run.last_step_ms()
    if world > 1:  # device time, max over ranks
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    mu, sg, it, active = run.state()
    out = {"workload": f"CrossEntropy {a.objective.lower()} n={a.n}, {a.samples} samples/pass, k={k}", "n_gpus": world,
           "sample_evals_per_s": a.samples * a.iters / (ms * 1e-3), "ms_per_pass": ms / a.iters,
           "kernel_launches_per_pass": launches / a.iters, "passes": it, "sigma_inf": float(np.abs(sg).max())}
    if not a.no_peaks and world == 1:
        out["measured_fp64_tflops"] = ctx.measure_fp64()
        out["measured_copy_gbs"] = ctx.measure_copy(1 << 30)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        run.close()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
