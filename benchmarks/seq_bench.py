#!/usr/bin/env python3
"""ParticleSwarm in SEQUENTIAL_EXACT mode (the reference's in-loop gbest semantics, particle_swarm.rs:399-427) next to
the synchronous mode on one GPU: particle-evals/s and the repair overhead, C2 shape by default.

    python benchmarks/seq_bench.py [--dim 32] [--particles 1048576] [--objective ROSENBROCK] [--iters 40]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", dest="n", type=int, default=32)
    ap.add_argument("--particles", type=int, default=1 << 20)
    ap.add_argument("--objective", default="ROSENBROCK")
    ap.add_argument("--iters", type=int, default=40)
    ap.add_argument("--lower", type=float, default=-5.0)
    ap.add_argument("--upper", type=float, default=10.0)
    a = ap.parse_args()
    import eqsolver_b200 as E
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), 0
    if world > 1:  # one process per GPU, weak scaling: --particles is the PER-GPU count
        import torch
        import torch.distributed as dist
        from eqsolver_b200 import distributed as ED
        rank, world, local = ED.env_rank_world()
        torch.cuda.set_device(local)
        dist.init_process_group(backend="cpu:gloo,cuda:nccl", device_id=torch.device("cuda", local))
        ctx = ED.init_context(local)
        a.particles *= world
    else:
        ctx = E.Context(0)
    out = {"workload": f"ParticleSwarm {a.objective.lower()} n={a.n}, {a.particles} particles, {a.iters} passes",
           "n_gpus": world}
    for mode in (E.PsoMode.SYNCHRONOUS, E.PsoMode.SEQUENTIAL_EXACT):
        ps = (E.ParticleSwarm.new(getattr(E.Objective, a.objective), [a.lower] * a.n, [a.upper] * a.n)
              .with_particle_count(a.particles).with_tol(0.0).with_seed(1729).with_mode(mode).with_context(ctx))
        run = ps.start(np.zeros(a.n))
        run.step(3)
        run.sync()
        imp0 = run.gbest()[4]
        if world > 1:
            dist.barrier()
        run.step(a.iters)
        run.sync()
        ms, launches = run.last_step_ms()
        if world > 1:  # device time, max over ranks
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        g, cost, it, stalled, imp = run.gbest()
        out[mode.name.lower()] = {"evals_per_s": a.particles * a.iters / (ms * 1e-3), "ms_per_pass": ms / a.iters,
                                  "launches_per_pass": launches / a.iters, "gbest_cost": cost,
                                  "gbest_improvements_per_pass": (imp - imp0) / a.iters}
        run.close()
    out["sequential_over_synchronous_time"] = out["sequential_exact"]["ms_per_pass"] / out["synchronous"]["ms_per_pass"]
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
