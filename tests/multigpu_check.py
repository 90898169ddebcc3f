#!/usr/bin/env python3
"""Multi-GPU parity check, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tests/multigpu_check.py

Every rank holds a contiguous block of the population; per pass the ranks exchange one record over NCCL
(csrc/comm.cu).  The sharded GPU run must equal the UNSHARDED oracle: bit for bit for ParticleSwarm (min-loc with
lowest-index ties is order-free), to 1e-12 for CrossEntropy (moment sums are taken per rank, then in rank order).
Also run by tests/test_multigpu.py when the box has >= 2 GPUs.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import eqsolver_b200 as E
    from eqsolver_b200 import distributed as ED
    from oracle import oracle as O

    rank, world, local = ED.env_rank_world()
    torch.cuda.set_device(local)
    dist.init_process_group(backend="cpu:gloo,cuda:nccl", device_id=torch.device("cuda", local))
    ctx = ED.init_context(local)
    assert ctx.rank_world == (rank, world)

    def same_everywhere(a, what):
        a = np.ascontiguousarray(a, dtype=np.float64)
        t = torch.from_numpy(a.copy()).cuda()
        gathered = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        for r, g in enumerate(gathered):
            assert np.array_equal(g.cpu().numpy().view(np.uint64), a.view(np.uint64)), f"{what}: rank {r} differs"

    # ---- ParticleSwarm, synchronous: sharded GPU == unsharded oracle, bit for bit --------------------------------
    for (obj, n, N, iters) in [(E.Objective.ROSENBROCK, 9, 10007, 12), (E.Objective.SPHERE, 4, 3 * world + 1, 5),
                               (E.Objective.USER, 6, 4099, 6)]:
        lower, upper, x0 = [-5.0] * n, [10.0] * n, np.full(n, 3.0)
        ps = (E.ParticleSwarm.new(obj, lower, upper).with_particle_count(N).with_tol(0.0).with_seed(1729)
              .with_mode(E.PsoMode.SYNCHRONOUS).with_context(ctx))
        run = ps.start(x0)
        o = O.Pso(int(obj), lower, upper, particle_count=N, tol=0.0, mode=O.SYNCHRONOUS, seed=1729)
        o.init(x0)
        first, local_n = E.shard_range(N, world, rank)
        assert (run.first, run.local) == (first, local_n)
        for t in range(iters + 1):
            g, c, it, stalled, _ = run.gbest()
            assert np.array_equal(g.view(np.uint64), o.gbest().view(np.uint64)), f"PSO {obj.name} gbest pass {t}"
            assert c == o.gbest_cost() and it == o.iters()
            assert np.array_equal(run.ring()[0].view(np.uint64), o.ring()[0].view(np.uint64)) and run.ring()[1] == o.ring()[1]
            x, v, pb, pbc = run.state()
            ox, ov, opb, opbc = o.state()
            sl = slice(first, first + local_n)
            for a, b, nm in ((x, ox[sl], "x"), (v, ov[sl], "v"), (pb, opb[sl], "pbest"), (pbc, opbc[sl], "pbest cost")):
                assert np.array_equal(np.ascontiguousarray(a).view(np.uint64), np.ascontiguousarray(b).view(np.uint64)), \
                    f"PSO {obj.name} {nm} pass {t} rank {rank}"
            same_everywhere(g, "gbest")
            if t < iters:
                run.step(1)
                o.step()
        run.close()

    # ---- ParticleSwarm, sequential-exact across ranks: tile-cyclic distribution, device-driven windowed rounds with the
    # first-improving index exchanged over the peer mailboxes (pso_seqp_kernel); EQS_PSO_SEQ_HOST=1: the host-driven
    # whole-suffix rounds on contiguous blocks.  Every particle's state against the UNSHARDED oracle, bit for bit. ------
    def seq_case(obj, n, N, steps, seed, window=None, host=False, exact=True):
        for k, v in (("EQS_PSO_SEQ_WINDOW", window), ("EQS_PSO_SEQ_HOST", "1" if host else None)):
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = str(v)
        lower, upper, x0 = [-5.0] * n, [10.0] * n, np.full(n, 3.0)
        ps = (E.ParticleSwarm.new(obj, lower, upper).with_particle_count(N).with_tol(0.0).with_seed(seed)
              .with_mode(E.PsoMode.SEQUENTIAL_EXACT).with_context(ctx))
        run = ps.start(x0)
        o = O.Pso(int(obj), lower, upper, particle_count=N, tol=0.0, mode=O.SEQUENTIAL, seed=seed)
        o.init(x0)
        idx = run.local_indices()
        what = f"seq {obj.name} n={n} N={N} window={window} host={host}"
        assert idx.size == run.local and (np.diff(idx) > 0).all(), what
        t_idx = torch.full((N,), -1, dtype=torch.int64, device="cuda")
        counts = torch.zeros(N, dtype=torch.int64, device="cuda")
        counts[torch.from_numpy(idx).cuda()] += 1
        dist.all_reduce(counts)
        assert bool((counts == 1).all()), what + ": the ranks' index sets do not partition 0..N-1"
        done = 0
        for k in [0] + list(steps):
            if k:
                run.step(k)
                for _ in range(k):
                    o.step()
                done += k
            g, c, it, _, imp = run.gbest()
            eq = (lambda a, b: np.array_equal(np.ascontiguousarray(a).view(np.uint64), np.ascontiguousarray(b).view(np.uint64))) \
                if exact else (lambda a, b: np.allclose(a, b, rtol=1e-12, atol=1e-12))
            assert eq(g, o.gbest()) and eq(np.array([c]), np.array([o.gbest_cost()])), f"{what} gbest after {done} passes"
            assert it == o.iters(), what
            assert eq(run.ring()[0], o.ring()[0]) and run.ring()[1] == o.ring()[1], f"{what} ring after {done} passes"
            x, v, pb, pbc = run.state()
            ox, ov, opb, opbc = o.state()
            for a, bb, nm in ((x, ox[idx], "x"), (v, ov[idx], "v"), (pb, opb[idx], "pbest"), (pbc, opbc[idx], "pbest cost")):
                assert eq(a, bb), f"{what} {nm} after {done} passes, rank {rank}"
            same_everywhere(g, "gbest")
        run.close()
        os.environ.pop("EQS_PSO_SEQ_WINDOW", None)
        os.environ.pop("EQS_PSO_SEQ_HOST", None)

    seq_case(E.Objective.ROSENBROCK, 5, 2003, [1, 1, 2], seed=7)                 # 8 tiles, the last one ragged
    seq_case(E.Objective.ROSENBROCK, 5, 2003, [1, 3], seed=7, window=300)        # windows that straddle the ranks' tiles
    seq_case(E.Objective.SPHERE, 4, 256 * world + 1, [2, 2], seed=3, window=100) # one rank holds a 1-particle tile
    seq_case(E.Objective.SPHERE, 3, 200, [3], seed=5)                            # fewer tiles than ranks: empty ranks
    seq_case(E.Objective.USER, 6, 70001, [1, 2], seed=11)                        # default window policy, 274 tiles
    seq_case(E.Objective.ROSENBROCK, 9, 40000, [2], seed=13, window=5000)
    seq_case(E.Objective.RASTRIGIN, 8, 9000, [2], seed=17, exact=False)
    seq_case(E.Objective.ROSENBROCK, 5, 2003, [1, 1], seed=7, host=True)         # the fallback path still works

    # ---- whole solve through the builder, stall exit on every rank at the same pass -------------------------------
    ps = (E.ParticleSwarm.new(E.Objective.SPHERE, [-1.0] * 2, [1.0] * 2).with_particle_count(64 * world).with_tol(1e-3)
          .with_iter_max(400).with_seed(9).with_mode(E.PsoMode.SYNCHRONOUS).with_context(ctx))
    got = ps.solve([0.9, -0.9])
    oo = O.Pso(O.SPHERE, [-1.0] * 2, [1.0] * 2, particle_count=64 * world, tol=1e-3, iter_max=400, mode=O.SYNCHRONOUS, seed=9)
    want = oo.solve([0.9, -0.9])
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), (got, want)
    assert ps.last_report.iters_run == oo.iters(), (ps.last_report.iters_run, oo.iters())
    assert ps.last_report.world == world and ps.last_report.stalled == (1 if oo.iters() < 400 else 0)

    # ---- CrossEntropy: distributed radix select + elite moments ---------------------------------------------------
    n, N, k, iters = 7, 20011, 199, 5
    x0, s0 = np.full(n, 3.0), np.full(n, 2.0)
    ce = (E.CrossEntropy.new(E.Objective.ROSENBROCK).with_sample_size(N).with_importance_selection_size(k)
          .with_iter_max(iters + 1).with_std_dev(s0).with_mean_divisor(E.MeanDivisor.ELITE_COUNT).with_seed(11)
          .with_tol(0.0).with_context(ctx))
    crun = ce.start(x0)
    oc = O.Ce(O.ROSENBROCK, x0, sample_size=N, elite_count=k, iter_max=iters + 1, std_dev=s0,
              mean_divisor=O.DIV_ELITE_COUNT, seed=11, tol=0.0)
    for t in range(iters):
        crun.step(1)
        assert oc.step() == 0
        mu, sg, it, act = crun.state()
        omu, osg = oc.state()
        np.testing.assert_allclose(mu, omu, rtol=1e-12, atol=5e-12, err_msg=f"CE mu pass {t}")
        np.testing.assert_allclose(sg, osg, rtol=1e-12, atol=5e-12, err_msg=f"CE sigma pass {t}")
        same_everywhere(mu, "CE mu")
        el = crun.elites()
        t_el = torch.zeros(k, dtype=torch.int64, device="cuda") - 1
        t_el[:el.size] = torch.from_numpy(el).cuda()
        gathered = [torch.empty_like(t_el) for _ in range(world)]
        dist.all_gather(gathered, t_el)
        allel = np.concatenate([g.cpu().numpy() for g in gathered])
        allel = np.sort(allel[allel >= 0])
        assert allel.tolist() == np.sort(oc.elites()).tolist(), f"CE elite set pass {t}"
    crun.close()

    # ---- MonteCarlo: samples sharded over the ranks, moments merged in rank order == unsharded oracle -----------
    for (integrand, n, count) in [(E.Objective.GAUSSIAN, 3, 100003), (E.Objective.SINCOS, 2, 2 * world + 1)]:
        lower, upper = [0.0] * n, [1.5] * n
        r = E.MonteCarlo.new(integrand).with_sample_count(count).with_context(ctx).integrate_with_seed(lower, upper, 5)
        st, mean, var = O.mc_integrate(int(integrand), lower, upper, count, 5)
        assert st == 0 and abs(r.mean - mean) <= 1e-12 * abs(mean) and abs(r.variance - var) <= 1e-10 * abs(var)
        same_everywhere([r.mean, r.variance], "MC result")

    # ---- seeded random shapes over the real exchanges (every rank draws the same shapes) -------------------------
    rng = np.random.default_rng(31 + int(os.environ.get("EQS_SWEEP_SEED", "0")))
    exact = [E.Objective.SPHERE, E.Objective.ROSENBROCK, E.Objective.USER]
    for case in range(24):
        obj = exact[int(rng.integers(0, 3))]
        n = int(rng.integers(1, 30))
        N = int(rng.choice([1, world - 1, world, world + 1, 255, 2049])) if rng.random() < 0.4 else int(rng.integers(1, 9000))
        N = max(N, 1)
        mode = E.PsoMode.SYNCHRONOUS if rng.random() < 0.7 else E.PsoMode.SEQUENTIAL_EXACT
        seed = int(rng.integers(0, 2 ** 40))
        x0 = rng.uniform(-4, 8, n)
        what = f"random case {case}: {obj.name} n={n} N={N} {mode.name} seed={seed} rank {rank}"
        run = (E.ParticleSwarm.new(obj, [-5.0] * n, [10.0] * n).with_particle_count(N).with_tol(0.0).with_seed(seed)
               .with_mode(mode).with_context(ctx)).start(x0)
        o = O.Pso(int(obj), [-5.0] * n, [10.0] * n, particle_count=N, tol=0.0, mode=int(mode), seed=seed)
        o.init(x0)
        run.step(3)
        for _ in range(3):
            o.step()
        g, c, it, _, _ = run.gbest()
        assert np.array_equal(g.view(np.uint64), o.gbest().view(np.uint64)) and c == o.gbest_cost() and it == 3, what
        assert np.array_equal(run.ring()[0].view(np.uint64), o.ring()[0].view(np.uint64)) and run.ring()[1] == o.ring()[1], what
        idx = run.local_indices()  # contiguous blocks (synchronous) or tiles dealt round-robin (sequential-exact)
        if mode == E.PsoMode.SYNCHRONOUS:
            first, local_n = E.shard_range(N, world, rank)
            assert idx.tolist() == list(range(first, first + local_n)), what
        if idx.size > 0:
            for a, b in zip(run.state(), o.state()):
                b = np.ascontiguousarray(b[idx])
                assert np.array_equal(np.ascontiguousarray(a).view(np.uint64), b.view(np.uint64)), what
        run.close()
    for case in range(8):
        obj = exact[int(rng.integers(0, 2))]
        n = int(rng.integers(1, 20))
        N = int(rng.integers(max(2, world), 20000))
        k = int(rng.integers(2, min(N, 400) + 1))
        seed = int(rng.integers(0, 2 ** 40))
        x0c, s0c = rng.uniform(-2, 2, n), rng.uniform(0.5, 2.0, n)
        what = f"random CE case {case}: {obj.name} n={n} N={N} k={k} seed={seed} rank {rank}"
        crun = (E.CrossEntropy.new(obj).with_sample_size(N).with_importance_selection_size(k).with_iter_max(4)
                .with_std_dev(s0c).with_mean_divisor(E.MeanDivisor.ELITE_COUNT).with_seed(seed).with_tol(0.0)
                .with_context(ctx)).start(x0c)
        oc = O.Ce(int(obj), x0c, sample_size=N, elite_count=k, iter_max=4, std_dev=s0c, mean_divisor=O.DIV_ELITE_COUNT,
                  seed=seed, tol=0.0)
        for t in range(2):
            crun.step(1)
            assert oc.step() == 0
            mu, sg, _, _ = crun.state()
            omu, osg = oc.state()
            scale = float(np.abs(omu).max() + np.abs(osg).max() + 1.0)
            np.testing.assert_allclose(mu, omu, rtol=1e-11, atol=1e-12 * scale, err_msg=what)
            np.testing.assert_allclose(sg, osg, rtol=1e-11, atol=1e-12 * scale, err_msg=what)
        crun.close()

    # ---- a peer that never shows up: the rank that waits in its kernel gives up after EQS_P2P_TIMEOUT_S and the call fails
    # with ExternalError; nothing hangs.  Last thing on this context: its exchange counters are out of step afterwards. ----
    if os.environ.get("EQS_P2P", "1") != "0":
        dist.barrier()
        os.environ["EQS_P2P_TIMEOUT_S"] = "2"
        n = 4
        run = (E.ParticleSwarm.new(E.Objective.SPHERE, [-1.0] * n, [1.0] * n).with_particle_count(4096 * world).with_tol(0.0)
               .with_seed(3).with_mode(E.PsoMode.SYNCHRONOUS).with_context(ctx)).start(np.full(n, 0.5))
        run.step(1)
        run.sync()
        dist.barrier()
        if rank == 0:   # only rank 0 enqueues the next pass
            import time
            t0 = time.monotonic()
            try:
                run.step(1)
                run.gbest()
                raise AssertionError("a pass without its peers must fail")
            except E.ExternalError as e:
                assert "peer" in str(e), str(e)
            waited = time.monotonic() - t0
            assert 1.5 < waited < 15.0, waited
        dist.barrier()
        os.environ.pop("EQS_P2P_TIMEOUT_S", None)

    dist.barrier()
    if rank == 0:
        print(f"multigpu_check OK on {world} GPUs")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
