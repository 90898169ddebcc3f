"""CPU, world_size 2 over gloo: the N>1 host logic -- rendezvous / unique-id shipping (eqsolver_b200/distributed.py),
shard ranges, and the exchange PROTOCOL of DESIGN.md §5 run on the sharded oracle with the records moved by
torch.distributed.  The CUDA kernels implement the same protocol (tests/test_pso_gpu.py virtual ranks,
tests/multigpu_check.py over NCCL)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT

WORLD = 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, results):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist
    import eqsolver_b200 as E
    from eqsolver_b200 import distributed as ED
    from oracle import oracle as O

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert ED.env_rank_world() == (rank, world, rank)
        # 1. the 128-byte communicator id travels from rank 0 to everyone
        uid = ED.exchange_unique_id(rank, make_id=lambda: bytes(range(128)))
        assert uid == bytes(range(128))

        def allgather(vec):
            t = torch.from_numpy(np.ascontiguousarray(vec, dtype=np.float64).copy())
            out = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(out, t)
            return np.stack([o.numpy() for o in out])

        # 2. ParticleSwarm (synchronous) sharded over the ranks == unsharded, bit for bit, ring included
        n, N, iters = 5, 203, 7
        lo, hi, x0 = [-5.0] * n, [10.0] * n, [3.0] * n
        first, local = E.shard_range(N, world, rank)
        ref = O.Pso(O.ROSENBROCK, lo, hi, particle_count=N, mode=O.SYNCHRONOUS, seed=3, tol=0.0)
        ref.init(x0)
        sh = O.Pso(O.ROSENBROCK, lo, hi, particle_count=N, mode=O.SYNCHRONOUS, seed=3, tol=0.0, first_index=first,
                   local_count=local)
        recs = allgather(sh.init_local(x0))                       # exchange 1: arg-min records
        f0 = O.objective(O.ROSENBROCK, x0)
        seed_cost = min([f0] + [recs[q, 0] for q in range(rank) if recs[q, 1] >= 0])
        pushes, tail = sh.init_ring_local(seed_cost)
        ring_recs = allgather(np.concatenate([[pushes], tail]))   # exchange 2: ordered ring pushes
        sh.init_apply(recs, ring_recs[:, 0].astype(np.int64), ring_recs[:, 1:])
        for t in range(iters + 1):
            assert np.array_equal(sh.gbest().view(np.uint64), ref.gbest().view(np.uint64)), f"gbest pass {t}"
            assert np.array_equal(sh.ring()[0].view(np.uint64), ref.ring()[0].view(np.uint64)) and sh.ring()[1] == ref.ring()[1]
            assert np.array_equal(sh.state()[0].view(np.uint64), ref.state()[0][first:first + local].view(np.uint64))
            if t < iters:
                stalled, rec = sh.step_local()
                sh.step_apply(allgather(rec))
                ref.step()

        # 3. CrossEntropy: global k-th cost + lowest-index ties, moment sums per rank then in rank order
        n, N, k = 4, 301, 23
        x0c, s0 = np.full(n, 3.0), np.full(n, 2.0)
        first, local = E.shard_range(N, world, rank)
        refc = O.Ce(O.ROSENBROCK, x0c, sample_size=N, elite_count=k, iter_max=6, std_dev=s0,
                    mean_divisor=O.DIV_ELITE_COUNT, seed=5, tol=0.0)
        shc = O.Ce(O.ROSENBROCK, x0c, sample_size=N, elite_count=k, iter_max=6, std_dev=s0,
                   mean_divisor=O.DIV_ELITE_COUNT, seed=5, tol=0.0, first_index=first, local_count=local)
        maxloc = max(E.shard_range(N, world, r)[1] for r in range(world))
        for t in range(4):
            refc.step()
            shc.sample_local()
            c = np.full(maxloc, np.inf); c[:local] = shc.costs()
            allc = allgather(c)
            flat = np.concatenate([allc[r, :E.shard_range(N, world, r)[1]] for r in range(world)])
            T = np.sort(flat)[k - 1]
            quota = k - int((flat < T).sum())
            ties_before = sum(int((allc[r, :E.shard_range(N, world, r)[1]] == T).sum()) for r in range(rank))
            mine = int((shc.costs() == T).sum())
            take = min(max(quota - ties_before, 0), mine)
            sums = allgather(shc.elite_sum_local(T, take))
            mu_new = np.zeros(n)
            for r in range(world):
                mu_new = mu_new + sums[r]
            mu_new = mu_new / k
            sq = allgather(shc.elite_sqsum_local(mu_new))
            tot = np.zeros(n)
            for r in range(world):
                tot = tot + sq[r]
            shc.apply(mu_new, np.sqrt(tot / (k - 1)))
            mu, sg = shc.state(); rmu, rsg = refc.state()
            np.testing.assert_allclose(mu, rmu, rtol=1e-12, atol=1e-12)
            np.testing.assert_allclose(sg, rsg, rtol=1e-12, atol=1e-12)

        # 4. MonteCarlo: points sharded over the ranks, Welford moments merged pairwise in rank order (csrc/mc.cu)
        lower, upper, count = [0.0, -1.0, 0.5], [1.0, 1.0, 2.0], 1001
        first, local = E.shard_range(count, world, rank)
        parts = allgather(O.mc_moments(O.GAUSSIAN, lower, upper, first, local, seed=9))
        cnt, mean, m2 = 0.0, 0.0, 0.0
        for r in range(world):
            nb, mb, sb = parts[r]
            if nb == 0:
                continue
            if cnt == 0:
                cnt, mean, m2 = nb, mb, sb
                continue
            tot, delta = cnt + nb, mb - mean
            mean, m2, cnt = mean + delta * (nb / tot), m2 + sb + delta * delta * (cnt * nb / tot), tot
        vol = np.prod(np.array(upper) - np.array(lower))
        st, rmean, rvar = O.mc_integrate(O.GAUSSIAN, lower, upper, count, seed=9)
        assert st == 0 and cnt == count
        assert abs(mean * vol - rmean) <= 1e-13 * abs(rmean) and abs(m2 / (count - 1) * vol * vol - rvar) <= 1e-12 * rvar
        # 5. ParticleSwarm in the reference's SEQUENTIAL semantics over the ranks, the protocol of csrc/pso.cu: pso_seqp_kernel /
        #    pso_seq_exchange: tiles of consecutive particles dealt round-robin; the ordered ring pushes of the
        #    init loop rebuilt from the tile minima of ALL ranks and merged by position; passes as windowed rounds whose only
        #    exchange is (first improving global index, its cost, its position).  Modelled in numpy on the oracle's draws and
        #    objective (the arithmetic below is the oracle's, bit for bit), against the UNSHARDED oracle.
        TILE, n, N, W, seed = 16, 4, 1003, 150, 21           # the kernels use tiles of 256; the protocol does not care
        lo, hi, x0 = np.full(n, -5.0), np.full(n, 10.0), np.full(n, 3.0)
        ref = O.Pso(O.ROSENBROCK, lo, hi, particle_count=N, mode=O.SEQUENTIAL, seed=seed, tol=0.0)
        ref.init(x0)
        own = np.array([g for g in range(N) if (g // TILE) % world == rank])
        f = lambda xx: O.objective(O.ROSENBROCK, xx)
        X, V = {}, {}
        for g in own:
            pu, vu = O.pso_init_draws(seed, int(g), n)
            dist_ = np.abs(hi - lo)
            X[g], V[g] = pu * (hi - lo) + lo, vu * (dist_ + dist_) + (-dist_)
        PB = {g: X[g].copy() for g in own}
        PBC = {g: f(X[g]) for g in own}
        # init, exchange 1: every rank's best (cost, index, position); lowest cost, then lowest index
        best = min(own, key=lambda g: (PBC[g], g)) if own.size else -1
        rec = allgather(np.concatenate([[PBC[best] if best >= 0 else np.inf, best], X[best] if best >= 0 else np.zeros(n)]))
        # init, exchange 2: the minima of ALL tiles, walked in global order
        ntiles = (N + TILE - 1) // TILE
        per = (ntiles + world - 1) // world
        gm = np.full(per, np.inf)
        for g in own:
            gm[(g // TILE) // world] = min(gm[(g // TILE) // world], PBC[g])
        gm_all = allgather(gm)
        f0 = f(x0)
        running, pushes = f0, []                              # (position, pushed value) of this rank's record lows
        for T in range(ntiles):
            tmin = gm_all[T % world, T // world]
            if tmin < running and T % world == rank:          # a tile with a record low that this rank owns
                e = running
                for g in range(T * TILE, min((T + 1) * TILE, N)):
                    if PBC[g] < e:
                        pushes.append((g, e))
                        e = PBC[g]
            running = min(running, tmin)
        kept = pushes[-50:]
        msg = np.full(1 + 100, -1.0)
        msg[0] = len(pushes)
        for j, (g, e) in enumerate(kept):
            msg[1 + j], msg[51 + j] = g, e
        allp = allgather(msg)
        total = int(allp[:, 0].sum())
        merged = sorted((int(allp[r, 1 + j]), allp[r, 51 + j]) for r in range(world) for j in range(50) if allp[r, 1 + j] >= 0)
        ring = np.full(50, np.inf)
        for back, (g, e) in enumerate(reversed(merged[-50:])):  # push number total-1-back goes to slot (number % 50)
            ring[(total - 1 - back) % 50] = e
        ring_i = total % 50
        win = min(range(world), key=lambda r: (rec[r, 0], rec[r, 1] if rec[r, 1] >= 0 else np.inf))
        G, Gc = (rec[win, 2:].copy(), rec[win, 0]) if rec[win, 1] >= 0 and rec[win, 0] < f0 else (x0.copy(), f0)
        assert np.array_equal(ring.view(np.uint64), ref.ring()[0].view(np.uint64)) and ring_i == ref.ring()[1], "seq init ring"
        assert np.array_equal(G.view(np.uint64), ref.gbest().view(np.uint64)) and Gc == ref.gbest_cost(), "seq init gbest"
        wins = [0] * world
        for t in range(4):
            start, rounds = 0, 0
            while start < N:
                wend = min(start + W, N)
                spec, first = {}, -1
                for g in own[(own >= start) & (own < wend)]:
                    rp, rg = O.pso_step_draws(seed, t, int(g), n)
                    v2 = 0.5 * V[g] + 1.0 * rp * (PB[g] - X[g]) + 1.0 * rg * (G - X[g])
                    x2 = X[g] + v2
                    spec[g] = (x2, v2, f(x2))
                    if first < 0 and spec[g][2] < PBC[g] and spec[g][2] < Gc:
                        first = g                             # own is ascending: the first hit is the lowest index
                got = allgather(np.concatenate([[first, spec[first][2] if first >= 0 else np.inf],
                                                spec[first][0] if first >= 0 else np.zeros(n)]))
                hits = [r for r in range(world) if got[r, 0] >= 0]
                wr = min(hits, key=lambda r: got[r, 0]) if hits else -1
                end = int(got[wr, 0]) if wr >= 0 else wend - 1
                for g in own[(own >= start) & (own <= end)]:  # final: commit, pbest test (:413-419)
                    X[g], V[g], c = spec[g]
                    if c < PBC[g]:
                        PB[g], PBC[g] = X[g].copy(), c
                if wr >= 0:                                   # :420-425
                    wins[wr] += 1
                    ring[ring_i] = Gc
                    ring_i = (ring_i + 1) % 50
                    G, Gc = got[wr, 2:].copy(), got[wr, 1]
                start, rounds = end + 1, rounds + 1
            ref.step()
            rx, rv, rpb, rpbc = ref.state()
            for g in own:
                assert np.array_equal(X[g].view(np.uint64), rx[g].view(np.uint64)), f"seq x {g} pass {t}"
                assert np.array_equal(V[g].view(np.uint64), rv[g].view(np.uint64)), f"seq v {g} pass {t}"
                assert np.array_equal(PB[g].view(np.uint64), rpb[g].view(np.uint64)) and PBC[g] == rpbc[g], f"seq pbest {g} pass {t}"
            assert np.array_equal(G.view(np.uint64), ref.gbest().view(np.uint64)) and Gc == ref.gbest_cost(), f"seq gbest pass {t}"
            assert np.array_equal(ring.view(np.uint64), ref.ring()[0].view(np.uint64)) and ring_i == ref.ring()[1], f"seq ring {t}"
            assert rounds >= (N + W - 1) // W
        assert all(w > 0 for w in wins), wins                 # gbest moved through particles of every rank
        results[rank] = "ok"
    except Exception as e:  # surface the failure to the parent
        import traceback
        results[rank] = "FAIL: " + "".join(traceback.format_exception(e))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    results = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(WORLD, port, results), nprocs=WORLD, join=True)
    assert dict(results) == {0: "ok", 1: "ok"}, "\n".join(str(v) for v in results.values())
