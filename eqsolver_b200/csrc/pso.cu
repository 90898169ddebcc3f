// pso.cu -- ParticleSwarm population step on sm_100a (f64, no tensor cores: nothing here is a contraction).
//
// Replaces the body of ParticleSwarm::solve, reference src/solvers/global_optimisers/particle_swarm.rs:356-431:
//   pso_x0_kernel        :357-359   gbest = x0, ring = +inf
//   pso_init_kernel      :361-389   sample U[lo,hi] / U[-|hi-lo|,|hi-lo|], evaluate, pbest = x   (K1)
//   pso_ring_scan_kernel :376-380   the ordered gbest record-lows of the init loop -> ring pushes
//   pso_step_kernel      :399-419   fused velocity/position update + objective + pbest compare     (K2)
//                        + block/grid min-loc reduction (K3) + (single rank) the gbest update :420-425
//   pso_spec/commit      :399-427   sequential-exact mode: speculate-and-repair
//   pso_apply_*          :420-425, :395-397, :433-441  gbest update, ring push, stall test
//
// DATA LAYOUT.  The reference keeps an array of structs `Vec<Particle>` (:14-22).  Here every
// per-particle vector field is one buffer [ceil(n/4)][ld][4] f64: four consecutive dimensions of ONE particle
// form a 32-byte DRAM sector, particles are contiguous within a 4-dimension group.  A warp reads 1 KB of
// consecutive bytes with one 256-bit access per lane (LDG.E.ENL2.256), and every per-particle decision (which
// buffer to write, whether to read pbest at all) is sector-granular, so it never causes partial-sector traffic.
//
// ROLE-SWAPPING POSITION BUFFERS (synchronous mode).  Position and pbest position live in two buffers A (= P.x) and
// B (= P.pb) with one flag byte per particle: bit 1 = "x lives in B", bit 0 = "stale": the last pass improved
// pbest, so pbest position == position and only one buffer holds live data.  A stale particle writes its new
// position into the OTHER buffer and flips the role bit: the old buffer now IS the pbest position.  `best_position
// = position.clone()` (:417-419) therefore costs no memory traffic, and the pbest sector of a stale particle is not
// even read: <= 40n bytes per particle-evaluation whatever the improvement rate (ncu showed 70 % of the particles of
// config 2 improving per pass, +14 % DRAM bytes in a copy-on-improve scheme; profiles/).
#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>

#ifdef EQS_PSO_COS_TABLE // tuning sweeps: where the step kernels' cosine takes its coefficients from (devmath.cuh)
#define EQS_COS_TABLE EQS_PSO_COS_TABLE
#endif
#include "objectives.cuh"
#include "pso.cuh"

namespace eqs {

namespace {

// tuning knobs of the step kernel (measured on B200, profiles/r01_sweep.md): groups per load batch, min CTAs per SM
#ifndef EQS_PSO_UQ
#define EQS_PSO_UQ 2
#endif
// cache-policy qualifiers of the 256-bit population accesses (tuning sweeps; empty = default policy)
#ifndef EQS_PSO_LDHINT
#define EQS_PSO_LDHINT ""
#endif
#ifndef EQS_PSO_STHINT
#define EQS_PSO_STHINT ""
#endif
#ifndef EQS_PSO_MINB
#define EQS_PSO_MINB 2
#endif
// groups per load batch in the persistent sequential-exact kernel: a round there is latency-bound (every thread moves at
// most one particle per round: n / (4 UQ) dependent load batches), not bandwidth-bound
#ifndef EQS_PSO_SEQ_UQ
#define EQS_PSO_SEQ_UQ EQS_PSO_UQ
#endif
// 1: the objective takes the 4 * UQ coordinates of a load batch at once, after the batch's stores, instead of 4 at a time
// between them (measured on one box, profiles/r02_sweep.md section 3: C2 0.866 -> 0.892, C4 0.875 -> 0.881 of measured HBM)
#ifndef EQS_PSO_FEED_BATCH
#define EQS_PSO_FEED_BATCH 1
#endif
// 1: the cosine objectives test their argument range once per particle instead of once per 4 coordinates (devmath.cuh)
#ifndef EQS_PSO_COS_DEFER
#define EQS_PSO_COS_DEFER 1
#endif

constexpr long long kNoIndex = 0x7fffffffffffffffLL;
__device__ __forceinline__ double dinf() { return __longlong_as_double(0x7ff0000000000000LL); }

// ---- 256-bit global accesses: one 32-byte sector = 4 dimensions of one particle ------------------------------
struct D4 {
    double e[4];
};

// Both are volatile with a "memory" clobber: to the compiler an asm statement without them is a pure function of its
// register operands, which it may hoist over the stores that produced the data or merge with an earlier identical load
// (the commit-then-speculate sequence of pso_seqw_kernel, the random-access objectives that read x back with plain
// loads).  Program order already is "all loads of a batch, math, stores", so nothing is lost by pinning it.
__device__ __forceinline__ D4 ld256(const double *p)
{
    D4 r;
    asm volatile("ld.global" EQS_PSO_LDHINT ".v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(r.e[0]), "=d"(r.e[1]), "=d"(r.e[2]), "=d"(r.e[3])
                 : "l"(p)
                 : "memory");
    return r;
}

__device__ __forceinline__ void st256(double *p, const D4 &v)
{
    asm volatile("st.global" EQS_PSO_STHINT ".v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.e[0]), "d"(v.e[1]), "d"(v.e[2]), "d"(v.e[3])
                 : "memory");
}

__device__ __forceinline__ long long goff(const PsoDev &P, int q, long long li)
{
    EQS_CHECK(q >= 0 && q < ((P.n + 3) >> 2) && li >= 0 && li < P.ld);
    return ((long long)q * P.ld + li) * 4;
}

// ---- draws: (rp, rg) of particle_swarm.rs:407-408 ------------------------------------------------------------
template <bool REPLAY>
__device__ __forceinline__ void step_draw(const PsoDev &P, uint32_t t, long long gi, int d, double &rp, double &rg)
{
    if constexpr (REPLAY) {
        const double *src = P.replay + ((((long long)t - P.replay_t0) * P.ntotal + gi) * P.n + d) * 2;
        rp = src[0];
        rg = src[1];
    } else {
        const uint4 w = philox_draw(P.key, STREAM_PSO_STEP, t, gi, (uint32_t)d);
        rp = u01(w.x, w.y);
        rg = u01(w.z, w.w);
    }
}

// particle_swarm.rs:409-413 with rustc's evaluation order ((w v) + ((c1 rp)(pb - x))) + ((c2 rg)(G - x)); x += v
__device__ __forceinline__ void pso_update(const PsoDev &P, double &x, double &v, double pb, double g, double rp, double rg)
{
    v = P.w * v + P.c1 * rp * (pb - x) + P.c2 * rg * (g - x);
    x = x + v;
}

// stalled_too_long, particle_swarm.rs:433-441
__device__ __forceinline__ int stall_test(const PsoCtrl *c, double tol)
{
    int stalled = 1;
    for (int j = 0; j < EQS_STALL_WINDOW; ++j)
        if (!(fabs(c->ring[j] - c->Gc) < tol)) stalled = 0;
    return stalled;
}

__device__ __forceinline__ void ring_push(PsoCtrl *c, double v)
{
    c->ring[c->ring_i] = v;
    c->ring_i = (c->ring_i + 1) % EQS_STALL_WINDOW;
}

// one 4-dimension group already in registers: draws, update, objective; `dims` = live dimensions in the group
template <class Obj, bool REPLAY, bool FULL, bool FEED = true>
__device__ __forceinline__ void pso_group_math(const PsoDev &P, uint32_t t, long long gi, int q, const double *sG, ObjAcc &acc,
                                               D4 &x, D4 &v, const D4 &pb, int dims, unsigned &hi_max)
{
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (FULL || j < dims) {
            const int d = 4 * q + j;
            double rp, rg;
            step_draw<REPLAY>(P, t, gi, d, rp, rg);
            pso_update(P, x.e[j], v.e[j], pb.e[j], sG[d], rp, rg);
            if constexpr (Obj::streaming && !FULL) Obj::feed(acc, x.e[j], d);
        }
    }
    if constexpr (Obj::streaming && FULL && FEED) { // the group's 4 coordinates at once
        if constexpr (EQS_PSO_COS_DEFER) objective_feed_vec_defer<Obj, 4>(acc, x.e, 4 * q, hi_max);
        else objective_feed_vec<Obj, 4>(acc, x.e, 4 * q);
    }
}

// K consecutive groups of one particle: all loads first (3K independent 32-byte requests), then math, then stores
template <class Obj, bool REPLAY, int K>
__device__ __forceinline__ void pso_groups(const PsoDev &P, uint32_t t, long long li, int q, const double *sG, ObjAcc &acc,
                                           const double *xin, const double *vin, const double *pbin, double *xout,
                                           double *vout, bool load_pb, unsigned &hi_max)
{
    D4 xs[K], vs[K], ps[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const long long off = goff(P, q + k, li);
        xs[k] = ld256(xin + off);
        vs[k] = ld256(vin + off);
        if (load_pb) ps[k] = ld256(pbin + off);
    }
    if (!load_pb) {
#pragma unroll
        for (int k = 0; k < K; ++k) ps[k] = xs[k];
    }
    constexpr bool batch_feed = EQS_PSO_FEED_BATCH && Obj::streaming && K > 1;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        pso_group_math<Obj, REPLAY, true, !batch_feed>(P, t, pso_gidx(P, li), q + k, sG, acc, xs[k], vs[k], ps[k], 4, hi_max);
        const long long off = goff(P, q + k, li);
        st256(xout + off, xs[k]);
        st256(vout + off, vs[k]);
    }
    if constexpr (batch_feed) { // all 4 K coordinates of the batch in lock step: a coefficient is materialised once per 4 K uses
        double xx[4 * K];
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int j = 0; j < 4; ++j) xx[4 * k + j] = xs[k].e[j];
        if constexpr (EQS_PSO_COS_DEFER) objective_feed_vec_defer<Obj, 4 * K>(acc, xx, 4 * q, hi_max);
        else objective_feed_vec<Obj, 4 * K>(acc, xx, 4 * q);
    }
}

// all dimensions of one particle; returns its new cost
template <class Obj, bool REPLAY, int UQ = EQS_PSO_UQ>
__device__ __forceinline__ double pso_move(const PsoDev &P, uint32_t t, long long li, const double *sG, const double *xin,
                                           const double *vin, const double *pbin, double *xout, double *vout, bool load_pb)
{
    const int n = P.n, nfull = n >> 2, rem = n & 3;
    ObjAcc acc;
    unsigned hi_max = 0;
    if constexpr (Obj::streaming) Obj::begin(acc, n);
    int q = 0;
    for (; q + UQ <= nfull; q += UQ) {
        pso_groups<Obj, REPLAY, UQ>(P, t, li, q, sG, acc, xin, vin, pbin, xout, vout, load_pb, hi_max);
    }
    for (; q < nfull; ++q) pso_groups<Obj, REPLAY, 1>(P, t, li, q, sG, acc, xin, vin, pbin, xout, vout, load_pb, hi_max);
    if (rem) { // ragged last group: the padding dimensions are carried through unchanged
        const long long off = goff(P, q, li);
        D4 x = ld256(xin + off), v = ld256(vin + off), pb = x;
        if (load_pb) pb = ld256(pbin + off);
        pso_group_math<Obj, REPLAY, false>(P, t, pso_gidx(P, li), q, sG, acc, x, v, pb, rem, hi_max);
        st256(xout + off, x);
        st256(vout + off, v);
    }
    if constexpr (Obj::streaming) {
        // a deferred range test fired (never on sane inputs): the sums hold garbage, evaluate the stored position again
        // through the scalar path, whose cosine handles every argument
        if (EQS_PSO_COS_DEFER && hi_max >= kCosHiLimit) return objective_on<Obj>(GroupedView{xout + li * 4, P.ld * 4}, n);
        return Obj::end(acc, n);
    } else {
        return Obj::eval(GroupedView{xout + li * 4, P.ld * 4}, n);
    }
}

// end of a particle's pass in synchronous mode: pbest cost + flag byte (:417-419), running arg-min
__device__ __forceinline__ void pso_finish_particle(const PsoDev &P, long long li, double c, double pbc_old, unsigned flags,
                                                    MinLoc &best)
{
    const bool stale = flags & 1u, role = (flags >> 1) & 1u;
    const bool imp = c < pbc_old;
    if (imp) P.pbc[li] = c;
    const unsigned nf = (imp ? 1u : 0u) | ((role != stale) ? 2u : 0u); // a stale particle moved to the other buffer
    if (nf != flags) P.pbflag[li] = (unsigned char)nf;
    const double cs = sanitize_cost(c);
    if (cs < best.c) {
        best.c = cs;
        best.i = pso_gidx(P, li);
    }
}

// the last block to finish reduces the per-block partials; returns true in that block only (all threads)
__device__ __forceinline__ bool grid_minloc_last_block(const PsoDev &P, MinLoc mine, MinLoc *s_red, MinLoc &winner)
{
    __shared__ int s_last;
    __shared__ MinLoc s_win;
    MinLoc b = block_minloc(mine, s_red);
    if (threadIdx.x == 0) {
        P.part_cost[blockIdx.x] = b.c;
        P.part_idx[blockIdx.x] = b.i;
        __threadfence();
        const unsigned ticket = atomicAdd(&P.ctrl->block_counter, 1u);
        s_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return false;
    __threadfence();
    MinLoc m{dinf(), kNoIndex};
    for (int j = threadIdx.x; j < (int)gridDim.x; j += blockDim.x)
        m = minloc_better(m, MinLoc{__ldcg(P.part_cost + j), __ldcg(P.part_idx + j)});
    m = block_minloc(m, s_red);
    if (threadIdx.x == 0) {
        s_win = m;
        P.ctrl->block_counter = 0;
    }
    __syncthreads();
    winner = s_win;
    return true;
}

// this rank's exchange record: (cost, global index, the particle's position).  src == nullptr: the buffer that
// holds x is named by the particle's role bit.
// CG: the data was written by OTHER CTAs of this grid -> read it from L2 (__ldcg); false inside the single-CTA kernel.
template <bool CG = true>
__device__ __forceinline__ void write_record(const PsoDev &P, double cost, long long gidx, const double *src)
{
    const bool none = gidx == kNoIndex || gidx < 0;
    const long long li = none ? 0 : pso_lidx(P, gidx);
    if (!src && !none) src = ((CG ? __ldcg(P.pbflag + li) : P.pbflag[li]) & 2u) ? P.pb : P.x;
    if (threadIdx.x == 0) {
        P.send[REC_COST] = cost;
        P.send[REC_IDX] = none ? -1.0 : (double)gidx;
    }
    for (int d = threadIdx.x; d < P.n; d += blockDim.x) {
        const double *p = src + goff(P, d >> 2, li) + (d & 3);
        P.send[REC_X + d] = none ? 0.0 : (CG ? __ldcg(p) : *p);
    }
}

// lowest cost, then lowest global index, over the ranks' records; -1 when every record is empty
__device__ __forceinline__ int best_record(const double *recs, int world, long long R, double &bc)
{
    int win = -1;
    double bi = 0.0;
    bc = dinf();
    for (int r = 0; r < world; ++r) {
        const double *rec = recs + r * R;
        const double idx = rec[REC_IDX], c = rec[REC_COST];
        if (idx < 0.0) continue;
        if (win < 0 || c < bc || (c == bc && idx < bi)) {
            win = r;
            bc = c;
            bi = idx;
        }
    }
    return win;
}

// synchronous gbest update (one per pass): particle_swarm.rs:420-425 applied to the pass's arg-min
__device__ __forceinline__ void pso_apply_step(const PsoDev &P, const double *recs, int world)
{
    __shared__ int s_winner;
    const long long R = record_doubles(P.n);
    if (threadIdx.x == 0) {
        PsoCtrl *c = P.ctrl;
        double bc;
        const int win = best_record(recs, world, R, bc);
        const bool improved = win >= 0 && bc < c->Gc;
        if (improved) {
            ring_push(c, c->Gc);
            c->Gc = bc;
            c->improvements += 1;
        }
        c->iter += 1;
        c->stalled = stall_test(c, P.tol);
        s_winner = improved ? win : -1;
    }
    __syncthreads();
    if (s_winner >= 0)
        for (int d = threadIdx.x; d < P.n; d += blockDim.x) P.G[d] = recs[s_winner * R + REC_X + d];
}

// ---- the exchange fused into the step kernel (one kernel per pass at any world size) ---------------------------
// Called by the LAST CTA of the step kernel once P.send holds this rank's record.  Only (cost, index, position) matter in
// a step; they go to every peer with the low-latency protocol of common.cuh (ll_*: payload and arrival in one 8-byte
// store, one NVLink one-way trip; round 1's flag protocol needed stores, a system fence's round trip and a flag store:
// 7 us per pass on two GPUs against 3), then the same gbest selection as pso_step_apply_kernel.
__device__ __forceinline__ void pso_p2p_exchange_apply(const PsoDev &P)
{
    __shared__ int s_bad, s_wr;
    const long long R = record_doubles(P.n);
    const int tid = threadIdx.x, world = P.world, n = P.n;
    if (tid == 0) s_bad = 0;
    const LLExchange X = ll_begin(P.mailbox, P.peers, P.rank, world); // barrier inside: P.send (write_record) is complete
    const bool have = P.send[REC_IDX] >= 0.0;
    const int nd = have ? 2 + n : 2;
    for (int d = tid; d < nd; d += blockDim.x)
        ll_put(X, d, d == 0 ? P.send[REC_COST] : (d == 1 ? P.send[REC_IDX] : P.send[REC_X + d - 2]));
    if (tid < 2 * world) { // every rank's (cost, index)
        const int r = tid >> 1, j = tid & 1;
        double v = j ? P.send[REC_IDX] : P.send[REC_COST];
        if (r != P.rank && !ll_get(X, r, j, v, P.p2p_timeout_ns)) {
            s_bad = 1;
            v = j ? -1.0 : dinf();
        }
        P.recv[r * R + (j ? REC_IDX : REC_COST)] = v;
    }
    __syncthreads();
    if (tid == 0) {
        double bc;
        s_wr = best_record(P.recv, world, R, bc);
    }
    __syncthreads();
    const int wr = s_wr;
    if (wr >= 0) { // only the winner's position is needed
        for (int d = tid; d < n; d += blockDim.x) {
            double v = P.send[REC_X + d];
            if (wr != P.rank && !ll_get(X, wr, 2 + d, v, P.p2p_timeout_ns)) s_bad = 1;
            P.recv[wr * R + REC_X + d] = v;
        }
    }
    __syncthreads();
    if (tid == 0) {
        ll_end(X);
        if (s_bad) P.ctrl->p2p_timeout = 1;
    }
    __threadfence_block();
    __syncthreads();
    pso_apply_step(P, P.recv, world);
}

__device__ __forceinline__ void load_gbest(const PsoDev &P, double *sG)
{
    const int npad = (P.n + 3) & ~3;
    for (int d = threadIdx.x; d < npad; d += blockDim.x) sG[d] = d < P.n ? P.G[d] : 0.0;
    __syncthreads();
}

// ---- kernels -------------------------------------------------------------------------------------------------

// :357-359.  G already holds x0 (H2D copy).
template <class Obj> __global__ void pso_x0_kernel(const PsoDev P)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    PsoCtrl *c = P.ctrl;
    c->Gc = objective_on<Obj>(StridedView{P.G, 1}, P.n);
    for (int j = 0; j < EQS_STALL_WINDOW; ++j) c->ring[j] = dinf();
    c->ring_i = 0;
    c->iter = 0;
    c->improvements = 0;
    c->spec_first = -1;
    c->stalled = 0;
    c->pass_done = 0;
    c->block_counter = 0;
    c->p2p_timeout = 0;
    c->seq_start = 0;
    c->seq_commit_lo = 1;
    c->seq_commit_hi = 0;
    c->seq_window = 0;
    c->seq_target = 0;
    c->seq_parity = 0;
    c->seq_commit_buf = 0;
    c->seq_round = 0ULL;
    c->seq_barrier = 0u;
    c->seq_xdone = 0ULL;
    c->seq_slot[0] = c->seq_slot[1] = c->seq_slot[2] = ~0ULL;
}

// :361-389 without the (order-dependent) gbest bookkeeping, which pso_ring_scan_kernel reconstructs
template <class Obj, bool REPLAY> __global__ void __launch_bounds__(PSO_BLOCK) pso_init_kernel(const PsoDev P)
{
    __shared__ MinLoc s_red[32];
    const int n = P.n, nq = (n + 3) >> 2;
    const long long ngroups = (P.nloc + PSO_GROUP - 1) / PSO_GROUP;
    MinLoc best{dinf(), kNoIndex};
    for (long long g = blockIdx.x; g < ngroups; g += gridDim.x) {
        const long long li = g * PSO_GROUP + threadIdx.x;
        double c = dinf();
        if (li < P.nloc) {
            const long long gi = pso_gidx(P, li);
            ObjAcc acc;
            if constexpr (Obj::streaming) Obj::begin(acc, n);
            for (int q = 0; q < nq; ++q) {
                D4 x4, v4;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int d = 4 * q + j;
                    x4.e[j] = 0.0;
                    v4.e[j] = 0.0;
                    if (d < n) {
                        double pu, vu;
                        if constexpr (REPLAY) {
                            pu = P.replay[gi * 2 * n + d];
                            vu = P.replay[gi * 2 * n + n + d];
                        } else {
                            const uint4 w = philox_draw(P.key, STREAM_PSO_INIT, 0u, gi, (uint32_t)d);
                            pu = u01(w.x, w.y);
                            vu = u01(w.z, w.w);
                        }
                        const double lo = P.lower[d], hi = P.upper[d];
                        x4.e[j] = pu * (hi - lo) + lo;              // Uniform::new_inclusive(lo, hi), :124-129
                        const double dist = fabs(hi - lo);          // :135
                        v4.e[j] = vu * (dist + dist) + (-dist);     // Uniform::new_inclusive(-dist, dist)
                        if constexpr (Obj::streaming) Obj::feed(acc, x4.e[j], d);
                    }
                }
                const long long off = goff(P, q, li);
                st256(P.x + off, x4);
                st256(P.v + off, v4);
                if (!P.sync) st256(P.pb + off, x4);                 // best_position = position, :385 (sync: flag below)
            }
            double cost;
            if constexpr (Obj::streaming) cost = Obj::end(acc, n);
            else cost = Obj::eval(GroupedView{P.x + li * 4, P.ld * 4}, n);
            P.pbc[li] = cost;
            P.cost[li] = cost;
            P.pbflag[li] = P.sync ? 1 : 0;                          // sync: stale = pbest position is the position
            c = sanitize_cost(cost);
            if (c < best.c) {
                best.c = c;
                best.i = gi;
            }
        }
        const MinLoc gm = block_minloc(MinLoc{c, li}, s_red);
        if (threadIdx.x == 0) P.group_min[g] = gm.c;
    }
    MinLoc win;
    if (grid_minloc_last_block(P, best, s_red, win)) write_record(P, win.c, win.i, P.x);
}

// The reference updates gbest INSIDE the init loop (:376-380), pushing the old gbest cost on the stall ring at
// every record low, in particle order.  Which positions are record lows is an exclusive prefix-min problem:
// particle i pushes e_i = min(seed, c_0..c_{i-1}) iff c_i < e_i.  One CTA walks the per-tile minima (1024 tiles
// per scan), then only the tiles that contain a record low (O(log N) of them for random costs).
// One kernel for both distributions.  Contiguous blocks: the rank walks its own tiles, seeded with f(x0) and the minima
// of the lower ranks (their records).  Tile-cyclic: the rank walks ALL global tiles in order -- their minima were gathered
// from every rank (gm_all) -- and opens only the flagged tiles it owns.  Either way it reports how many pushes its own
// particles made and the last <= 50 of them with the pushing particle's global index; pso_init_apply_kernel merges the
// ranks' lists by that position.
__global__ void __launch_bounds__(1024) pso_ring_scan_kernel(const PsoDev P)
{
    __shared__ double s_scan_d[33];
    __shared__ int s_scan_i[33];
    __shared__ long long s_list[1024];
    __shared__ double s_seedlist[1024];
    __shared__ double s_ring[EQS_STALL_WINDOW], s_rpos[EQS_STALL_WINDOW];
    __shared__ double s_seed;
    const int tid = threadIdx.x;
    const long long R = record_doubles(P.n);
    if (tid == 0) {
        double seed = P.ctrl->Gc; // f(x0), :358
        if (!P.cyclic)
            for (int r = 0; r < P.rank; ++r) { // particles of lower ranks come first in the reference's order
                const double *rec = P.recv + r * R;
                if (rec[REC_IDX] >= 0.0 && rec[REC_COST] < seed) seed = rec[REC_COST];
            }
        s_seed = seed;
    }
    if (tid < EQS_STALL_WINDOW) {
        s_ring[tid] = dinf();
        s_rpos[tid] = -1.0;
    }
    __syncthreads();
    double running = s_seed;
    long long count = 0;
    // tiles in the reference's order: this rank's own (contiguous) or every rank's (cyclic)
    const long long ntiles = P.cyclic ? (P.ntotal + PSO_GROUP - 1) / PSO_GROUP : (P.nloc + PSO_GROUP - 1) / PSO_GROUP;
    for (long long g0 = 0; g0 < ntiles; g0 += 1024) {
        const long long T = g0 + tid;
        double gm = dinf();
        if (T < ntiles) gm = P.cyclic ? P.gm_all[(T % P.world) * P.tiles_per_rank + T / P.world] : P.group_min[T];
        double chunk_min;
        double ex = block_exclusive_scan(gm, dinf(), OpMin(), s_scan_d, chunk_min);
        ex = ex < running ? ex : running;
        const int flagged = gm < ex;
        int nflag;
        const int pos = block_exclusive_scan(flagged, 0, OpAddI(), s_scan_i, nflag);
        if (flagged) {
            s_list[pos] = T;
            s_seedlist[pos] = ex;
        }
        __syncthreads();
        for (int q = 0; q < nflag; ++q) {
            const long long Tq = s_list[q];
            if (P.cyclic && Tq % P.world != P.rank) continue; // another rank's particles (uniform over the block)
            const long long li = (P.cyclic ? Tq / P.world : Tq) * PSO_GROUP + tid;
            const double gseed = s_seedlist[q];
            const double c = (tid < PSO_GROUP && li < P.nloc) ? sanitize_cost(P.cost[li]) : dinf();
            double dummy;
            double e = block_exclusive_scan(c, dinf(), OpMin(), s_scan_d, dummy);
            e = e < gseed ? e : gseed;
            const int rec = c < e;
            int tot;
            const long long rk = count + block_exclusive_scan(rec, 0, OpAddI(), s_scan_i, tot);
            if (rec && rk >= count + tot - EQS_STALL_WINDOW) {
                s_ring[rk % EQS_STALL_WINDOW] = e;
                s_rpos[rk % EQS_STALL_WINDOW] = (double)pso_gidx(P, li);
            }
            count += tot;
            __syncthreads();
        }
        running = chunk_min < running ? chunk_min : running;
    }
    __syncthreads();
    if (tid < EQS_STALL_WINDOW) {
        const long long kept = count < EQS_STALL_WINDOW ? count : EQS_STALL_WINDOW;
        P.send[REC_RING + tid] = tid < kept ? s_ring[(count - kept + tid) % EQS_STALL_WINDOW] : dinf();
        P.send[REC_RPOS + tid] = tid < kept ? s_rpos[(count - kept + tid) % EQS_STALL_WINDOW] : -1.0;
    }
    if (tid == 0) P.send[REC_PUSHES] = (double)count;
}

// merge the ranks' init records: gbest (strict <, lowest index) and the ring in global particle order
__global__ void pso_init_apply_kernel(const PsoDev P)
{
    __shared__ int s_winner;
    const long long R = record_doubles(P.n);
    if (threadIdx.x == 0) {
        PsoCtrl *c = P.ctrl;
        double bc;
        const int win = best_record(P.recv, P.world, R, bc);
        const bool improved = win >= 0 && bc < c->Gc;
        if (improved) c->Gc = bc;
        // every rank kept its own last <= 50 pushes in position order; the last 50 of ALL pushes are the 50 largest
        // positions of their union (a push among the global last 50 is among its own rank's last 50), merged from the
        // back.  Push number p (0-based, global) goes to ring slot p % 50 like CircularArray::push (:24-46).
        long long total = 0;
        double *recv = P.recv; // word 3 of every record (padding) counts the entries of that rank still to be merged
        for (int r = 0; r < P.world; ++r) {
            const long long K = (long long)recv[r * R + REC_PUSHES];
            total += K;
            recv[r * R + 3] = (double)(K < EQS_STALL_WINDOW ? K : EQS_STALL_WINDOW);
        }
        const long long m = total < EQS_STALL_WINDOW ? total : EQS_STALL_WINDOW;
        for (long long j = m - 1; j >= 0; --j) {
            int pick = -1;
            double ppos = -1.0;
            for (int r = 0; r < P.world; ++r) {
                const int left = (int)recv[r * R + 3];
                if (left > 0 && recv[r * R + REC_RPOS + left - 1] > ppos) {
                    ppos = recv[r * R + REC_RPOS + left - 1];
                    pick = r;
                }
            }
            if (pick < 0) break;
            const int left = (int)recv[pick * R + 3];
            c->ring[(total - m + j) % EQS_STALL_WINDOW] = recv[pick * R + REC_RING + left - 1];
            recv[pick * R + 3] = (double)(left - 1);
        }
        c->ring_i = total % EQS_STALL_WINDOW;
        c->improvements = total;
        c->stalled = stall_test(c, P.tol);
        s_winner = improved ? win : -1;
    }
    __syncthreads();
    if (s_winner >= 0)
        for (int d = threadIdx.x; d < P.n; d += blockDim.x) P.G[d] = P.recv[s_winner * R + REC_X + d];
}

// K2 + K3: the synchronous population step.  One thread owns one particle and walks its 4-dimension groups with
// 256-bit accesses; Philox draws, update and objective stay in registers.  HBM sees x, v (and pbest unless stale)
// once in and x, v once out.
template <class Obj, bool REPLAY>
__global__ void __launch_bounds__(PSO_BLOCK, EQS_PSO_MINB) pso_step_kernel(const PsoDev P, int inline_apply)
{
    extern __shared__ double sG[];
    __shared__ MinLoc s_red[32];
    PsoCtrl *ctrl = P.ctrl;
    if (ctrl->stalled) return; // :395-397 -- passes after the stall test fires are no-ops
    const uint32_t t = (uint32_t)ctrl->iter;
    load_gbest(P, sG);
    MinLoc best{dinf(), kNoIndex};
    for (long long li = (long long)blockIdx.x * blockDim.x + threadIdx.x; li < P.nloc;
         li += (long long)gridDim.x * blockDim.x) {
        const unsigned f = P.pbflag[li];
        const bool stale = f & 1u, role = f & 2u;
        const double pbc_old = P.pbc[li]; // issued now, consumed after the last dimension
        double *xb = role ? P.pb : P.x, *ob = role ? P.x : P.pb;
        const double c = pso_move<Obj, REPLAY>(P, t, li, sG, xb, P.v, ob, stale ? ob : xb, P.v, !stale);
        pso_finish_particle(P, li, c, pbc_old, f, best);
    }
    MinLoc win;
    if (grid_minloc_last_block(P, best, s_red, win)) {
        write_record(P, win.c, win.i, nullptr);
        if (inline_apply == 1) {
            __syncthreads();
            pso_apply_step(P, P.send, 1);
        } else if (inline_apply == 2) {
            pso_p2p_exchange_apply(P);
        }
    }
}

// (Per-thread L2 prefetches -- prefetch.global.L2 / CCTL.E.PF2 of exactly the sectors the thread loads 1, 2 or 4 batches
// later, next particle included -- were built and measured in round 2 and are much slower: C4 0.81 / 0.72 / 0.62 of measured
// HBM against 0.88 without, C2 0.81 / 0.71 / 0.62 against 0.88.  profiles/r02_sweep.md section 3.)
// (A variant of this kernel that streams the population through a two-stage shared-memory ring with per-thread cp.async
// was built and measured in round 2 -- bit-identical, but slower: 0.64-0.87 of measured HBM on C2 and 0.59-0.67 on C4
// against 0.88 / 0.82 here, because a 16-byte LDGSTS per lane doubles the L2 sector requests and the ring's LDS / LDGSTS
// traffic stalls the MIO queue.  profiles/r02_sweep.md has the ncu rows; the code is in the history, not in the product.)

__global__ void pso_step_apply_kernel(const PsoDev P)
{
    if (P.ctrl->stalled) return; // the step kernel of a stalled pass did nothing: neither may this (:395-397)
    pso_apply_step(P, P.recv, P.world);
}

// ---- sequential-exact mode (the reference's semantics, F6) ---------------------------------------------------
// speculate: move every particle with global index >= start against the CURRENT gbest into (x2, v2, cost)
// without committing, and find the FIRST particle whose cost beats gbest.
template <class Obj, bool REPLAY>
__global__ void __launch_bounds__(PSO_BLOCK) pso_spec_kernel(const PsoDev P, long long start_g)
{
    extern __shared__ double sG[];
    __shared__ MinLoc s_red[32];
    PsoCtrl *ctrl = P.ctrl;
    if (ctrl->stalled) return;
    const uint32_t t = (uint32_t)ctrl->iter;
    const double Gc = ctrl->Gc;
    load_gbest(P, sG);
    MinLoc first{0.0, kNoIndex}; // min over index only: all costs equal => lowest index wins
    for (long long li = (long long)blockIdx.x * blockDim.x + threadIdx.x; li < P.nloc;
         li += (long long)gridDim.x * blockDim.x) {
        const long long gi = P.first + li;
        if (gi < start_g) continue;
        const double cost = pso_move<Obj, REPLAY>(P, t, li, sG, P.x, P.v, P.pb, P.x2, P.v2, true);
        P.cost[li] = cost;
        // :417,:420 -- the gbest test is nested in the pbest test; the two differ when the pbest cost is NaN (then the
        // particle can never update either), so both are spelled out
        if (cost < P.pbc[li] && cost < Gc && gi < first.i) first.i = gi;
    }
    MinLoc win;
    if (grid_minloc_last_block(P, first, s_red, win)) {
        const bool none = win.i == kNoIndex;
        write_record(P, none ? dinf() : __ldcg(P.cost + (win.i - P.first)), win.i, P.x2);
    }
}

__global__ void pso_spec_apply_kernel(const PsoDev P)
{
    __shared__ int s_winner;
    const long long R = record_doubles(P.n);
    if (threadIdx.x == 0) {
        PsoCtrl *c = P.ctrl;
        int win = -1;
        double bi = 0.0;
        for (int r = 0; r < P.world; ++r) {
            const double idx = P.recv[r * R + REC_IDX];
            if (idx >= 0.0 && (win < 0 || idx < bi)) {
                win = r;
                bi = idx;
            }
        }
        if (win >= 0) { // :420-425
            ring_push(c, c->Gc);
            c->Gc = P.recv[win * R + REC_COST];
            c->improvements += 1;
            c->spec_first = (long long)bi;
        } else {
            c->spec_first = -1;
        }
        const int done = (win < 0) || ((long long)bi == P.ntotal - 1);
        c->pass_done = done;
        if (done) {
            c->iter += 1;
            c->stalled = stall_test(c, P.tol);
        }
        s_winner = win;
    }
    __syncthreads();
    if (s_winner >= 0)
        for (int d = threadIdx.x; d < P.n; d += blockDim.x) P.G[d] = P.recv[s_winner * R + REC_X + d];
}

// commit the speculated state of global indices [start_g, end_g]: x, v <- x2, v2 and the pbest update :417-419
__global__ void __launch_bounds__(PSO_BLOCK) pso_commit_kernel(const PsoDev P, long long start_g, long long end_g)
{
    const int nq = (P.n + 3) >> 2;
    for (long long li = (long long)blockIdx.x * blockDim.x + threadIdx.x; li < P.nloc;
         li += (long long)gridDim.x * blockDim.x) {
        const long long gi = P.first + li;
        if (gi < start_g || gi > end_g) continue;
        const double c = P.cost[li];
        const bool imp = c < P.pbc[li];
        if (imp) P.pbc[li] = c;
        for (int q = 0; q < nq; ++q) {
            const long long off = goff(P, q, li);
            const D4 xv = ld256(P.x2 + off);
            st256(P.x + off, xv);
            st256(P.v + off, ld256(P.v2 + off));
            if (imp) st256(P.pb + off, xv);
        }
    }
}

// Sequential-exact mode on one rank, device-driven.  A round speculates only a WINDOW of particles [start, start + W)
// against the current gbest: A = the committed (x, v), B = the other buffer pair, A -> B.  The last CTA finds the first
// improving particle i* of the window (none: the whole window is final), moves gbest (:420-425), records [start, i*]
// as final and advances start; the NEXT launch performs the pbest update (:417-419) of that range before it speculates.
// When start reaches N the pass is over: iter + 1, stall test, and A / B swap roles -- nothing is copied.  A repair
// therefore costs at most one window instead of the rest of the population, and no host round trip is needed: the
// host only enqueues launches (no-ops once iter == seq_target).  With k improvements per pass, a fixed cost c_r per
// round and c_p per particle, a pass costs (N/W + k/2) c_r + (N + k W/2) c_p, least at W = sqrt(2 N c_r / (k c_p)):
// W is re-chosen after every pass from a running estimate of k (measured: profiles/r01_sweep.md).
__global__ void pso_seq_arm_kernel(const PsoDev P, long long iters, long long wmin, long long wmax, long long w0, double wscale)
{
    PsoCtrl *c = P.ctrl;
    c->seq_barrier = 0u;
    c->seq_target = c->iter + iters;
    c->seq_wmin = wmin;
    c->seq_wmax = wmax;
    c->seq_wscale = wscale;
    if (c->seq_window == 0) { // first use after init
        c->seq_window = w0;
        c->seq_imp0 = c->improvements;
        c->seq_kest = 4.0;
    }
    if (c->seq_window < wmin) c->seq_window = wmin;
    if (c->seq_window > wmax) c->seq_window = wmax;
}

// Round 2: the rounds no longer are launches.  ONE persistent kernel (cooperative launch: every CTA resident) runs all
// rounds of all `iters` passes; the rounds are separated by a grid-wide barrier (one atomic per CTA on a monotonic
// counter) instead of a kernel boundary.  Every CTA keeps its OWN copy of the loop state (gbest cost, stall ring, window,
// parity, pass counter) in shared memory and advances it with the same deterministic rule from the same inputs, so the
// only thing that crosses CTAs per round is the first improving index: an atomicMin on one of three rotating slots before
// the barrier, one load after it (the slot of round r is reset in round r+1's shadow for round r+3... see below).  The
// improving particle's position is read straight from the speculated buffer by every CTA into its shared-memory gbest.
// Fixed cost of a round: ~11 us as a launch (launch gap + tail + last-CTA reduction) -> the barrier's ~2 us, which also
// moves the best window size down by 2x (W ~ sqrt(c_r)).
struct SeqLocal {
    double Gc;
    double ring[EQS_STALL_WINDOW];
    double kest;
    long long ring_i, iter, improvements, imp0;
    long long start, window, commit_lo, commit_hi;
    unsigned long long round;
    int parity, commit_buf, stalled;
};

// all threads of every CTA call; returns when every CTA of the grid has arrived.  `counter` only grows: arrival number
// k of this launch waits for k * gridDim.x (reset to 0 by pso_seq_arm_kernel before the launch).
__device__ __forceinline__ void grid_barrier(unsigned int *counter, unsigned int &target)
{
    __threadfence(); // this thread's writes, before the block's arrival is published
    __syncthreads();
    target += gridDim.x;
    if (threadIdx.x == 0) {
        atomicAdd(counter, 1u);
        unsigned int seen;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
        } while (seen < target);
    }
    __syncthreads();
}

// Several ranks (tile-cyclic distribution, pso.cuh): every rank runs this kernel on its share of the SAME global window,
// with the same replicated loop state.  The LAST CTA of a rank to finish its share of the round (ticket on the arrival
// counter: no separate grid barrier) sends (cost, first improving global index, its speculated position) to the other
// ranks with the low-latency protocol of common.cuh (ll_*: one NVLink one-way trip, no fence, no flag), takes the lowest
// index of all ranks and publishes (cost, index, position) in local memory; every CTA of the rank waits for that.
__device__ __forceinline__ void pso_seq_exchange(const PsoDev &P, const double *xb, unsigned long long round)
{
    __shared__ unsigned long long s_loc;
    __shared__ double s_hdr[MB_MAX_WORLD][2];
    __shared__ int s_wr, s_bad;
    PsoCtrl *ctrl = P.ctrl;
    const int tid = threadIdx.x, world = P.world, n = P.n;
    if (tid == 0) {
        unsigned long long w;
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(&ctrl->seq_slot[round % 3]) : "memory");
        s_loc = w;
        s_bad = 0;
        // the slot of round r+2: its last readers (this CTA, round r-1) are done, its next writers (round r+2) wait for
        // the publication of rounds r and r+1, which come after this store
        ctrl->seq_slot[(round + 2) % 3] = ~0ULL;
    }
    const LLExchange X = ll_begin(P.mailbox, P.peers, P.rank, world); // barrier inside: s_loc is visible
    const bool have = s_loc != ~0ULL;
    const long long li = have ? pso_lidx(P, (long long)s_loc) : 0;
    // this rank's record: [0] cost, [1] global index (-1 = none), [2 .. 2+n) position (only with an index)
    const int nd = have ? 2 + n : 2;
    for (int d = tid; d < nd; d += blockDim.x) {
        double v;
        if (d == 0) v = have ? __ldcg(P.cost + li) : dinf();
        else if (d == 1) v = have ? (double)(long long)s_loc : -1.0;
        else v = __ldcg(xb + goff(P, (d - 2) >> 2, li) + ((d - 2) & 3));
        ll_put(X, d, v);
        if (d < 2) s_hdr[P.rank][d] = v;
    }
    if (tid < 2 * world && (tid >> 1) != P.rank) { // the peers' (cost, index)
        double v = 0.0;
        if (!ll_get(X, tid >> 1, tid & 1, v, P.p2p_timeout_ns)) s_bad = 1;
        s_hdr[tid >> 1][tid & 1] = v;
    }
    __syncthreads();
    if (tid == 0) {
        int wr = -1;
        double bi = 0.0;
        for (int r = 0; r < world; ++r) {
            const double idx = s_hdr[r][1];
            if (idx >= 0.0 && (wr < 0 || idx < bi)) {
                wr = r;
                bi = idx;
            }
        }
        s_wr = s_bad ? -1 : wr;
    }
    __syncthreads();
    const int wr = s_wr;
    if (wr >= 0) { // the winner's position: from the mailbox, or this rank's own particle
        for (int d = tid; d < n; d += blockDim.x) {
            double v = 0.0;
            if (wr == P.rank) v = __ldcg(xb + goff(P, d >> 2, li) + (d & 3));
            else if (!ll_get(X, wr, 2 + d, v, P.p2p_timeout_ns)) s_bad = 1;
            P.recv[2 + d] = v;
        }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        ll_end(X);
        if (s_bad) ctrl->p2p_timeout = 1;
        P.recv[0] = wr >= 0 ? s_hdr[wr][0] : dinf();
        P.recv[1] = s_bad ? -2.0 : (wr >= 0 ? s_hdr[wr][1] : -1.0); // -2: a peer never arrived, every CTA leaves the loop
        __threadfence();
        const unsigned long long done = round + 1;
        asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(&ctrl->seq_xdone), "l"(done) : "memory");
    }
}

template <class Obj, bool REPLAY>
__global__ void __launch_bounds__(PSO_BLOCK) pso_seqp_kernel(const PsoDev P)
{
    extern __shared__ double sG[];
    __shared__ MinLoc s_red[32];
    __shared__ SeqLocal S;
    __shared__ long long s_first;   // first improving GLOBAL index of the round, -1 = none, -2 = abort (peer timeout)
    __shared__ double s_first_cost;
    __shared__ int s_last;
    PsoCtrl *ctrl = P.ctrl;
    const int tid = threadIdx.x;
    const long long N = P.ntotal;   // the loop state below is in GLOBAL particle indices
    const bool multi = P.world > 1;
    if (tid == 0) { // every CTA reads the same control block: nothing writes it until the kernel's last lines
        S.Gc = ctrl->Gc;
        for (int j = 0; j < EQS_STALL_WINDOW; ++j) S.ring[j] = ctrl->ring[j];
        S.ring_i = ctrl->ring_i;
        S.iter = ctrl->iter;
        S.improvements = ctrl->improvements;
        S.imp0 = ctrl->seq_imp0;
        S.kest = ctrl->seq_kest;
        S.start = ctrl->seq_start;
        S.window = ctrl->seq_window;
        S.commit_lo = ctrl->seq_commit_lo;
        S.commit_hi = ctrl->seq_commit_hi;
        S.commit_buf = ctrl->seq_commit_buf;
        S.parity = ctrl->seq_parity;
        S.stalled = ctrl->stalled;
        S.round = ctrl->seq_round;
    }
    const long long target_iter = ctrl->seq_target, wmin = ctrl->seq_wmin, wmax = ctrl->seq_wmax;
    const double wscale = ctrl->seq_wscale;
    load_gbest(P, sG); // ends with a barrier: S is visible
    unsigned int arrivals = 0;
    const long long stride = (long long)gridDim.x * blockDim.x, gtid = (long long)blockIdx.x * blockDim.x + tid;
    // local particle li is always handled by thread li mod stride, so a particle's pending commit precedes its speculation
    auto first_at = [&](long long lo) {
        long long r = (gtid - lo) % stride;
        if (r < 0) r += stride;
        return lo + r;
    };
    for (;;) {
        const bool active = !S.stalled && S.iter < target_iter;
        const long long clo = S.commit_lo, chi = S.commit_hi, start = S.start, W = S.window;
        const int cbuf = S.commit_buf, parity = S.parity;
        const double Gc = S.Gc;
        const unsigned long long round = S.round;
        if (clo <= chi) { // :417-419 for the particles the previous round made final (this rank's share of [clo, chi])
            const double *src = cbuf ? P.x2 : P.x;
            const int nq = (P.n + 3) >> 2;
            const long long l1 = pso_lcount(P, chi + 1);
            for (long long li = first_at(pso_lcount(P, clo)); li < l1; li += stride) {
                const double c = P.cost[li];
                if (c < P.pbc[li]) {
                    P.pbc[li] = c;
                    for (int q = 0; q < nq; ++q) {
                        const long long off = goff(P, q, li);
                        st256(P.pb + off, ld256(src + off));
                    }
                }
            }
        }
        if (!active) break; // the same decision in every CTA (of every rank); the pending commit above was the last thing to do
        const uint32_t t = (uint32_t)S.iter;
        const long long wend = start + W < N ? start + W : N;
        const double *xa = parity ? P.x2 : P.x, *va = parity ? P.v2 : P.v;
        double *xb = parity ? P.x : P.x2, *vb = parity ? P.v : P.v2;
        MinLoc first{0.0, kNoIndex}; // min over the index only
        const long long s1 = pso_lcount(P, wend);
        for (long long li = first_at(pso_lcount(P, start)); li < s1; li += stride) {
            const double cost = pso_move<Obj, REPLAY, EQS_PSO_SEQ_UQ>(P, t, li, sG, xa, va, P.pb, xb, vb, true);
            P.cost[li] = cost;
            if (cost < P.pbc[li] && cost < Gc && li < first.i) first.i = li; // :417,:420 (nested; see pso_spec_kernel)
        }
        const MinLoc b = block_minloc(first, s_red);
        if (tid == 0 && b.i != kNoIndex) atomicMin(&ctrl->seq_slot[round % 3], (unsigned long long)pso_gidx(P, b.i));
        if (multi) { // arrive; the last CTA of this rank talks to the other ranks, everybody waits for what it publishes
            __threadfence();
            __syncthreads();
            arrivals += gridDim.x;
            if (tid == 0) s_last = atomicAdd(&ctrl->seq_barrier, 1u) == arrivals - 1;
            __syncthreads();
            if (s_last) pso_seq_exchange(P, xb, round);
        } else {
            grid_barrier(&ctrl->seq_barrier, arrivals);
        }
        if (tid == 0) {
            if (multi) {
                unsigned long long seen;
                do {
                    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(&ctrl->seq_xdone) : "memory");
                } while (seen < round + 1);
                s_first = (long long)__ldcg(P.recv + 1);
                s_first_cost = __ldcg(P.recv);
            } else {
                unsigned long long w;
                asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(&ctrl->seq_slot[round % 3]) : "memory");
                s_first = w == ~0ULL ? -1 : (long long)w;
                s_first_cost = w == ~0ULL ? 0.0 : __ldcg(P.cost + w);
            }
            // the slot of round r+2: its last readers (round r-1) are behind this barrier, its next writers (round r+2)
            // are behind the NEXT barrier, which this thread reaches only after the store
            if (!multi && blockIdx.x == 0) ctrl->seq_slot[(round + 2) % 3] = ~0ULL;
        }
        __syncthreads();
        const long long istar = s_first;
        if (istar == -2) break; // a peer rank never reached the exchange: p2p_timeout is set, the host reports it
        const bool found = istar >= 0;
        if (found) { // gbest <- the improving particle's speculated position, into this CTA's copy
            if (multi) {
                for (int d = tid; d < P.n; d += blockDim.x) sG[d] = __ldcg(P.recv + 2 + d);
            } else {
                for (int d = tid; d < P.n; d += blockDim.x) sG[d] = __ldcg(xb + goff(P, d >> 2, istar) + (d & 3));
            }
        }
        if (tid == 0) {
            const long long hi = found ? istar : wend - 1;
            S.commit_lo = start;
            S.commit_hi = hi;
            S.commit_buf = parity ^ 1;
            if (found) { // :420-425
                S.ring[S.ring_i] = S.Gc;
                S.ring_i = (S.ring_i + 1) % EQS_STALL_WINDOW;
                S.Gc = s_first_cost;
                S.improvements += 1;
            }
            if (hi + 1 >= N) { // the pass is complete
                S.iter += 1;
                int stalled = 1; // stalled_too_long, :433-441
                for (int j = 0; j < EQS_STALL_WINDOW; ++j)
                    if (!(fabs(S.ring[j] - S.Gc) < P.tol)) stalled = 0;
                S.stalled = stalled;
                S.start = 0;
                S.parity = parity ^ 1;
                S.kest = 0.5 * S.kest + 0.5 * (double)(S.improvements - S.imp0);
                S.imp0 = S.improvements;
                const double w = sqrt(wscale * (double)N / fmax(S.kest, 0.02));
                long long wn = w < 9.0e18 ? (long long)w : wmax;
                wn = wn < wmin ? wmin : wn;
                S.window = wn > wmax ? wmax : wn;
            } else {
                S.start = hi + 1;
            }
            S.round = round + 1;
        }
        __syncthreads();
    }
    // every CTA holds the same final state; one of them stores it
    if (blockIdx.x == 0) {
        for (int d = tid; d < P.n; d += blockDim.x) P.G[d] = sG[d];
        if (tid == 0) {
            ctrl->Gc = S.Gc;
            for (int j = 0; j < EQS_STALL_WINDOW; ++j) ctrl->ring[j] = S.ring[j];
            ctrl->ring_i = S.ring_i;
            ctrl->iter = S.iter;
            ctrl->improvements = S.improvements;
            ctrl->seq_imp0 = S.imp0;
            ctrl->seq_kest = S.kest;
            ctrl->seq_start = S.start;
            ctrl->seq_window = S.window;
            ctrl->seq_commit_lo = 1; // nothing pending: the loop only ends after the last round's commit
            ctrl->seq_commit_hi = 0;
            ctrl->seq_commit_buf = S.commit_buf;
            ctrl->seq_parity = S.parity;
            ctrl->stalled = S.stalled;
            ctrl->seq_round = S.round;
        }
    }
}

// K8: small populations (nloc <= PSO_SMALL_MAX, one rank).  The reference's own shapes -- 100 particles x 50 passes
// (benchmarks/multi.rs:70-84), 1000 x 1000 (BASELINE config 1) -- fit on ONE SM, and a pass is a few microseconds, so
// kernel launches and host round trips are the whole cost.  One CTA therefore runs `iters` passes back to back: same
// device functions, same buffers, __syncthreads instead of kernel boundaries.  Sequential-exact mode runs its
// speculate-and-repair rounds (see pso_spec_kernel) inside the kernel too.
template <class Obj, bool SEQ>
__global__ void __launch_bounds__(PSO_SMALL_BLOCK) pso_small_kernel(const PsoDev P, long long iters)
{
    extern __shared__ double sG[];
    __shared__ MinLoc s_red[32];
    __shared__ MinLoc s_win;
    __shared__ long long s_start;
    PsoCtrl *c = P.ctrl;
    const int tid = threadIdx.x;
    for (long long pass = 0; pass < iters; ++pass) {
        if (c->stalled) break; // :395-397 (uniform: written before the last barrier of the previous pass)
        const uint32_t t = (uint32_t)c->iter;
        load_gbest(P, sG);
        if constexpr (!SEQ) {
            MinLoc best{dinf(), kNoIndex};
            for (long long li = tid; li < P.nloc; li += blockDim.x) {
                const unsigned f = P.pbflag[li];
                const bool stale = f & 1u, role = f & 2u;
                const double pbc_old = P.pbc[li];
                double *xb = role ? P.pb : P.x, *ob = role ? P.x : P.pb;
                const double cost = pso_move<Obj, false>(P, t, li, sG, xb, P.v, ob, stale ? ob : xb, P.v, !stale);
                pso_finish_particle(P, li, cost, pbc_old, f, best);
            }
            const MinLoc b = block_minloc(best, s_red);
            if (tid == 0) s_win = b;
            __threadfence_block();
            __syncthreads();
            write_record<false>(P, s_win.c, s_win.i, nullptr);
            __threadfence_block();
            __syncthreads();
            pso_apply_step(P, P.send, 1);
            __syncthreads();
        } else {
            const int nq = (P.n + 3) >> 2;
            if (tid == 0) s_start = 0;
            __syncthreads();
            for (;;) {
                const long long start = s_start;
                const double Gc = c->Gc;
                MinLoc first{0.0, kNoIndex};
                for (long long li = tid; li < P.nloc; li += blockDim.x) {
                    const long long gi = P.first + li;
                    if (gi < start) continue;
                    const double cost = pso_move<Obj, false>(P, t, li, sG, P.x, P.v, P.pb, P.x2, P.v2, true);
                    P.cost[li] = cost;
                    if (cost < P.pbc[li] && cost < Gc && gi < first.i) first.i = gi; // :417,:420 (nested)
                }
                const MinLoc b = block_minloc(first, s_red);
                if (tid == 0) s_win = b;
                __threadfence_block();
                __syncthreads();
                const long long istar = s_win.i;
                const bool found = istar != kNoIndex;
                const long long end = found ? istar : P.ntotal - 1;
                for (long long li = tid; li < P.nloc; li += blockDim.x) { // commit [start, end], :413-419
                    const long long gi = P.first + li;
                    if (gi < start || gi > end) continue;
                    const double cost = P.cost[li];
                    const bool imp = cost < P.pbc[li];
                    if (imp) P.pbc[li] = cost;
                    for (int q = 0; q < nq; ++q) {
                        const long long off = goff(P, q, li);
                        const D4 xv = ld256(P.x2 + off);
                        st256(P.x + off, xv);
                        st256(P.v + off, ld256(P.v2 + off));
                        if (imp) st256(P.pb + off, xv);
                    }
                }
                if (found) { // :420-425
                    const long long wl = istar - P.first;
                    for (int d = tid; d < P.n; d += blockDim.x) P.G[d] = P.x2[goff(P, d >> 2, wl) + (d & 3)];
                    if (tid == 0) {
                        ring_push(c, c->Gc);
                        c->Gc = P.cost[wl];
                        c->improvements += 1;
                        s_start = istar + 1;
                    }
                }
                const bool done = !found || istar == P.ntotal - 1;
                __threadfence_block();
                __syncthreads();
                if (done) break;
                load_gbest(P, sG);
            }
            if (tid == 0) {
                c->iter += 1;
                c->stalled = stall_test(c, P.tol);
            }
            __syncthreads();
        }
    }
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

template <class K> int persistent_grid(Ctx *ctx, K kernel, int block, size_t smem, long long units_per_block_iter,
                                       long long units, int &grid)
{
    int nb = 0;
    if (smem > kSmemOptIn) EQS_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    EQS_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, block, smem));
    if (nb < 1) return ctx->fail(EQS_ERR_EXTERNAL, "kernel does not fit on an SM (smem %zu)", smem);
    const long long need = std::max(1LL, (units + units_per_block_iter - 1) / units_per_block_iter);
    grid = (int)std::min<long long>(std::min<long long>((long long)nb * ctx->sm_count, need), 64LL * ctx->sm_count);
    return EQS_OK;
}

} // namespace

// ---- host driver ---------------------------------------------------------------------------------------------

PsoSolver::~PsoSolver()
{
    if (ctx_ && ctx_->device >= 0) cudaSetDevice(ctx_->device);
    if (arena_ || h_ctrl_) cudaStreamSynchronize(ctx_->stream); // nothing of this solver may still be in flight
    ctx_->release_arena(arena_, arena_bytes_);
    if (replay_buf_) cudaFree(replay_buf_);
    ctx_->release_pinned(h_ctrl_);
    if (ev0_) cudaEventDestroy(ev0_);
    if (ev1_) cudaEventDestroy(ev1_);
}

int validate_pso_params(const eqs_pso_params *p, std::string *why)
{
    auto bad = [&](int st, const char *m) {
        if (why) *why = m;
        return st;
    };
    if (!p) return bad(EQS_ERR_INCORRECT_INPUT, "null parameters");
    if (p->n <= 0) return bad(EQS_ERR_INCORRECT_INPUT, "dimension n must be positive");
    if (p->n > 12800) return bad(EQS_ERR_INCORRECT_INPUT, "dimension n out of range (1..12800: gbest tile in shared memory)");
    if (p->particle_count < 0 || p->iter_max < 0) return bad(EQS_ERR_INCORRECT_INPUT, "negative count");
    if (p->particle_count > (1LL << 40)) return bad(EQS_ERR_INCORRECT_INPUT, "particle count out of range (0..2^40)");
    if (p->objective_id < 0 || p->objective_id >= EQS_OBJ_COUNT) return bad(EQS_ERR_INCORRECT_INPUT, "unknown objective id");
    if ((p->objective_id == EQS_OBJ_THREE_CIRCLES || p->objective_id == EQS_OBJ_SINCOS) && p->n != 2)
        return bad(EQS_ERR_INCORRECT_INPUT, "this objective needs n = 2");
    if (p->mode != EQS_PSO_SEQUENTIAL_EXACT && p->mode != EQS_PSO_SYNCHRONOUS) return bad(EQS_ERR_INCORRECT_INPUT, "unknown mode");
    if (!p->lower || !p->upper) return bad(EQS_ERR_INCORRECT_INPUT, "null bounds");
    // Uniform::new_inclusive(low, high): rand's error strings, surfaced as SolverError::ExternalError (:129,:139)
    for (int d = 0; d < p->n; ++d) {
        const double lo = p->lower[d], hi = p->upper[d];
        if (!(lo <= hi)) return bad(EQS_ERR_EXTERNAL, "low > high (or equal if exclusive) in uniform distribution");
        if (!std::isfinite(lo) || !std::isfinite(hi) || !std::isfinite(hi - lo))
            return bad(EQS_ERR_EXTERNAL, "Non-finite range in uniform distribution");
    }
    return EQS_OK;
}

int PsoSolver::create(const eqs_pso_params *p)
{
    std::string why;
    const int st = validate_pso_params(p, &why);
    if (st != EQS_OK) return ctx_->fail(st, "%s", why.c_str());
    p_ = *p;
    const int n = p->n;
    lower_.assign(p->lower, p->lower + n);
    upper_.assign(p->upper, p->upper + n);
    p_.lower = lower_.data();
    p_.upper = upper_.data();
    int rank = ctx_->rank, world = ctx_->world;
    if (p->exchange == EQS_EXCHANGE_MANUAL) {
        rank = p->virtual_rank;
        world = p->virtual_world;
        if (world < 1 || rank < 0 || rank >= world) return ctx_->fail(EQS_ERR_INCORRECT_INPUT, "bad virtual rank/world");
        if (p->mode == EQS_PSO_SEQUENTIAL_EXACT && world > 1)
            return ctx_->fail(EQS_ERR_INCORRECT_INPUT, "sequential-exact mode needs exchange = AUTO");
    }
    manual_ = p->exchange == EQS_EXCHANGE_MANUAL && world > 1;
    EQS_CUDA(ctx_, cudaSetDevice(ctx_->device));
    int64_t first = 0, nloc = 0;
    eqs_shard_range(p->particle_count, world, rank, &first, &nloc);
    PsoDev &D = d_;
    D.n = n;
    D.world = world;
    D.rank = rank;
    // peer mailboxes (NVLink, CUDA IPC): the fused exchange of the synchronous step and of the sequential rounds
    char *mb = nullptr;
    char *const *peers = nullptr;
    const bool p2p = !manual_ && world > 1 && 2 * (2 + (long long)n) <= MB_LL_CAP && comm_p2p(ctx_, &mb, &peers);
    // Sequential-exact mode over several ranks: tile-cyclic distribution + device-driven rounds (pso_seqp_kernel); without
    // peer memory (or with EQS_PSO_SEQ_HOST=1) the host-driven whole-suffix rounds on contiguous blocks
    D.cyclic = p->mode == EQS_PSO_SEQUENTIAL_EXACT && p2p && !getenv("EQS_PSO_SEQ_HOST");
    D.tiles_per_rank = 0;
    if (D.cyclic) {
        long long cn = 0, per = 1;
        pso_cyclic_shard(p->particle_count, world, rank, &cn, &per);
        nloc = cn;
        D.tiles_per_rank = per;
        first = (int64_t)rank * PSO_GROUP; // global index of local particle 0 (informational: see local_indices)
    }
    D.first = first;
    D.nloc = nloc;
    D.ntotal = p->particle_count;
    D.ld = (long long)align_up((size_t)std::max<int64_t>(nloc, 1), 32); // rows start on 1 KB boundaries
    D.w = p->w;
    D.c1 = p->c1;
    D.c2 = p->c2;
    D.tol = p->tol;
    D.key = philox_key(p->seed);
    D.sync = p->mode == EQS_PSO_SYNCHRONOUS;
    D.replay = nullptr;
    D.replay_t0 = 0;
    const int nq = (n + 3) / 4;
    if ((size_t)nq * 4 * sizeof(double) > 100 * 1024) return ctx_->fail(EQS_ERR_INCORRECT_INPUT, "n too large for the shared-memory gbest tile");

    const bool seq = p->mode == EQS_PSO_SEQUENTIAL_EXACT;
    const size_t mat = (size_t)nq * 4 * D.ld * sizeof(double);
    const size_t vec = (size_t)D.ld * sizeof(double);
    const size_t R = (size_t)record_doubles(n) * sizeof(double);
    const long long ngroups = std::max<long long>(std::max<long long>((nloc + PSO_GROUP - 1) / PSO_GROUP, D.tiles_per_rank), 1);
    const int max_grid = 64 * std::max(ctx_->sm_count, 1);
    // one arena: plan the offsets, allocate once, then bind the pointers
    size_t off = 0;
    auto plan = [&](size_t bytes) {
        const size_t o = off;
        off += align_up(bytes, 256);
        return o;
    };
    const size_t o_x = plan(mat), o_v = plan(mat), o_pb = plan(mat), o_pbc = plan(vec), o_cost = plan(vec);
    const size_t o_flag = plan((size_t)D.ld);
    const size_t o_x2 = seq ? plan(mat) : 0, o_v2 = seq ? plan(mat) : 0;
    const size_t o_G = plan(n * sizeof(double)), o_lo = plan(n * sizeof(double)), o_hi = plan(n * sizeof(double));
    const size_t o_ctrl = plan(sizeof(PsoCtrl));
    const size_t o_pc = plan(max_grid * sizeof(double)), o_pi = plan(max_grid * sizeof(long long));
    const size_t o_gm = plan(ngroups * sizeof(double));
    const size_t o_gma = D.cyclic ? plan((size_t)world * D.tiles_per_rank * sizeof(double)) : 0;
    const size_t o_send = plan(R);
    const bool own_recv = !(world == 1 && !manual_);
    const size_t o_recv = own_recv ? plan(R * world) : o_send;
    EQS_CUDA(ctx_, ctx_->acquire_arena(off, &arena_, &arena_bytes_));
    char *base = (char *)arena_;
    D.x = (double *)(base + o_x);
    D.v = (double *)(base + o_v);
    D.pb = (double *)(base + o_pb);
    D.pbc = (double *)(base + o_pbc);
    D.cost = (double *)(base + o_cost);
    D.pbflag = (unsigned char *)(base + o_flag);
    D.x2 = seq ? (double *)(base + o_x2) : nullptr;
    D.v2 = seq ? (double *)(base + o_v2) : nullptr;
    D.G = (double *)(base + o_G);
    double *lo = (double *)(base + o_lo), *hi = (double *)(base + o_hi);
    D.lower = lo;
    D.upper = hi;
    D.ctrl = (PsoCtrl *)(base + o_ctrl);
    D.part_cost = (double *)(base + o_pc);
    D.part_idx = (long long *)(base + o_pi);
    D.group_min = (double *)(base + o_gm);
    D.gm_all = D.cyclic ? (double *)(base + o_gma) : nullptr;
    D.send = (double *)(base + o_send);
    D.recv = (double *)(base + o_recv);
    max_grid_ = max_grid;
    D.mailbox = nullptr;
    D.peers = nullptr;
    fused_exchange_ = false;
    if (p2p && (D.sync || D.cyclic)) {
        D.mailbox = mb;
        D.peers = peers;
        D.p2p_timeout_ns = comm_p2p_timeout_ns();
        fused_exchange_ = D.sync;
    }
    if (D.cyclic) { // tiles this rank does not have (ragged end of the cyclic deal) never hold a record low
        std::vector<double> inf((size_t)ngroups, HUGE_VAL);
        EQS_CUDA(ctx_, cudaMemcpyAsync(D.group_min, inf.data(), inf.size() * sizeof(double), cudaMemcpyHostToDevice, ctx_->stream));
        EQS_CUDA(ctx_, cudaStreamSynchronize(ctx_->stream));
    }
    EQS_CUDA(ctx_, cudaMemcpyAsync(lo, lower_.data(), n * sizeof(double), cudaMemcpyHostToDevice, ctx_->stream));
    EQS_CUDA(ctx_, cudaMemcpyAsync(hi, upper_.data(), n * sizeof(double), cudaMemcpyHostToDevice, ctx_->stream));
    EQS_CUDA(ctx_, cudaMemsetAsync(D.ctrl, 0, sizeof(PsoCtrl), ctx_->stream));
    EQS_CUDA(ctx_, cudaMemsetAsync(D.send, 0, R, ctx_->stream));
    static_assert(sizeof(PsoCtrl) <= Ctx::kPinnedBytes, "pinned block too small");
    EQS_CUDA(ctx_, ctx_->acquire_pinned((void **)&h_ctrl_));
    memset(h_ctrl_, 0, sizeof(PsoCtrl));
    EQS_CUDA(ctx_, cudaEventCreate(&ev0_));
    EQS_CUDA(ctx_, cudaEventCreate(&ev1_));
    // sequential mode on one rank beyond the single-CTA kernel's reach: device-driven windowed rounds
    // (EQS_PSO_SEQ_HOST=1 keeps the host-driven whole-suffix rounds, which ranks > 1 always use)
    seq_windowed_ = p->mode == EQS_PSO_SEQUENTIAL_EXACT &&
                    (d_.cyclic || (d_.world == 1 && !manual_ && d_.nloc > PSO_SMALL_MAX && !getenv("EQS_PSO_SEQ_HOST")));
    return EQS_OK;
}

int PsoSolver::upload_replay(const double *host, int64_t doubles)
{
    const size_t bytes = (size_t)doubles * sizeof(double);
    if (bytes > replay_cap_) {
        if (replay_buf_) EQS_CUDA(ctx_, cudaFree(replay_buf_));
        replay_buf_ = nullptr;
        replay_cap_ = 0; // a failed cudaMalloc below must not leave the old capacity behind
        EQS_CUDA(ctx_, cudaMalloc(&replay_buf_, bytes));
        replay_cap_ = bytes;
    }
    EQS_CUDA(ctx_, cudaMemcpyAsync(replay_buf_, host, bytes, cudaMemcpyHostToDevice, ctx_->stream));
    return EQS_OK;
}

int PsoSolver::exchange()
{
    if (manual_ || d_.world == 1) return EQS_OK;
    if (fused_exchange_ && in_sync_step_) return EQS_OK; // done by the step kernel itself over peer memory
    return comm_allgather_f64(ctx_, d_.send, d_.recv, record_doubles(d_.n), ctx_->stream);
}

int PsoSolver::fetch_ctrl()
{
    EQS_CUDA(ctx_, cudaMemcpyAsync(h_ctrl_, d_.ctrl, sizeof(PsoCtrl), cudaMemcpyDeviceToHost, ctx_->stream));
    EQS_CUDA(ctx_, cudaStreamSynchronize(ctx_->stream));
    if (h_ctrl_->p2p_timeout) return ctx_->fail(EQS_ERR_EXTERNAL, "a peer rank did not reach the gbest exchange (peer-memory mailbox timed out)");
    return EQS_OK;
}

int PsoSolver::launch_init_eval()
{
    const bool replay = p_.replay_init != nullptr;
    int rc = EQS_OK;
    if (replay) {
        EQS_TRY(upload_replay(p_.replay_init, p_.particle_count * 2 * (int64_t)p_.n));
        d_.replay = replay_buf_;
    } else {
        d_.replay = nullptr;
    }
    const long long ngroups = (d_.nloc + PSO_GROUP - 1) / PSO_GROUP;
    dispatch_objective(p_.objective_id, [&]<class Obj>() {
        pso_x0_kernel<Obj><<<1, 32, 0, ctx_->stream>>>(d_);
        launches_++;
        auto run = [&](auto kernel) {
            int grid = 1;
            rc = persistent_grid(ctx_, kernel, PSO_BLOCK, 0, 1, ngroups, grid);
            if (rc != EQS_OK) return;
            kernel<<<grid, PSO_BLOCK, 0, ctx_->stream>>>(d_);
            launches_++;
        };
        if (replay) run(pso_init_kernel<Obj, true>);
        else run(pso_init_kernel<Obj, false>);
    });
    EQS_TRY(rc);
    EQS_CUDA(ctx_, cudaGetLastError());
    d_.replay = nullptr;
    return EQS_OK;
}

int PsoSolver::launch_ring_scan()
{
    pso_ring_scan_kernel<<<1, 1024, 0, ctx_->stream>>>(d_);
    launches_++;
    EQS_CUDA(ctx_, cudaGetLastError());
    return EQS_OK;
}

// tile-cyclic distribution: the ordered ring scan walks the tiles of ALL ranks, so every rank needs every rank's minima
int PsoSolver::gather_tile_minima()
{
    if (!d_.cyclic) return EQS_OK;
    return comm_allgather_f64(ctx_, d_.group_min, d_.gm_all, d_.tiles_per_rank, ctx_->stream);
}

int PsoSolver::launch_init_apply()
{
    pso_init_apply_kernel<<<1, 128, 0, ctx_->stream>>>(d_);
    launches_++;
    EQS_CUDA(ctx_, cudaGetLastError());
    return EQS_OK;
}

int PsoSolver::init_local(const double *x0, int phase)
{
    EQS_CUDA(ctx_, cudaSetDevice(ctx_->device));
    if (phase == 0) {
        EQS_CUDA(ctx_, cudaMemcpyAsync(d_.G, x0, d_.n * sizeof(double), cudaMemcpyHostToDevice, ctx_->stream));
        return launch_init_eval();
    }
    return launch_ring_scan();
}

int PsoSolver::init_apply()
{
    EQS_TRY(launch_init_apply());
    initialised_ = true;
    return EQS_OK;
}

int PsoSolver::init(const double *x0)
{
    if (manual_) return ctx_->fail(EQS_ERR_INCORRECT_INPUT, "exchange = MANUAL: use the split-phase calls");
    EQS_TRY(init_local(x0, 0));
    EQS_TRY(exchange());
    EQS_TRY(gather_tile_minima());
    EQS_TRY(init_local(x0, 1));
    EQS_TRY(exchange());
    return init_apply();
}

int PsoSolver::launch_step()
{
    const bool replay = d_.replay != nullptr;
    const int inline_apply = (d_.world == 1 && !manual_) ? 1 : (fused_exchange_ ? 2 : 0);
    int rc = EQS_OK;
    const size_t smem = (size_t)((d_.n + 3) / 4) * 4 * sizeof(double);
    dispatch_objective(p_.objective_id, [&]<class Obj>() {
        auto run = [&](auto kernel) {
            if (grid_step_ == 0) {
                rc = persistent_grid(ctx_, kernel, PSO_BLOCK, smem, PSO_BLOCK, d_.nloc, grid_step_);
                if (rc != EQS_OK) return;
            }
            kernel<<<grid_step_, PSO_BLOCK, smem, ctx_->stream>>>(d_, inline_apply);
            launches_++;
        };
        if (replay) run(pso_step_kernel<Obj, true>);
        else run(pso_step_kernel<Obj, false>);
    });
    EQS_TRY(rc);
    EQS_CUDA(ctx_, cudaGetLastError());
    return EQS_OK;
}

int PsoSolver::launch_step_apply()
{
    if ((d_.world == 1 && !manual_) || fused_exchange_) return EQS_OK; // applied by the last block of the step kernel
    pso_step_apply_kernel<<<1, 128, 0, ctx_->stream>>>(d_);
    launches_++;
    EQS_CUDA(ctx_, cudaGetLastError());
    return EQS_OK;
}

int PsoSolver::sequential_pass()
{
    EQS_TRY(fetch_ctrl());
    if (h_ctrl_->stalled) return EQS_OK;
    const bool replay = d_.replay != nullptr;
    const size_t smem = (size_t)((d_.n + 3) / 4) * 4 * sizeof(double);
    long long start = 0;
    for (;;) {
        int rc = EQS_OK;
        dispatch_objective(p_.objective_id, [&]<class Obj>() {
            auto run = [&](auto kernel) {
                int grid = 1;
                rc = persistent_grid(ctx_, kernel, PSO_BLOCK, smem, PSO_BLOCK, d_.nloc, grid);
                if (rc != EQS_OK) return;
                kernel<<<grid, PSO_BLOCK, smem, ctx_->stream>>>(d_, start);
                launches_++;
            };
            if (replay) run(pso_spec_kernel<Obj, true>);
            else run(pso_spec_kernel<Obj, false>);
        });
        EQS_TRY(rc);
        EQS_CUDA(ctx_, cudaGetLastError());
        EQS_TRY(exchange());
        pso_spec_apply_kernel<<<1, 128, 0, ctx_->stream>>>(d_);
        launches_++;
        EQS_TRY(fetch_ctrl());
        const long long istar = h_ctrl_->spec_first;
        const long long end = istar >= 0 ? istar : d_.ntotal - 1;
        int grid = 1;
        EQS_TRY(persistent_grid(ctx_, pso_commit_kernel, PSO_BLOCK, 0, PSO_BLOCK, d_.nloc, grid));
        pso_commit_kernel<<<grid, PSO_BLOCK, 0, ctx_->stream>>>(d_, start, end);
        launches_++;
        EQS_CUDA(ctx_, cudaGetLastError());
        if (h_ctrl_->pass_done) break;
        start = istar + 1;
    }
    return EQS_OK;
}

// `iters` passes of the sequential mode as device-driven windowed rounds: one cooperative launch of the persistent
// kernel (pso_seqp_kernel) behind a one-thread kernel that arms the target pass count and the window policy.
int PsoSolver::sequential_windowed(int64_t iters)
{
    const bool replay = d_.replay != nullptr;
    const size_t smem = (size_t)((d_.n + 3) / 4) * 4 * sizeof(double);
    int rc = EQS_OK;
    const void *kernel_fn = nullptr;
    dispatch_objective(p_.objective_id, [&]<class Obj>() {
        auto pick = [&](auto kernel) {
            kernel_fn = (const void *)kernel;
            if (grid_seqw_ == 0) rc = persistent_grid(ctx_, kernel, PSO_BLOCK, smem, PSO_BLOCK, d_.nloc, grid_seqw_);
        };
        if (replay) pick(pso_seqp_kernel<Obj, true>);
        else pick(pso_seqp_kernel<Obj, false>);
    });
    EQS_TRY(rc);
    // a round costs about 4 us beyond its particles (grid barrier, ramp-up and tail of the wave; a launch per round cost
    // 11 -- swept on C2: 1 / 2.5 / 4 / 8 us give 0.447 / 0.402 / 0.386 / 0.416 ms per pass, profiles/r02_sweep.md) and a
    // particle (40n + 16) B of HBM traffic at ~5.7 TB/s
    // Several ranks: the window is in GLOBAL particles and every rank works on 1 / world of it, a round also pays one
    // NVLink round trip of the peer mailboxes (EQS_PSO_SEQ_ROUND_US, swept on 2 GPUs: profiles/r02_sweep.md section 4).
    // The window policy must be the same on every rank: it only uses quantities all ranks agree on.
    const int world = d_.world;
    const long long wave = (long long)ctx_->sm_count * 2 * PSO_BLOCK * world;
    long long wmin = std::min<long long>(4096LL * world, wave), wmax = std::max<long long>(d_.ntotal, wmin), w0 = std::min(wave, wmax);
    double round_us = world > 1 ? 6.0 : 4.0;
    if (const char *e = getenv("EQS_PSO_SEQ_ROUND_US")) round_us = std::max(0.1, atof(e));
    const double wscale = 2.0 * round_us * 5.7e6 * world / (40.0 * d_.n + 16.0);
    if (const char *e = getenv("EQS_PSO_SEQ_WINDOW")) wmin = wmax = w0 = std::max<long long>(1, atoll(e));
    pso_seq_arm_kernel<<<1, 1, 0, ctx_->stream>>>(d_, (long long)iters, wmin, wmax, w0, wscale);
    launches_++;
    void *args[] = {(void *)&d_};
    EQS_CUDA(ctx_, cudaLaunchCooperativeKernel(kernel_fn, dim3(grid_seqw_), dim3(PSO_BLOCK), args, smem, ctx_->stream));
    launches_++;
    return EQS_OK;
}

int PsoSolver::step_local(const double *replay_step)
{
    if (!initialised_) return ctx_->fail(EQS_ERR_INCORRECT_INPUT, "the population has not been initialised (eqs_pso_init)");
    EQS_CUDA(ctx_, cudaSetDevice(ctx_->device));
    if (p_.mode != EQS_PSO_SYNCHRONOUS) return ctx_->fail(EQS_ERR_INCORRECT_INPUT, "split-phase stepping is synchronous-mode only");
    if (replay_step) {
        EQS_TRY(fetch_ctrl());
        EQS_TRY(upload_replay(replay_step, p_.particle_count * 2 * (int64_t)p_.n));
        d_.replay = replay_buf_;
        d_.replay_t0 = h_ctrl_->iter;
    } else {
        d_.replay = nullptr;
    }
    grid_step_ = 0;
    return launch_step();
}

int PsoSolver::step_apply() { return launch_step_apply(); }

int PsoSolver::step(int64_t iters, const double *replay_step)
{
    if (!initialised_) return ctx_->fail(EQS_ERR_INCORRECT_INPUT, "the population has not been initialised (eqs_pso_init)");
    EQS_CUDA(ctx_, cudaSetDevice(ctx_->device));
    if (manual_) return ctx_->fail(EQS_ERR_INCORRECT_INPUT, "exchange = MANUAL: use the split-phase calls");
    if (replay_step) {
        EQS_TRY(fetch_ctrl());
        EQS_TRY(upload_replay(replay_step, iters * p_.particle_count * 2 * (int64_t)p_.n));
        d_.replay = replay_buf_;
        d_.replay_t0 = h_ctrl_->iter;
    } else {
        d_.replay = nullptr;
    }
    grid_step_ = 0;
    const int64_t l0 = launches_;
    EQS_CUDA(ctx_, cudaEventRecord(ev0_, ctx_->stream));
    const bool small = d_.world == 1 && !manual_ && !replay_step && d_.nloc >= 1 && d_.nloc <= PSO_SMALL_MAX &&
                       !getenv("EQS_PSO_NO_SMALL");
    if (small && iters > 0) {
        // K8: the whole batch of passes in one single-CTA launch
        const size_t smem = (size_t)((d_.n + 3) / 4) * 4 * sizeof(double);
        const int block = (int)std::min<long long>(PSO_SMALL_BLOCK, (d_.nloc + 31) / 32 * 32);
        const bool seq = p_.mode == EQS_PSO_SEQUENTIAL_EXACT;
        int rc = EQS_OK;
        dispatch_objective(p_.objective_id, [&]<class Obj>() {
            auto run = [&](auto kernel) {
                if (smem > kSmemOptIn &&
                    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
                    rc = ctx_->fail(EQS_ERR_EXTERNAL, "cannot reserve %zu bytes of shared memory", smem);
                if (rc != EQS_OK) return;
                kernel<<<1, block, smem, ctx_->stream>>>(d_, (long long)iters);
                launches_++;
            };
            if (seq) run(pso_small_kernel<Obj, true>);
            else run(pso_small_kernel<Obj, false>);
        });
        EQS_TRY(rc);
        EQS_CUDA(ctx_, cudaGetLastError());
    }
    if (!small && seq_windowed_ && iters > 0) {
        EQS_TRY(sequential_windowed(iters));
    }
    for (int64_t it = 0; it < ((small || seq_windowed_) ? 0 : iters); ++it) {
        if (p_.mode == EQS_PSO_SYNCHRONOUS) {
            in_sync_step_ = true;
            EQS_TRY(launch_step());
            EQS_TRY(exchange());
            EQS_TRY(launch_step_apply());
            in_sync_step_ = false;
        } else {
            EQS_TRY(sequential_pass());
        }
    }
    EQS_CUDA(ctx_, cudaEventRecord(ev1_, ctx_->stream));
    last_launches_ = launches_ - l0;
    return EQS_OK;
}

int PsoSolver::sync()
{
    EQS_CUDA(ctx_, cudaSetDevice(ctx_->device));
    EQS_CUDA(ctx_, cudaStreamSynchronize(ctx_->stream));
    return EQS_OK;
}

int PsoSolver::read_gbest(double *g, double *cost, int64_t *iters, int32_t *stalled, int64_t *improvements)
{
    if (!initialised_) return ctx_->fail(EQS_ERR_INCORRECT_INPUT, "the population has not been initialised (eqs_pso_init)");
    EQS_CUDA(ctx_, cudaSetDevice(ctx_->device));
    if (g) EQS_CUDA(ctx_, cudaMemcpyAsync(g, d_.G, d_.n * sizeof(double), cudaMemcpyDeviceToHost, ctx_->stream));
    EQS_TRY(fetch_ctrl());
    if (cost) *cost = h_ctrl_->Gc;
    if (iters) *iters = h_ctrl_->iter;
    if (stalled) *stalled = h_ctrl_->stalled;
    if (improvements) *improvements = h_ctrl_->improvements;
    return EQS_OK;
}

int PsoSolver::read_state(double *x, double *v, double *pb, double *pbc)
{
    if (!initialised_) return ctx_->fail(EQS_ERR_INCORRECT_INPUT, "the population has not been initialised (eqs_pso_init)");
    EQS_CUDA(ctx_, cudaSetDevice(ctx_->device));
    const int n = d_.n, nq = (n + 3) / 4;
    const long long L = d_.nloc, ld = d_.ld;
    const size_t m = (size_t)nq * 4 * ld;
    auto pull = [&](const void *dev, void *host, size_t bytes) -> int {
        EQS_CUDA(ctx_, cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx_->stream));
        EQS_CUDA(ctx_, cudaStreamSynchronize(ctx_->stream));
        return EQS_OK;
    };
    auto at = [&](const std::vector<double> &buf, long long i, int d) { return buf[((size_t)(d >> 2) * ld + i) * 4 + (d & 3)]; };
    // device layout -> the reference's particle-major Vec<Particle>, resolving the role / stale bits of synchronous mode
    std::vector<unsigned char> flag((size_t)std::max<long long>(L, 1), 0);
    if (d_.sync && L > 0) EQS_TRY(pull(d_.pbflag, flag.data(), (size_t)L));
    // windowed sequential mode: the committed (x, v) may live in (x2, v2)
    bool swapped = false;
    if (seq_windowed_) {
        EQS_TRY(fetch_ctrl());
        swapped = h_ctrl_->seq_parity != 0;
    }
    if (x || pb) {
        std::vector<double> A(m), B(m);
        EQS_TRY(pull(swapped ? d_.x2 : d_.x, A.data(), m * sizeof(double)));
        EQS_TRY(pull(d_.pb, B.data(), m * sizeof(double)));
        for (long long i = 0; i < L; ++i) {
            const bool stale = flag[i] & 1, role = flag[i] & 2;
            const std::vector<double> &X = role ? B : A, &O = role ? A : B;
            for (int d = 0; d < n; ++d) {
                if (x) x[i * n + d] = at(X, i, d);
                if (pb) pb[i * n + d] = stale ? at(X, i, d) : at(O, i, d);
            }
        }
    }
    if (v) {
        std::vector<double> V(m);
        EQS_TRY(pull(swapped ? d_.v2 : d_.v, V.data(), m * sizeof(double)));
        for (long long i = 0; i < L; ++i)
            for (int d = 0; d < n; ++d) v[i * n + d] = at(V, i, d);
    }
    if (pbc && L > 0) EQS_TRY(pull(d_.pbc, pbc, (size_t)L * sizeof(double)));
    return EQS_OK;
}

int PsoSolver::read_ring(double *ring, int64_t *ring_i)
{
    EQS_TRY(fetch_ctrl());
    if (ring) memcpy(ring, h_ctrl_->ring, sizeof(h_ctrl_->ring));
    if (ring_i) *ring_i = h_ctrl_->ring_i;
    return EQS_OK;
}

// global index of every local particle, in local order (what read_state's rows are)
int PsoSolver::local_indices(int64_t *gidx)
{
    for (long long li = 0; li < d_.nloc; ++li) gidx[li] = pso_gidx(d_, li);
    return EQS_OK;
}

int PsoSolver::get_record(double *rec)
{
    EQS_CUDA(ctx_, cudaSetDevice(ctx_->device));
    EQS_CUDA(ctx_, cudaMemcpyAsync(rec, d_.send, record_doubles(d_.n) * sizeof(double), cudaMemcpyDeviceToHost, ctx_->stream));
    EQS_CUDA(ctx_, cudaStreamSynchronize(ctx_->stream));
    return EQS_OK;
}

int PsoSolver::set_records(const double *recs)
{
    EQS_CUDA(ctx_, cudaSetDevice(ctx_->device));
    EQS_CUDA(ctx_, cudaMemcpyAsync(d_.recv, recs, d_.world * record_doubles(d_.n) * sizeof(double), cudaMemcpyHostToDevice,
                                   ctx_->stream));
    EQS_CUDA(ctx_, cudaStreamSynchronize(ctx_->stream));
    return EQS_OK;
}

// ParticleSwarm::solve, particle_swarm.rs:356-431
int PsoSolver::solve(const double *x0, double *out_x, eqs_report *rep)
{
    Event e0, e1, e2;
    EQS_CUDA(ctx_, e0.create());
    EQS_CUDA(ctx_, e1.create());
    EQS_CUDA(ctx_, e2.create());
    const int64_t l0 = launches_;
    EQS_CUDA(ctx_, cudaEventRecord(e0, ctx_->stream));
    int st = init(x0);
    if (st == EQS_OK) st = cudaEventRecord(e1, ctx_->stream) == cudaSuccess ? EQS_OK : EQS_ERR_EXTERNAL;
    // the stall test lives on the device; the host only looks every kBatch passes to stop launching no-ops
    const int64_t kBatch = 64;
    int64_t left = p_.iter_max;
    while (st == EQS_OK && left > 0) {
        const int64_t b = std::min(kBatch, left);
        st = step(b, nullptr);
        left -= b;
        if (st == EQS_OK && left > 0) {
            st = fetch_ctrl();
            if (st == EQS_OK && h_ctrl_->stalled) break;
        }
    }
    if (st == EQS_OK) {
        cudaEventRecord(e2, ctx_->stream);
        st = read_gbest(out_x, nullptr, nullptr, nullptr, nullptr); // Ok(global_best_position), :430
    }
    if (st == EQS_OK && rep) {
        float ms_init = 0.f, ms_loop = 0.f;
        cudaEventElapsedTime(&ms_init, e0, e1);
        cudaEventElapsedTime(&ms_loop, e1, e2);
        rep->iters_run = h_ctrl_->iter;
        rep->evals = 1 + p_.particle_count * (1 + h_ctrl_->iter);
        rep->best_cost = h_ctrl_->Gc;
        rep->stalled = h_ctrl_->stalled;
        rep->world = d_.world;
        rep->kernel_launches = launches_ - l0;
        rep->init_ms = ms_init;
        rep->loop_ms = ms_loop;
    }
    return st;
}

} // namespace eqs
