// ce.cu -- CrossEntropy population step on sm_100a (f64; FP64-pipe bound: the sample matrix never touches HBM).
//
// Replaces the body of CrossEntropy::solve, reference src/solvers/global_optimisers/cross_entropy.rs:245-306:
//   ce_sample_kernel   :256-274  x = mu + sigma z (counter-based Philox + Box-Muller in registers), cost only    (K4)
//                                + min / max cost key and NaN count of the pass
//   ce_level_kernel    :276      the full sort is replaced by a radix SELECT of the k-th cost key over the key RANGE (K5)
//                                of the pass: 11 bits per level, level A over all keys, level B over all keys while it
//                                compacts the chosen bucket into a candidate list, further levels over that list
//   ce_count / compact :284,291  elite set = keys below the k-th key + lowest-index ties, in sample order
//   ce_elite_gen       :284,291  elites are REGENERATED from their counters into a k x n matrix
//   ce_rowsum          :281-297  two-pass mean / Bessel-corrected sigma per dimension, fixed summation order        (K6)
//                      :255,299-302  mu, sigma <- new; iter += 1; loop condition for the next pass
// A pass is EIGHT launches.  Every step that needs all ranks' data (key range, histogram of a level, moment sums) is done
// by the LAST CTA of the kernel that produced this rank's share: it all-gathers the shares over the peer-memory
// mailboxes (common.cuh: ll_allgather_u64; nothing to gather on one rank) and applies them, so no collective and no
// extra kernel sits between the stages.  The same device functions also run as separate PHASES with the exchange done
// outside (virtual ranks driven by a test harness, NCCL when peer mapping is unavailable).
#include <algorithm>
#include <cmath>
#include <new>

#define EQS_COS_TABLE 1 // issue-bound sample kernels: constant-operand cos (devmath.cuh)
#ifndef EQS_CE_COS_DEFER // 1: the objective's range test once per sample instead of once per trip (devmath.cuh: DEFER)
#define EQS_CE_COS_DEFER 1
#endif
#define EQS_DEVMATH_TABLES // Box-Muller's logarithm reads its tables from shared memory (devmath.cuh: log_unit_tab)
#include "common.cuh"
#include "objectives.cuh"
#include "philox.cuh"

namespace eqs {

namespace {

constexpr int CE_BLOCK = 256;
// tuning knobs of ce_sample_kernel (profiles/r01_sweep.md): coordinates per trip of the dimension loop, min CTAs per SM
#ifndef EQS_CE_DIMS
#define EQS_CE_DIMS 4
#endif
#ifndef EQS_CE_MINBLOCKS
#define EQS_CE_MINBLOCKS 2
#endif
constexpr int CE_BINS = 2048;  // 11 bits per level
constexpr int CE_LEVELS = 6;   // ceil(64 / 11): what a 64-bit key range can need
constexpr int CE_SEL_BLOCK = 512; // threads of the select kernels: few, fat CTAs keep the histogram flush (global atomics) short
constexpr int CE_TILE = 2048;  // samples per block in count / compact (256 threads x 8, blocked)
constexpr int CE_GEN_BLOCK = 64; // elite regeneration: k threads in all, small CTAs spread them over the SMs
constexpr int CE_CHUNK = 4096; // elites per block in the row sums
constexpr int CE_PHASES = 10;
constexpr int CE_SMALL_MAX = 1024; // samples the single-CTA whole-solve kernel takes (one per thread)

// how the stages of a pass are joined
enum { CE_PHASED = 0,      // exchange outside the kernels (virtual ranks, NCCL): one phase per exchange
       CE_FUSED_LOCAL = 1, // one rank: the last CTA applies its own data
       CE_FUSED_P2P = 2 }; // the last CTA all-gathers over the peer mailboxes, then applies

struct CeCtrl {
    unsigned long long base;      // lowest key of the bucket that holds the k-th key (level by level); at the end the k-th key
    unsigned long long bucket_hi; // highest key of that bucket
    unsigned long long n_cand;    // keys in the candidate list (level B)
    long long k_rem;              // rank of the k-th key inside the current bucket (1-based)
    long long tie_take;           // samples with key == k-th key this rank contributes (lowest indices first)
    long long n_elite_local;
    long long iter;               // the reference's `iter`, starts at 1 (:253)
    double sigma_inf;             // max |sigma| (:255)
    int shift;                    // bin of the current level = (key - base) >> shift
    int sel_done;                 // the k-th key is known
    int active;                   // loop condition :255 for the NEXT pass
    int nan_cost;                 // a NaN cost was seen on SOME rank: the reference panics in the sort (:276)
    int bad_sigma;                // the loop would go on with a non-finite sigma: the reference panics in Normal::new (:259)
    int p2p_timeout;              // a peer never showed up in a fused exchange
    unsigned int ticket;          // last-CTA ticket of the fused kernels (left at 0 by each of them)
    unsigned int sample_ticket;   // the same for the sample kernel's work counter (every mode)
    unsigned long long work_next; // sample kernel: first local sample nobody has taken yet (left at 0 by the kernel)
};

struct CeDev {
    double *mu, *sigma, *mu_new; // [n]
    double *cost;                // [ld]
    double *xs;                  // [n][ld] only for strided objectives
    unsigned long long *hsend, *hrecv; // [BINS], [world][BINS]; words 0..2 double as (min key, max key, NaN count) of the pass
    double *msend, *mrecv;       // [n], [world][n]
    unsigned long long *cand;    // [ld] keys of the bucket chosen at level A, unordered
    int *blk_less, *blk_tie;     // [ntiles]
    long long *blk_tie_off, *blk_el_off;
    long long *elite_idx;        // [kcap] local sample indices, ascending
    double *Ex;                  // [n][kld] regenerated elites
    double *part;                // [n][nchunks]
    CeCtrl *ctrl;
    char *mailbox;               // peer-memory exchange (common.cuh), or nullptr
    char *const *peers;
    long long p2p_timeout_ns;
    const double *replay;
    long long replay_t0;
    long long ld, nloc, first, ntotal, k, kcap, kld, ntiles;
    long long iter_max;
    int nchunks;
    int n, world, rank, divisor_mode;
    double tol;
    PhiloxKey key;
};

// order-preserving f64 -> u64: a < b  <=>  key(a) < key(b); NaN sorts last
__device__ __forceinline__ unsigned long long cost_key(double c)
{
    const long long b = __double_as_longlong(c);
    const unsigned long long u = (unsigned long long)b;
    return b < 0 ? ~u : (u | 0x8000000000000000ULL);
}

// the coordinates of sample gi in dimension order: x_d = mu_d + sigma_d z_d  (Normal::sample, :265-271)
template <bool REPLAY, class F>
__device__ __forceinline__ void ce_sample_coords(const CeDev &P, uint32_t t, long long gi, const double *smu,
                                                 const double *ssig, F &&f)
{
    const int n = P.n;
    if constexpr (REPLAY) {
        const double *z = P.replay + (((long long)t - P.replay_t0) * P.ntotal + gi) * n;
        for (int d = 0; d < n; ++d) f(smu[d] + ssig[d] * z[d], d);
    } else {
        for (int p = 0; 2 * p < n; ++p) {
            double z0, z1;
            normal_pair(philox_draw(P.key, STREAM_CE, t, gi, (uint32_t)p), z0, z1);
            f(smu[2 * p] + ssig[2 * p] * z0, 2 * p);
            if (2 * p + 1 < n) f(smu[2 * p + 1] + ssig[2 * p + 1] * z1, 2 * p + 1);
        }
    }
}

// the same coordinates fed to a streaming objective, four (two Philox blocks) at a time where the functor has a
// feed_vec: its per-coordinate operations and their order are unchanged (objectives.cuh)
template <class Obj, bool REPLAY>
__device__ __forceinline__ double ce_sample_cost(const CeDev &P, uint32_t t, long long gi, const double *smu, const double *ssig)
{
    const int n = P.n;
    ObjAcc acc;
    Obj::begin(acc, n);
    if constexpr (REPLAY) {
        const double *z = P.replay + (((long long)t - P.replay_t0) * P.ntotal + gi) * n;
        for (int d = 0; d < n; ++d) Obj::feed(acc, smu[d] + ssig[d] * z[d], d);
    } else {
        int d = 0;
        unsigned hi_max = 0; // the objective's deferred range test (devmath.cuh: eqs_cos_vec DEFER), looked at once per sample
        constexpr int V = EQS_CE_DIMS; // coordinates per trip: V/2 Philox blocks and Box-Muller pairs in flight
        for (; d + V <= n; d += V) {
            double z[V], x[V];
#pragma unroll
            for (int j = 0; j < V; j += 2)
                normal_pair(philox_draw(P.key, STREAM_CE, t, gi, (uint32_t)((d + j) >> 1)), z[j], z[j + 1]);
#pragma unroll
            for (int j = 0; j < V; ++j) x[j] = smu[d + j] + ssig[d + j] * z[j];
            if constexpr (EQS_CE_COS_DEFER) objective_feed_vec_defer<Obj, V>(acc, x, d, hi_max);
            else objective_feed_vec<Obj, V>(acc, x, d);
        }
        for (; d < n; d += 2) {
            double z0, z1;
            normal_pair(philox_draw(P.key, STREAM_CE, t, gi, (uint32_t)(d >> 1)), z0, z1);
            Obj::feed(acc, smu[d] + ssig[d] * z0, d);
            if (d + 1 < n) Obj::feed(acc, smu[d + 1] + ssig[d + 1] * z1, d + 1);
        }
        if (EQS_CE_COS_DEFER && hi_max >= kCosHiLimit) {
            // some coordinate was outside the fast cosine's range (never on sane inputs): the sums hold garbage.  The sample
            // is a pure function of its counters: draw it again and feed it through the scalar path, whose cosine is total.
            Obj::begin(acc, n);
            ce_sample_coords<false>(P, t, gi, smu, ssig, [&](double x, int dd) { Obj::feed(acc, x, dd); });
        }
    }
    return Obj::end(acc, n);
}

__device__ __forceinline__ void load_mu_sigma(const CeDev &P, double *smu, double *ssig)
{
    for (int d = threadIdx.x; d < P.n; d += blockDim.x) {
        smu[d] = P.mu[d];
        ssig[d] = P.sigma[d];
    }
    __syncthreads();
}

// ---- joining the stages ------------------------------------------------------------------------------------------

// true in exactly one CTA of the grid, the last one to get here (all of its threads); everything the other CTAs wrote
// or added before their call is visible to it.  Leaves the ticket at 0 for the next kernel.
__device__ __forceinline__ bool ce_last_block(CeCtrl *c, unsigned total_blocks)
{
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(&c->ticket, 1u);
        s_last = (t == total_blocks - 1u);
        if (s_last) c->ticket = 0u;
    }
    __syncthreads();
    if (s_last) __threadfence();
    return s_last != 0;
}

// every rank's 8-byte words of one exchange, wherever the exchange left them: word i of rank r at p[r * stride + i]
struct Gathered {
    const unsigned long long *p;
    long long stride;
    int world;
    // L2 loads (the words were written by other CTAs, by this CTA's other threads, or by the NCCL kernel before this one):
    // they overlap, where volatile loads complete one after the other (32 of them per thread on 8 ranks)
    __device__ __forceinline__ unsigned long long operator()(int r, long long i) const { return __ldcg(p + r * stride + i); }
    __device__ __forceinline__ double f64(int r, long long i) const { return __longlong_as_double((long long)(*this)(r, i)); }
};

// one CTA, fused modes: `words` words at `send` (global memory, complete) from every rank
__device__ __forceinline__ Gathered ce_gather(const CeDev &P, int fused, const unsigned long long *send, long long words)
{
    if (fused == CE_FUSED_P2P) { // low-latency protocol (common.cuh: ll_*), into the receive buffers of the phased mode
        const bool hist = send == P.hsend;
        unsigned long long *recv = hist ? P.hrecv : reinterpret_cast<unsigned long long *>(P.mrecv);
        const long long stride = hist ? CE_BINS : P.n;
        if (!ll_allgather_u64(P.mailbox, P.peers, P.rank, P.world, send, words, recv, stride, P.p2p_timeout_ns))
            P.ctrl->p2p_timeout = 1;
        __syncthreads();
        return Gathered{recv, stride, P.world};
    }
    return Gathered{send, 0, 1};
}

// ---- the select, level by level ----------------------------------------------------------------------------------

// start of the select: key range of the pass over all ranks -> first level; a NaN cost on ANY rank stops the pass on
// EVERY rank before mu, sigma or iter change (the reference panics in the sort, :276).  One CTA.
__device__ __forceinline__ void ce_select_begin(const CeDev &P, Gathered g)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        CeCtrl *c = P.ctrl;
        unsigned long long lo = ~0ULL, hi = 0ULL, nan = 0ULL;
        for (int r = 0; r < g.world; ++r) {
            const unsigned long long l = g(r, 0), h = g(r, 1);
            lo = l < lo ? l : lo;
            hi = h > hi ? h : hi;
            nan += g(r, 2);
        }
        if (nan) {
            c->nan_cost = 1;
            c->active = 0;
        }
        if (c->p2p_timeout) c->active = 0;
        const unsigned long long range = hi >= lo ? hi - lo : 0ULL;
        const int bits = range ? 64 - __clzll((long long)range) : 0;
        c->base = lo;
        c->bucket_hi = hi >= lo ? hi : lo;
        c->shift = bits > 11 ? bits - 11 : 0;
        c->k_rem = P.k;
        c->sel_done = 0;
        c->n_cand = 0ULL;
        P.hsend[0] = P.hsend[1] = P.hsend[2] = 0ULL; // the histogram of level A accumulates here
    }
    __syncthreads();
}

// pick the bin that holds the k-th key from all ranks' histograms of the current level and descend into it.  The select
// ends (a) after the level with shift 0 -- the bucket is ONE key: split its ties over the ranks, lowest sample indices
// first -- or (b) as soon as the k-th key is the LAST key of the chosen bin: then every key up to the bin's upper end is
// elite and nothing has to be resolved inside it (the usual way out: two levels leave a handful of keys per bin).
// One CTA of CE_SEL_BLOCK threads, CE_BINS / CE_SEL_BLOCK bins per thread.
__device__ __forceinline__ void ce_pick(const CeDev &P, Gathered g)
{
    __shared__ long long s_scan[33];
    __shared__ int s_bin;
    __shared__ long long s_excl, s_cnt;
    CeCtrl *c = P.ctrl;
    const int tid = threadIdx.x;
    constexpr int PER = CE_BINS / CE_SEL_BLOCK;
    long long h[PER], tot = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        long long v = 0;
        for (int r = 0; r < g.world; ++r) v += (long long)g(r, tid * PER + j);
        h[j] = v;
        tot += v;
    }
    const long long k_rem = ((const volatile CeCtrl *)c)->k_rem;
    if (tid == 0) {
        s_bin = 0;
        s_excl = 0;
        s_cnt = 0;
    }
    long long total;
    long long excl = block_exclusive_scan(tot, 0LL, OpAddLL(), s_scan, total);
    if (excl < k_rem && k_rem <= excl + tot) {
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            if (excl < k_rem && k_rem <= excl + h[j]) {
                s_bin = tid * PER + j;
                s_excl = excl;
                s_cnt = h[j];
            }
            excl += h[j];
        }
    }
    __syncthreads();
    if (tid == 0) {
        const int shift = c->shift;
        const unsigned long long base = c->base + ((unsigned long long)s_bin << shift);
        unsigned long long top = base + ((1ULL << shift) - 1ULL);
        top = top < c->bucket_hi ? top : c->bucket_hi;
        const long long rem = k_rem - s_excl;
        c->k_rem = rem;
        if (shift == 0) { // rem = how many of the samples with key == k-th key are elite, lowest index first
            long long before = 0;
            for (int r = 0; r < P.rank && r < g.world; ++r) before += (long long)g(r, s_bin);
            const long long mine = (long long)g(g.world == 1 ? 0 : P.rank, s_bin);
            long long take = rem - before;
            take = take < 0 ? 0 : (take > mine ? mine : take);
            c->base = base;
            c->bucket_hi = base;
            c->tie_take = take;
            c->sel_done = 1;
        } else if (rem == s_cnt) { // the whole bin is elite: threshold = its upper end, every key equal to that is taken
            c->base = top;
            c->bucket_hi = top;
            c->tie_take = 0x7fffffffffffffffLL;
            c->sel_done = 1;
        } else {
            c->base = base;
            c->bucket_hi = top;
            c->shift = shift > 11 ? shift - 11 : 0;
        }
        if (c->p2p_timeout) c->active = 0;
    }
    __syncthreads();
    for (int b = tid; b < CE_BINS; b += blockDim.x) P.hsend[b] = 0ULL; // ready for the next level
    __threadfence_block();
    __syncthreads();
}

// histogram of the current level over the candidate list into P.hsend (plain stores: one CTA owns it)
__device__ __forceinline__ void ce_cand_hist(const CeDev &P, unsigned int *sh)
{
    const volatile CeCtrl *c = P.ctrl;
    const unsigned long long base = c->base, bhi = c->bucket_hi, M = c->n_cand;
    const int shift = c->shift;
    for (int b = threadIdx.x; b < CE_BINS; b += blockDim.x) sh[b] = 0u;
    __syncthreads();
    for (unsigned long long j = threadIdx.x; j < M; j += blockDim.x) {
        const unsigned long long key = __ldcg(P.cand + j);
        if (key >= base && key <= bhi) {
            EQS_CHECK(((key - base) >> shift) < CE_BINS);
            atomicAdd(&sh[(key - base) >> shift], 1u);
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < CE_BINS; b += blockDim.x) P.hsend[b] = (unsigned long long)sh[b];
    __threadfence();
    __syncthreads();
}

// K4: one thread per sample; nothing but the 8-byte cost goes to HBM.  Also the pass's key range and NaN count; in the
// fused modes the last CTA opens the select.
template <class Obj, bool REPLAY> __global__ void __launch_bounds__(CE_BLOCK, EQS_CE_MINBLOCKS) ce_sample_kernel(const CeDev P, int fused)
{
    extern __shared__ double s_ms[];
    __shared__ unsigned long long s_lo[CE_BLOCK / 32], s_hi[CE_BLOCK / 32];
    __shared__ unsigned int s_nan[CE_BLOCK / 32];
    CeCtrl *c = P.ctrl;
    if (!c->active) return;
    const uint32_t t = (uint32_t)(c->iter - 1);
    double *smu = s_ms, *ssig = s_ms + P.n;
    if constexpr (!REPLAY) devmath_tables_load();
    load_mu_sigma(P, smu, ssig);
    unsigned long long klo = ~0ULL, khi = 0ULL;
    unsigned int nan = 0u;
    // Work is handed out dynamically, 32 samples (one per lane) at a time.  With a fixed grid-stride assignment the warps of
    // a scheduler do not advance at the same rate in this issue-bound kernel: the favoured ones finish early and the tail
    // runs with 1-2 warps per scheduler (ncu: 13.3 of 16 warps active on average, profiles/r02y_ncu_ce_sample_c3.csv).
    // A sample's cost depends on its index only, so who computes it changes nothing.  The next chunk is requested before
    // the current one is computed: the atomic's latency is never waited for.
    const int lane = threadIdx.x & 31;
    unsigned long long next = 0;
    if (lane == 0) next = atomicAdd(&c->work_next, 32ULL);
    for (;;) {
        const long long base = (long long)__shfl_sync(0xffffffffu, next, 0);
        if (base >= P.nloc) break;
        if (lane == 0) next = atomicAdd(&c->work_next, 32ULL);
        const long long li = base + lane;
        if (li >= P.nloc) continue;
        const long long gi = P.first + li;
        double cost;
        if constexpr (Obj::streaming) {
            cost = ce_sample_cost<Obj, REPLAY>(P, t, gi, smu, ssig);
        } else {
            ce_sample_coords<REPLAY>(P, t, gi, smu, ssig, [&](double x, int d) { P.xs[(long long)d * P.ld + li] = x; });
            cost = Obj::eval(StridedView{P.xs + li, P.ld}, P.n);
        }
        cost = cost + 0.0; // -0 -> +0 so that key order and `<` agree
        P.cost[li] = cost;
        if (cost != cost) {
            nan++;
        } else {
            const unsigned long long key = cost_key(cost);
            klo = key < klo ? key : klo;
            khi = key > khi ? key : khi;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long l = __shfl_down_sync(0xffffffffu, klo, o), h = __shfl_down_sync(0xffffffffu, khi, o);
        klo = l < klo ? l : klo;
        khi = h > khi ? h : khi;
        nan += __shfl_down_sync(0xffffffffu, nan, o);
    }
    if ((threadIdx.x & 31) == 0) {
        s_lo[threadIdx.x >> 5] = klo;
        s_hi[threadIdx.x >> 5] = khi;
        s_nan[threadIdx.x >> 5] = nan;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < CE_BLOCK / 32; ++w) {
            klo = s_lo[w] < klo ? s_lo[w] : klo;
            khi = s_hi[w] > khi ? s_hi[w] : khi;
            nan += s_nan[w];
        }
        if (klo <= khi) {
            atomicMin(&P.hsend[0], klo);
            atomicMax(&P.hsend[1], khi);
        }
        if (nan) atomicAdd(&P.hsend[2], (unsigned long long)nan);
    }
    // the last CTA to finish leaves the work counter at 0 for the next pass (a ticket of its own: the phased mode joins
    // its stages on the host and must not touch the fused kernels' ticket protocol)
    __shared__ int s_last_sample;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        s_last_sample = atomicAdd(&c->sample_ticket, 1u) == gridDim.x - 1;
        if (s_last_sample) {
            c->sample_ticket = 0u;
            c->work_next = 0ULL;
        }
    }
    __syncthreads();
    if (fused == CE_PHASED || !s_last_sample) return;
    __threadfence();
    ce_select_begin(P, ce_gather(P, fused, P.hsend, 4));
}

// K5, levels A and B: histogram of the current level over ALL keys of this rank; level B (`compact`) also copies the keys
// of the current bucket -- the one level A chose -- into the candidate list, which every later level works on.  Fused
// modes: the last CTA picks the bin and, at level B, runs the remaining levels over the candidates by itself.
// Grid: two CTAs of 512 threads per SM at most -- every CTA ends with up to 2048 global atomics, so few fat CTAs.
__global__ void __launch_bounds__(CE_SEL_BLOCK) ce_level_kernel(const CeDev P, int compact, int fused)
{
    __shared__ unsigned int sh[CE_BINS];
    CeCtrl *c = P.ctrl;
    if (!c->active || c->sel_done) return;
    for (int b = threadIdx.x; b < CE_BINS; b += blockDim.x) sh[b] = 0u;
    __syncthreads();
    const unsigned long long base = c->base, bhi = c->bucket_hi;
    const int shift = c->shift;
    const int lane = threadIdx.x & 31;
    const long long npair = (P.nloc + 1) >> 1; // P.cost is 256-byte aligned and padded to ld >= nloc rounded up to 16
    for (long long b0 = (long long)blockIdx.x * blockDim.x; b0 < npair; b0 += (long long)gridDim.x * blockDim.x) {
        const long long pi = b0 + threadIdx.x;
        unsigned long long key[2] = {0ULL, 0ULL};
        bool in[2] = {false, false};
        if (pi < npair) {
            const double2 cv = reinterpret_cast<const double2 *>(P.cost)[pi];
            key[0] = cost_key(cv.x);
            key[1] = cost_key(cv.y);
            in[0] = key[0] >= base && key[0] <= bhi;
            in[1] = 2 * pi + 1 < P.nloc && key[1] >= base && key[1] <= bhi;
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            EQS_CHECK(!in[q] || ((key[q] - base) >> shift) < CE_BINS);
            if (in[q]) atomicAdd(&sh[(key[q] - base) >> shift], 1u);
            if (compact) { // one atomic per warp; the order of the list does not matter
                const unsigned m = __ballot_sync(0xffffffffu, in[q]);
                if (m) {
                    const int leader = __ffs(m) - 1;
                    unsigned long long at = 0ULL;
                    if (lane == leader) at = atomicAdd(&c->n_cand, (unsigned long long)__popc(m));
                    at = __shfl_sync(0xffffffffu, at, leader);
                    EQS_CHECK(at + __popc(m) <= (unsigned long long)P.ld);
                    if (in[q]) P.cand[at + __popc(m & ((1u << lane) - 1u))] = key[q];
                }
            }
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < CE_BINS; b += blockDim.x)
        if (sh[b]) atomicAdd(&P.hsend[b], (unsigned long long)sh[b]);
    if (fused == CE_PHASED) return;
    if (!ce_last_block(c, gridDim.x)) return;
    ce_pick(P, ce_gather(P, fused, P.hsend, CE_BINS));
    if (!compact) return;
    for (int level = 2; level < CE_LEVELS; ++level) { // the same number of rounds on every rank: the picks are global
        if (((const volatile CeCtrl *)c)->sel_done) break;
        ce_cand_hist(P, sh);
        ce_pick(P, ce_gather(P, fused, P.hsend, CE_BINS));
    }
}

// ---- phased mode: the same steps as kernels of their own, the exchange happens between them --------------------------
__global__ void ce_begin_kernel(const CeDev P)
{
    if (!P.ctrl->active) return;
    ce_select_begin(P, Gathered{P.hrecv, CE_BINS, P.world});
}

__global__ void __launch_bounds__(CE_SEL_BLOCK) ce_pick_kernel(const CeDev P)
{
    if (!P.ctrl->active || P.ctrl->sel_done) return;
    ce_pick(P, Gathered{P.hrecv, CE_BINS, P.world});
}

__global__ void __launch_bounds__(CE_SEL_BLOCK) ce_cand_level_kernel(const CeDev P)
{
    __shared__ unsigned int sh[CE_BINS];
    if (!P.ctrl->active || P.ctrl->sel_done) return;
    ce_cand_hist(P, sh);
}

// ---- the elite set in sample order --------------------------------------------------------------------------------

constexpr int CE_PER_THREAD = CE_TILE / CE_BLOCK;

// the keys of this thread's CE_PER_THREAD consecutive samples of tile `tile`; `valid` bit q = sample exists
__device__ __forceinline__ unsigned tile_keys(const CeDev &P, long long tile, unsigned long long (&key)[CE_PER_THREAD])
{
    const long long base = tile * CE_TILE + (long long)threadIdx.x * CE_PER_THREAD;
    unsigned valid = 0u;
#pragma unroll
    for (int q = 0; q < CE_PER_THREAD; q += 2) {
        double2 cv = make_double2(0.0, 0.0);
        if (base + q < P.nloc) cv = *reinterpret_cast<const double2 *>(P.cost + base + q); // ld is a multiple of 16
        key[q] = cost_key(cv.x);
        key[q + 1] = cost_key(cv.y);
        if (base + q < P.nloc) valid |= 1u << q;
        if (base + q + 1 < P.nloc) valid |= 1u << (q + 1);
    }
    return valid;
}

// per tile: samples below the k-th key and samples equal to it; the last CTA turns the counts into offsets
__global__ void __launch_bounds__(CE_BLOCK) ce_count_kernel(const CeDev P)
{
    __shared__ int s_a[33], s_b[33];
    __shared__ long long s_scan[33];
    CeCtrl *c = P.ctrl;
    if (!c->active) return;
    const unsigned long long T = c->base;
    unsigned long long key[CE_PER_THREAD];
    const unsigned valid = tile_keys(P, blockIdx.x, key);
    int less = 0, ties = 0;
#pragma unroll
    for (int q = 0; q < CE_PER_THREAD; ++q) {
        const bool v = (valid >> q) & 1u;
        less += v && key[q] < T;
        ties += v && key[q] == T;
    }
    int tl, tt;
    block_exclusive_scan(less, 0, OpAddI(), s_a, tl);
    block_exclusive_scan(ties, 0, OpAddI(), s_b, tt);
    if (threadIdx.x == 0) {
        P.blk_less[blockIdx.x] = tl;
        P.blk_tie[blockIdx.x] = tt;
    }
    if (!ce_last_block(c, gridDim.x)) return;
    // offsets: every thread owns a run of consecutive tiles; two block scans in all (ties first: how many of a tile's
    // ties are elite depends on the ties before it)
    const long long tie_take = c->tie_take;
    const long long per = (P.ntiles + blockDim.x - 1) / blockDim.x;
    const long long t0 = (long long)threadIdx.x * per, t1 = t0 + per < P.ntiles ? t0 + per : P.ntiles;
    long long my_tie = 0;
    for (long long t = t0; t < t1; ++t) my_tie += __ldcg(P.blk_tie + t);
    long long total;
    const long long ex_tie0 = block_exclusive_scan(my_tie, 0LL, OpAddLL(), s_scan, total);
    long long my_el = 0, run_tie = ex_tie0;
    for (long long t = t0; t < t1; ++t) {
        const long long tc = __ldcg(P.blk_tie + t);
        long long tie_el = tie_take - run_tie;
        tie_el = tie_el < 0 ? 0 : (tie_el > tc ? tc : tie_el);
        my_el += __ldcg(P.blk_less + t) + tie_el;
        run_tie += tc;
    }
    long long run_el = block_exclusive_scan(my_el, 0LL, OpAddLL(), s_scan, total);
    run_tie = ex_tie0;
    for (long long t = t0; t < t1; ++t) {
        const long long tc = __ldcg(P.blk_tie + t);
        long long tie_el = tie_take - run_tie;
        tie_el = tie_el < 0 ? 0 : (tie_el > tc ? tc : tie_el);
        P.blk_tie_off[t] = run_tie;
        P.blk_el_off[t] = run_el;
        run_el += __ldcg(P.blk_less + t) + tie_el;
        run_tie += tc;
    }
    if (threadIdx.x == 0) c->n_elite_local = total;
}

// elite local indices in ascending sample order (deterministic summation order downstream)
__global__ void __launch_bounds__(CE_BLOCK) ce_compact_kernel(const CeDev P)
{
    __shared__ int s_a[33];
    if (!P.ctrl->active) return;
    const unsigned long long T = P.ctrl->base;
    const long long tie_take = P.ctrl->tie_take;
    const long long base = (long long)blockIdx.x * CE_TILE + (long long)threadIdx.x * CE_PER_THREAD;
    unsigned long long key[CE_PER_THREAD];
    const unsigned valid = tile_keys(P, blockIdx.x, key);
    unsigned less = 0u, tie = 0u;
#pragma unroll
    for (int q = 0; q < CE_PER_THREAD; ++q) {
        const bool v = (valid >> q) & 1u;
        if (v && key[q] < T) less |= 1u << q;
        if (v && key[q] == T) tie |= 1u << q;
    }
    int dummy;
    long long tie_rank = P.blk_tie_off[blockIdx.x] + block_exclusive_scan(__popc(tie), 0, OpAddI(), s_a, dummy);
    unsigned el = less;
#pragma unroll
    for (int q = 0; q < CE_PER_THREAD; ++q) {
        if ((tie >> q) & 1u) {
            if (tie_rank < tie_take) el |= 1u << q;
            tie_rank++;
        }
    }
    long long pos = P.blk_el_off[blockIdx.x] + block_exclusive_scan(__popc(el), 0, OpAddI(), s_a, dummy);
#pragma unroll
    for (int q = 0; q < CE_PER_THREAD; ++q)
        if ((el >> q) & 1u) {
            EQS_CHECK(pos >= 0 && pos < P.kcap && base + q < P.nloc);
            P.elite_idx[pos++] = base + q;
        }
}

// regenerate the elites from their counters into Ex[d][j] (the N x n sample matrix never existed)
template <bool REPLAY> __global__ void __launch_bounds__(CE_GEN_BLOCK) ce_elite_gen_kernel(const CeDev P)
{
    extern __shared__ double s_ms[];
    CeCtrl *c = P.ctrl;
    if (!c->active) return;
    const uint32_t t = (uint32_t)(c->iter - 1);
    double *smu = s_ms, *ssig = s_ms + P.n;
    if constexpr (!REPLAY) devmath_tables_load();
    load_mu_sigma(P, smu, ssig);
    const long long ne = c->n_elite_local;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < ne; j += (long long)gridDim.x * blockDim.x) {
        const long long li = P.elite_idx[j];
        EQS_CHECK(li >= 0 && li < P.nloc && j < P.kld);
        ce_sample_coords<REPLAY>(P, t, P.first + li, smu, ssig, [&](double x, int d) { P.Ex[(long long)d * P.kld + j] = x; });
    }
}

// ---- K6: moments ---------------------------------------------------------------------------------------------------

// mu_new = (sum over ranks, in rank order) / 10 or k  (:287).  Any number of threads of one CTA.
__device__ __forceinline__ void ce_mean_apply(const CeDev &P, Gathered g)
{
    for (int d = threadIdx.x; d < P.n; d += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < g.world; ++r) s += g.f64(r, d);
        P.mu_new[d] = s / (P.divisor_mode == EQS_CE_DIV_REFERENCE_TEN ? 10.0 : (double)P.k);
    }
}

// sigma_new = sqrt(sum / (k-1)) (:295-296); mus, sigmas <- new; iter += 1 (:299-302); loop condition (:255).  One CTA.
// A non-finite sigma is an error only if the loop goes on with it: Normal::new (:259) is what panics, and only on sigma.
__device__ __forceinline__ void ce_sigma_apply(const CeDev &P, Gathered g)
{
    __shared__ double s_max[32];
    __shared__ int s_bad[32];
    CeCtrl *c = P.ctrl;
    double mx = 0.0;
    int bad = 0;
    for (int d = threadIdx.x; d < P.n; d += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < g.world; ++r) s += g.f64(r, d);
        const double sg = sqrt(s / (double)(P.k - 1));
        P.sigma[d] = sg;
        P.mu[d] = P.mu_new[d];
        const double a = fabs(sg);
        if (a > mx) mx = a;
        if (!isfinite(sg)) bad = 1;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double w = __shfl_down_sync(0xffffffffu, mx, o);
        mx = w > mx ? w : mx;
        bad |= __shfl_down_sync(0xffffffffu, bad, o);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
        s_max[threadIdx.x >> 5] = mx;
        s_bad[threadIdx.x >> 5] = bad;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)((blockDim.x + 31) >> 5); ++w) {
            mx = s_max[w] > mx ? s_max[w] : mx;
            bad |= s_bad[w];
        }
        c->sigma_inf = mx;
        c->iter += 1;
        const bool go_on = (mx > P.tol) && (c->iter < P.iter_max);
        if (bad && go_on) c->bad_sigma = 1;
        c->active = go_on && !c->bad_sigma && !c->p2p_timeout;
        P.hsend[0] = ~0ULL; // identities of the next pass's (min key, max key, NaN count)
        P.hsend[1] = 0ULL;
        P.hsend[2] = 0ULL;
    }
}

// chunks of CE_CHUNK elites that hold this rank's elites (at least one, so that a rank without elites still writes a zero)
__device__ __forceinline__ int ce_live_chunks(long long ne) { return ne > 0 ? (int)((ne + CE_CHUNK - 1) / CE_CHUNK) : 1; }

// per-dimension partial sums over a chunk of elites; mode 0: sum x, mode 1: sum (x - mu_new)^2 (:283-293).  Fused modes:
// the last CTA adds the chunks in chunk order, gathers the ranks' sums and applies them (mean, or sigma + loop control).
__global__ void __launch_bounds__(CE_BLOCK) ce_rowsum_kernel(const CeDev P, int mode, int fused)
{
    __shared__ double s_w[32];
    CeCtrl *c = P.ctrl;
    if (!c->active) return;
    const long long ne = c->n_elite_local;
    const int d = blockIdx.y;
    // The chunks are those of the elites this rank ACTUALLY holds (ce_live_chunks), walked by however many CTA columns the
    // host launched: the buffers are sized for the worst case (all k elites on one rank), the grid for twice the expected
    // share -- sized by the capacity, a pass over 8 ranks spent more time in empty CTAs than in the exchange.
    const int nact = ce_live_chunks(ne);
    const double m = mode ? P.mu_new[d] : 0.0;
    for (int ch = blockIdx.x; ch < nact; ch += gridDim.x) {
        const long long j0 = (long long)ch * CE_CHUNK, j1 = j0 + CE_CHUNK < ne ? j0 + CE_CHUNK : ne;
        double acc = 0.0;
        for (long long j = j0 + threadIdx.x; j < j1; j += blockDim.x) {
            const double v = P.Ex[(long long)d * P.kld + j];
            if (mode) {
                const double e = v - m;
                acc += e * e;
            } else {
                acc += v;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        __syncthreads(); // s_w of the previous chunk has been read
        if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int w = 0; w < CE_BLOCK / 32; ++w) s += s_w[w];
            P.part[(long long)d * P.nchunks + ch] = s;
        }
    }
    if (fused == CE_PHASED) return;
    if (!ce_last_block(c, gridDim.x * gridDim.y)) return;
    for (int dd = threadIdx.x; dd < P.n; dd += blockDim.x) {
        double s = 0.0;
        for (int ch = 0; ch < nact; ++ch) s += __ldcg(P.part + (long long)dd * P.nchunks + ch);
        P.msend[dd] = s;
    }
    __threadfence();
    __syncthreads();
    const Gathered g = ce_gather(P, fused, reinterpret_cast<const unsigned long long *>(P.msend), P.n);
    if (mode == 0) ce_mean_apply(P, g);
    else ce_sigma_apply(P, g);
}

// phased mode: this rank's per-dimension sum, chunks added in chunk order
__global__ void ce_moment_local_kernel(const CeDev P)
{
    if (!P.ctrl->active) return;
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= P.n) return;
    double s = 0.0;
    const int nact = ce_live_chunks(P.ctrl->n_elite_local);
    for (int ch = 0; ch < nact; ++ch) s += P.part[(long long)d * P.nchunks + ch];
    P.msend[d] = s;
}

__global__ void __launch_bounds__(CE_BLOCK) ce_mean_apply_kernel(const CeDev P)
{
    if (!P.ctrl->active) return;
    ce_mean_apply(P, Gathered{reinterpret_cast<const unsigned long long *>(P.mrecv), P.n, P.world});
}

__global__ void __launch_bounds__(CE_BLOCK) ce_sigma_apply_kernel(const CeDev P)
{
    if (!P.ctrl->active) return;
    ce_sigma_apply(P, Gathered{reinterpret_cast<const unsigned long long *>(P.mrecv), P.n, P.world});
}

// Small sample sizes (N <= CE_SMALL_MAX on one rank): the reference's default shape is 100 samples x 49 passes
// (cross_entropy.rs:8-9, benchmarks/multi.rs:86-100), where a 23-launch pass is nothing but launch latency.  One CTA
// runs `iters` passes: one thread per sample, costs ranked by counting in shared memory (rank = samples with a lower
// cost, ties by lower index -- the order of a stable sort, :276), the k elites regenerated into shared memory IN RANK
// ORDER, and mu / sigma summed by one thread per dimension in exactly the reference's order (:281-297), so with
// replayed normals and a transcendental-free objective the result is bit-identical to the reference's arithmetic.
template <class Obj, bool REPLAY> __global__ void __launch_bounds__(1024, 1) ce_small_kernel(const CeDev P, long long iters)
{
    extern __shared__ double s_dyn[];
    __shared__ int s_nan;
    __shared__ double s_red[32];
    __shared__ int s_bad[32];
    CeCtrl *c = P.ctrl;
    const int tid = threadIdx.x, n = P.n, N = (int)P.nloc, k = (int)P.k;
    double *smu = s_dyn, *ssig = s_dyn + n, *scost = s_dyn + 2 * n, *sEx = s_dyn + 2 * n + N; // [n] [n] [N] [k][n]
    if constexpr (!REPLAY) devmath_tables_load();
    for (long long pass = 0; pass < iters; ++pass) {
        if (!c->active) break; // :255 (uniform: written before the last barrier of the previous pass)
        const uint32_t t = (uint32_t)(c->iter - 1);
        if (tid == 0) s_nan = 0;
        load_mu_sigma(P, smu, ssig);
        if (tid < N) { // :264-274
            const long long gi = P.first + tid;
            double cost;
            if constexpr (Obj::streaming) {
                ObjAcc acc;
                Obj::begin(acc, n);
                ce_sample_coords<REPLAY>(P, t, gi, smu, ssig, [&](double x, int d) { Obj::feed(acc, x, d); });
                cost = Obj::end(acc, n);
            } else {
                ce_sample_coords<REPLAY>(P, t, gi, smu, ssig, [&](double x, int d) { P.xs[(long long)d * P.ld + tid] = x; });
                cost = Obj::eval(StridedView{P.xs + tid, P.ld}, n);
            }
            cost = cost + 0.0;
            scost[tid] = cost;
            P.cost[tid] = cost;
            if (cost != cost) s_nan = 1;
        }
        __syncthreads();
        if (s_nan) { // the reference panics in the sort (:276)
            if (tid == 0) {
                c->nan_cost = 1;
                c->active = 0;
            }
            break;
        }
        if (tid < N) { // rank by counting = position after a stable sort by cost
            const double mine = scost[tid];
            int rank = 0;
            for (int j = 0; j < N; ++j) {
                const double o = scost[j];
                rank += (o < mine) || (o == mine && j < tid);
            }
            if (rank < k) {
                P.elite_idx[rank] = tid;
                double *row = sEx + (size_t)rank * n;
                ce_sample_coords<REPLAY>(P, t, P.first + tid, smu, ssig, [&](double x, int d) { row[d] = x; });
            }
        }
        __syncthreads();
        double mx = 0.0;
        int bad = 0;
        for (int d = tid; d < n; d += blockDim.x) { // :281-297, elites in cost order like the reference
            double mu = 0.0;
            for (int j = 0; j < k; ++j) mu += sEx[(size_t)j * n + d];
            mu /= (P.divisor_mode == EQS_CE_DIV_REFERENCE_TEN ? 10.0 : (double)k);
            double sg = 0.0;
            for (int j = 0; j < k; ++j) {
                const double e = sEx[(size_t)j * n + d] - mu;
                sg += e * e;
            }
            sg = sqrt(sg / (double)(k - 1));
            P.mu[d] = mu;
            P.sigma[d] = sg;
            const double a = fabs(sg);
            if (a > mx) mx = a;
            if (!isfinite(sg)) bad = 1;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double w = __shfl_down_sync(0xffffffffu, mx, o);
            mx = w > mx ? w : mx;
            bad |= __shfl_down_sync(0xffffffffu, bad, o);
        }
        if ((tid & 31) == 0) {
            s_red[tid >> 5] = mx;
            s_bad[tid >> 5] = bad;
        }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
                mx = s_red[w] > mx ? s_red[w] : mx;
                bad |= s_bad[w];
            }
            c->sigma_inf = mx;
            c->n_elite_local = k;
            c->iter += 1; // :302
            const bool go_on = (mx > P.tol) && (c->iter < P.iter_max);
            if (bad && go_on) c->bad_sigma = 1; // Normal::new (:259) panics only if another pass samples with it
            c->active = go_on && !c->bad_sigma;
        }
        __threadfence_block();
        __syncthreads();
    }
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

} // namespace

int validate_ce_params(const eqs_ce_params *p, std::string *why)
{
    auto bad = [&](int st, const char *m) {
        if (why) *why = m;
        return st;
    };
    if (!p) return bad(EQS_ERR_INCORRECT_INPUT, "null parameters");
    if (p->n <= 0 || p->n > 12800) return bad(EQS_ERR_INCORRECT_INPUT, "dimension n out of range (1..12800: mu, sigma tile in shared memory)");
    if (p->sample_size <= 0 || p->sample_size > (1LL << 40)) return bad(EQS_ERR_INCORRECT_INPUT, "sample size out of range (1..2^40)");
    // the reference does not validate k: k < 2 divides by zero (:295) and k > N silently averages N samples (:284)
    if (p->elite_count < 2 || p->elite_count > p->sample_size)
        return bad(EQS_ERR_INCORRECT_INPUT, "importance selection size must satisfy 2 <= k <= sample size");
    if (p->iter_max < 0) return bad(EQS_ERR_INCORRECT_INPUT, "negative iteration count");
    if (p->objective_id < 0 || p->objective_id >= EQS_OBJ_COUNT) return bad(EQS_ERR_INCORRECT_INPUT, "unknown objective id");
    if ((p->objective_id == EQS_OBJ_THREE_CIRCLES || p->objective_id == EQS_OBJ_SINCOS) && p->n != 2)
        return bad(EQS_ERR_INCORRECT_INPUT, "this objective needs n = 2");
    if (p->mean_divisor != EQS_CE_DIV_REFERENCE_TEN && p->mean_divisor != EQS_CE_DIV_ELITE_COUNT)
        return bad(EQS_ERR_INCORRECT_INPUT, "unknown mean divisor mode");
    return EQS_OK;
}

class CeSolver {
public:
    explicit CeSolver(Ctx *ctx) : ctx_(ctx) {}
    ~CeSolver()
    {
        if (ctx_ && ctx_->device >= 0) cudaSetDevice(ctx_->device);
        if (arena_ || h_ctrl_) cudaStreamSynchronize(ctx_->stream);
        ctx_->release_arena(arena_, arena_bytes_);
        if (replay_buf_) cudaFree(replay_buf_);
        ctx_->release_pinned(h_ctrl_);
        if (ev0_) cudaEventDestroy(ev0_);
        if (ev1_) cudaEventDestroy(ev1_);
    }

    int create(const eqs_ce_params *p, const double *x0)
    {
        std::string why;
        const int st = validate_ce_params(p, &why);
        if (st != EQS_OK) return ctx_->fail(st, "%s", why.c_str());
        if (!x0) return ctx_->fail(EQS_ERR_INCORRECT_INPUT, "null x0");
        p_ = *p;
        const int n = p->n;
        int rank = ctx_->rank, world = ctx_->world;
        if (p->exchange == EQS_EXCHANGE_MANUAL) {
            rank = p->virtual_rank;
            world = p->virtual_world;
            if (world < 1 || rank < 0 || rank >= world) return ctx_->fail(EQS_ERR_INCORRECT_INPUT, "bad virtual rank/world");
        }
        manual_ = p->exchange == EQS_EXCHANGE_MANUAL && world > 1;
        // how the stages are joined: on one rank, and over the peer mailboxes, the last CTA of each stage applies it;
        // virtual ranks, and real ranks without peer mapping (NCCL), run the pass phase by phase
        d_.mailbox = nullptr;
        d_.peers = nullptr;
        d_.p2p_timeout_ns = comm_p2p_timeout_ns();
        fused_ = CE_PHASED;
        if (!manual_ && !getenv("EQS_CE_PHASED")) {
            char *mb = nullptr;
            char *const *peers = nullptr;
            if (world == 1) {
                fused_ = CE_FUSED_LOCAL;
            } else if (2 * (size_t)std::max(n, CE_BINS) <= (size_t)MB_LL_CAP && comm_p2p(ctx_, &mb, &peers)) {
                d_.mailbox = mb;
                d_.peers = peers;
                fused_ = CE_FUSED_P2P;
            }
        }
        EQS_CUDA(ctx_, cudaSetDevice(ctx_->device));
        int64_t first = 0, nloc = 0;
        eqs_shard_range(p->sample_size, world, rank, &first, &nloc);
        CeDev &D = d_;
        D.n = n;
        D.world = world;
        D.rank = rank;
        D.first = first;
        D.nloc = nloc;
        D.ntotal = p->sample_size;
        D.k = p->elite_count;
        D.ld = (long long)align_up((size_t)std::max<int64_t>(nloc, 1), 16);
        D.kcap = std::max<long long>(1, std::min<long long>(D.k, nloc));
        D.kld = (long long)align_up((size_t)D.kcap, 16);
        D.ntiles = (nloc + CE_TILE - 1) / CE_TILE;
        D.nchunks = (int)((D.kcap + CE_CHUNK - 1) / CE_CHUNK);
        D.iter_max = p->iter_max;
        D.divisor_mode = p->mean_divisor;
        D.tol = p->tol;
        D.key = philox_key(p->seed);
        D.replay = nullptr;
        D.replay_t0 = 0;
        if ((size_t)2 * n * sizeof(double) > 200 * 1024) return ctx_->fail(EQS_ERR_INCORRECT_INPUT, "n too large for the shared-memory mu/sigma tile");
        bool strided = false;
        dispatch_objective(p->objective_id, [&]<class Obj>() { strided = !Obj::streaming; });

        std::vector<double> sig(n, 1.0); // :247-251
        if (p->std_dev) sig.assign(p->std_dev, p->std_dev + n);
        double smax = 0.0;
        bad_sigma0_ = false;
        for (double s : sig) {
            if (!std::isfinite(s)) bad_sigma0_ = true;
            smax = std::max(smax, std::fabs(s));
        }

        size_t off = 0;
        auto plan = [&](size_t bytes) {
            const size_t o = off;
            off += align_up(std::max<size_t>(bytes, 8), 256);
            return o;
        };
        const size_t o_mu = plan(n * 8), o_sig = plan(n * 8), o_mun = plan(n * 8);
        const size_t o_cost = plan(D.ld * 8), o_cand = plan(D.ld * 8);
        const size_t o_xs = strided ? plan((size_t)n * D.ld * 8) : 0;
        const bool own_recv = !(world == 1 && !manual_);
        const size_t o_hs = plan(CE_BINS * 8), o_hr = own_recv ? plan((size_t)world * CE_BINS * 8) : o_hs;
        const size_t o_ms = plan(n * 8), o_mr = own_recv ? plan((size_t)world * n * 8) : o_ms;
        const size_t nt = (size_t)std::max<long long>(D.ntiles, 1);
        const size_t o_bl = plan(nt * 4), o_bt = plan(nt * 4), o_bto = plan(nt * 8), o_beo = plan(nt * 8);
        const size_t o_ei = plan(D.kcap * 8);
        const size_t o_ex = plan((size_t)n * D.kld * 8);
        const size_t o_part = plan((size_t)n * D.nchunks * 8);
        const size_t o_ctrl = plan(sizeof(CeCtrl));
        EQS_CUDA(ctx_, ctx_->acquire_arena(off, &arena_, &arena_bytes_));
        char *base = (char *)arena_;
        D.mu = (double *)(base + o_mu);
        D.sigma = (double *)(base + o_sig);
        D.mu_new = (double *)(base + o_mun);
        D.cost = (double *)(base + o_cost);
        D.cand = (unsigned long long *)(base + o_cand);
        D.xs = strided ? (double *)(base + o_xs) : nullptr;
        D.hsend = (unsigned long long *)(base + o_hs);
        D.hrecv = (unsigned long long *)(base + o_hr);
        D.msend = (double *)(base + o_ms);
        D.mrecv = (double *)(base + o_mr);
        D.blk_less = (int *)(base + o_bl);
        D.blk_tie = (int *)(base + o_bt);
        D.blk_tie_off = (long long *)(base + o_bto);
        D.blk_el_off = (long long *)(base + o_beo);
        D.elite_idx = (long long *)(base + o_ei);
        D.Ex = (double *)(base + o_ex);
        D.part = (double *)(base + o_part);
        D.ctrl = (CeCtrl *)(base + o_ctrl);
        static_assert(sizeof(CeCtrl) <= Ctx::kPinnedBytes, "pinned block too small");
        EQS_CUDA(ctx_, ctx_->acquire_pinned((void **)&h_ctrl_));
        memset(h_ctrl_, 0, sizeof(CeCtrl));
        h_ctrl_->iter = 1; // :253
        h_ctrl_->sigma_inf = smax;
        bad_sigma0_ = bad_sigma0_ && (smax > p->tol) && (1 < p->iter_max); // an error only if the loop samples with it
        h_ctrl_->active = (smax > p->tol) && (1 < p->iter_max) && !bad_sigma0_;
        cudaStream_t s = ctx_->stream;
        EQS_CUDA(ctx_, cudaMemcpyAsync(D.mu, x0, n * 8, cudaMemcpyHostToDevice, s)); // mus = x0, :246
        EQS_CUDA(ctx_, cudaMemcpyAsync(D.sigma, sig.data(), n * 8, cudaMemcpyHostToDevice, s));
        EQS_CUDA(ctx_, cudaMemcpyAsync(D.ctrl, h_ctrl_, sizeof(CeCtrl), cudaMemcpyHostToDevice, s));
        EQS_CUDA(ctx_, cudaMemsetAsync(D.hsend, 0, CE_BINS * 8, s));
        EQS_CUDA(ctx_, cudaMemsetAsync(D.hsend, 0xff, 8, s)); // identity of the pass's min key (words 1, 2: max key, NaN count)
        EQS_CUDA(ctx_, cudaStreamSynchronize(s));
        EQS_CUDA(ctx_, cudaEventCreate(&ev0_));
        EQS_CUDA(ctx_, cudaEventCreate(&ev1_));
        sample_grid_ = 0;
        return EQS_OK;
    }

    int upload_replay(const double *host, int64_t doubles)
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

    int fetch_ctrl()
    {
        EQS_CUDA(ctx_, cudaMemcpyAsync(h_ctrl_, d_.ctrl, sizeof(CeCtrl), cudaMemcpyDeviceToHost, ctx_->stream));
        EQS_CUDA(ctx_, cudaStreamSynchronize(ctx_->stream));
        return EQS_OK;
    }

    int grid_for(long long units) const
    {
        const long long need = std::max<long long>(1, (units + CE_BLOCK - 1) / CE_BLOCK);
        return (int)std::min<long long>(need, (long long)ctx_->sm_count * 8);
    }

    // phase kinds of exchange: 0 none, 1 histogram-sized (u64[BINS]: key range of the pass, or a level), 2 moments (f64[n])
    static int phase_exchange(int phase) { return phase <= 6 ? 1 : (phase <= 8 ? 2 : 0); }

    int launch_sample(int fused)
    {
        cudaStream_t s = ctx_->stream;
        const CeDev &D = d_;
        const bool replay = D.replay != nullptr;
        const size_t smem = (size_t)2 * D.n * sizeof(double);
        int rc = EQS_OK;
        dispatch_objective(p_.objective_id, [&]<class Obj>() {
            auto run = [&](auto kernel) {
                if (smem > kSmemOptIn) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (sample_grid_ == 0) {
                    int nb = 0;
                    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, CE_BLOCK, smem);
                    if (nb < 1) {
                        rc = ctx_->fail(EQS_ERR_EXTERNAL, "sample kernel does not fit on an SM");
                        return;
                    }
                    const long long need = std::max<long long>(1, (D.nloc + CE_BLOCK - 1) / CE_BLOCK);
                    sample_grid_ = (int)std::min<long long>(need, (long long)nb * ctx_->sm_count);
                }
                kernel<<<sample_grid_, CE_BLOCK, smem, s>>>(D, fused);
                launches_++;
            };
            if (replay) run(ce_sample_kernel<Obj, true>);
            else run(ce_sample_kernel<Obj, false>);
        });
        return rc;
    }

    void launch_level(int compact, int fused)
    {
        const long long need = std::max<long long>(1, ((d_.nloc + 1) / 2 + CE_SEL_BLOCK - 1) / CE_SEL_BLOCK);
        const int grid = (int)std::min<long long>(need, 2LL * ctx_->sm_count);
        ce_level_kernel<<<grid, CE_SEL_BLOCK, 0, ctx_->stream>>>(d_, compact, fused);
        launches_++;
    }

    // the elite list in sample order and the regenerated elites
    void launch_elites()
    {
        cudaStream_t s = ctx_->stream;
        const CeDev &D = d_;
        const size_t smem = (size_t)2 * D.n * sizeof(double);
        ce_count_kernel<<<(unsigned)std::max<long long>(D.ntiles, 1), CE_BLOCK, 0, s>>>(D);
        launches_++;
        if (D.ntiles > 0) {
            ce_compact_kernel<<<(unsigned)D.ntiles, CE_BLOCK, 0, s>>>(D);
            launches_++;
        }
        auto gen = [&](auto kernel) {
            if (smem > kSmemOptIn) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            const long long need = std::max<long long>(1, (D.kcap + CE_GEN_BLOCK - 1) / CE_GEN_BLOCK);
            kernel<<<(int)std::min<long long>(need, 32LL * ctx_->sm_count), CE_GEN_BLOCK, smem, s>>>(D);
            launches_++;
        };
        if (D.replay != nullptr) gen(ce_elite_gen_kernel<true>);
        else gen(ce_elite_gen_kernel<false>);
    }

    void launch_rowsum(int mode, int fused)
    {
        // CTA columns for twice the share of the elites a rank expects (k / world); the kernel walks whatever it really has
        const long long expect = (d_.k + d_.world - 1) / d_.world;
        const long long cols = std::min<long long>(d_.nchunks, (2 * expect + CE_CHUNK - 1) / CE_CHUNK + 1);
        ce_rowsum_kernel<<<dim3((unsigned)std::max<long long>(cols, 1), d_.n), CE_BLOCK, 0, ctx_->stream>>>(d_, mode, fused);
        launches_++;
    }

    // one pass, fused: 8 launches, every exchange inside the kernels
    int launch_pass_fused()
    {
        EQS_TRY(launch_sample(fused_));
        launch_level(0, fused_);
        launch_level(1, fused_);
        launch_elites();
        launch_rowsum(0, fused_);
        launch_rowsum(1, fused_);
        EQS_CUDA(ctx_, cudaGetLastError());
        return EQS_OK;
    }

    // one phase of the phased pass (the exchange named by phase_exchange follows it)
    int launch_phase(int phase)
    {
        cudaStream_t s = ctx_->stream;
        const CeDev &D = d_;
        const int dim_grid = (D.n + 127) / 128;
        auto pick = [&] {
            ce_pick_kernel<<<1, CE_SEL_BLOCK, 0, s>>>(D);
            launches_++;
        };
        if (phase == 0) {
            EQS_TRY(launch_sample(CE_PHASED));
        } else if (phase == 1) {
            ce_begin_kernel<<<1, 32, 0, s>>>(D);
            launches_++;
            launch_level(0, CE_PHASED);
        } else if (phase == 2) {
            pick();
            launch_level(1, CE_PHASED);
        } else if (phase <= 6) {
            pick();
            ce_cand_level_kernel<<<1, CE_SEL_BLOCK, 0, s>>>(D);
            launches_++;
        } else if (phase == 7) {
            pick();
            launch_elites();
            launch_rowsum(0, CE_PHASED);
            ce_moment_local_kernel<<<dim_grid, 128, 0, s>>>(D);
            launches_++;
        } else if (phase == 8) {
            ce_mean_apply_kernel<<<1, CE_BLOCK, 0, s>>>(D);
            launches_++;
            launch_rowsum(1, CE_PHASED);
            ce_moment_local_kernel<<<dim_grid, 128, 0, s>>>(D);
            launches_++;
        } else if (phase == 9) {
            ce_sigma_apply_kernel<<<1, CE_BLOCK, 0, s>>>(D);
            launches_++;
        } else {
            return ctx_->fail(EQS_ERR_INCORRECT_INPUT, "phase must be 0..%d", CE_PHASES - 1);
        }
        EQS_CUDA(ctx_, cudaGetLastError());
        return EQS_OK;
    }

    int exchange(int kind)
    {
        if (manual_ || d_.world == 1 || kind == 0) return EQS_OK;
        if (kind == 1)
            return comm_allgather_f64(ctx_, (const double *)d_.hsend, (double *)d_.hrecv, CE_BINS, ctx_->stream);
        return comm_allgather_f64(ctx_, d_.msend, d_.mrecv, d_.n, ctx_->stream);
    }

    int prepare_replay(const double *replay, int64_t iters)
    {
        if (replay) {
            EQS_TRY(fetch_ctrl());
            EQS_TRY(upload_replay(replay, iters * p_.sample_size * (int64_t)p_.n));
            d_.replay = replay_buf_;
            d_.replay_t0 = h_ctrl_->iter - 1;
        } else {
            d_.replay = nullptr;
        }
        sample_grid_ = 0;
        return EQS_OK;
    }

    int step(int64_t iters, const double *replay)
    {
        EQS_CUDA(ctx_, cudaSetDevice(ctx_->device));
        if (manual_) return ctx_->fail(EQS_ERR_INCORRECT_INPUT, "exchange = MANUAL: drive the phases with eqs_ce_phase");
        if (bad_sigma0_) return ctx_->fail(EQS_ERR_NAN, "non-finite standard deviation (the reference panics in Normal::new, cross_entropy.rs:259)");
        EQS_TRY(prepare_replay(replay, iters));
        const int64_t l0 = launches_;
        EQS_CUDA(ctx_, cudaEventRecord(ev0_, ctx_->stream));
        // small sample sizes on one rank: all passes of the batch in ONE single-CTA launch (ce_small_kernel)
        const size_t small_smem = ((size_t)2 * d_.n + d_.nloc + (size_t)d_.k * d_.n) * sizeof(double);
        const bool small = d_.world == 1 && d_.nloc >= 2 && d_.nloc <= CE_SMALL_MAX && small_smem <= 160 * 1024 &&
                           !getenv("EQS_CE_NO_SMALL");
        if (small && iters > 0) {
            const bool rp = d_.replay != nullptr;
            const int block = (int)std::min<long long>(1024, (std::max<long long>(d_.nloc, std::min(d_.n, 256)) + 31) / 32 * 32);
            int rc = EQS_OK;
            dispatch_objective(p_.objective_id, [&]<class Obj>() {
                auto run = [&](auto kernel) {
                    if (small_smem > kSmemOptIn &&
                        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_smem) != cudaSuccess)
                        rc = ctx_->fail(EQS_ERR_EXTERNAL, "cannot reserve %zu bytes of shared memory", small_smem);
                    if (rc != EQS_OK) return;
                    kernel<<<1, block, small_smem, ctx_->stream>>>(d_, (long long)iters);
                    launches_++;
                };
                if (rp) run(ce_small_kernel<Obj, true>);
                else run(ce_small_kernel<Obj, false>);
            });
            EQS_TRY(rc);
            EQS_CUDA(ctx_, cudaGetLastError());
            small_path_ = true;
        } else {
            small_path_ = false;
        }
        for (int64_t it = 0; it < (small ? 0 : iters); ++it) {
            if (fused_ != CE_PHASED) {
                EQS_TRY(launch_pass_fused());
                continue;
            }
            for (int ph = 0; ph < CE_PHASES; ++ph) {
                EQS_TRY(launch_phase(ph));
                EQS_TRY(exchange(phase_exchange(ph)));
            }
        }
        EQS_CUDA(ctx_, cudaEventRecord(ev1_, ctx_->stream));
        last_launches_ = launches_ - l0;
        return EQS_OK;
    }

    int check_flags()
    {
        if (h_ctrl_->p2p_timeout) return ctx_->fail(EQS_ERR_EXTERNAL, "a peer rank did not reach an exchange of the CrossEntropy pass (peer-memory mailbox timed out)");
        if (h_ctrl_->nan_cost) return ctx_->fail(EQS_ERR_NAN, "a sample cost is NaN (the reference panics in sort_unstable_by, cross_entropy.rs:276)");
        if (h_ctrl_->bad_sigma) return ctx_->fail(EQS_ERR_NAN, "a non-finite standard deviation was produced and the loop goes on (the reference panics in Normal::new, cross_entropy.rs:259)");
        return EQS_OK;
    }

    int read_state(double *mu, double *sigma, int64_t *iters, int32_t *active)
    {
        EQS_CUDA(ctx_, cudaSetDevice(ctx_->device));
        cudaStream_t s = ctx_->stream;
        if (mu) EQS_CUDA(ctx_, cudaMemcpyAsync(mu, d_.mu, d_.n * 8, cudaMemcpyDeviceToHost, s));
        if (sigma) EQS_CUDA(ctx_, cudaMemcpyAsync(sigma, d_.sigma, d_.n * 8, cudaMemcpyDeviceToHost, s));
        EQS_TRY(fetch_ctrl());
        if (iters) *iters = h_ctrl_->iter - 1;
        if (active) *active = h_ctrl_->active;
        return check_flags();
    }

    // CrossEntropy::solve, cross_entropy.rs:245-306
    int solve(double *out_mu, eqs_report *rep)
    {
        Event e0, e1;
        EQS_CUDA(ctx_, e0.create());
        EQS_CUDA(ctx_, e1.create());
        const int64_t l0 = launches_;
        cudaEventRecord(e0, ctx_->stream);
        int st = EQS_OK;
        const int64_t kBatch = 16;
        int64_t left = std::max<int64_t>(p_.iter_max - 1, 0); // iter starts at 1 (:253,:255)
        while (st == EQS_OK && left > 0) {
            const int64_t b = std::min(kBatch, left);
            st = step(b, nullptr);
            left -= b;
            if (st == EQS_OK) st = fetch_ctrl();
            if (st == EQS_OK) st = check_flags();
            if (st == EQS_OK && !h_ctrl_->active) break;
        }
        if (st == EQS_OK) {
            cudaEventRecord(e1, ctx_->stream);
            st = read_state(out_mu, nullptr, nullptr, nullptr); // Ok(mus), :305
        }
        if (st == EQS_OK && rep) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            rep->iters_run = h_ctrl_->iter - 1;
            rep->evals = p_.sample_size * (h_ctrl_->iter - 1);
            rep->best_cost = h_ctrl_->sigma_inf;
            rep->stalled = !(h_ctrl_->sigma_inf > p_.tol);
            rep->world = d_.world;
            rep->kernel_launches = launches_ - l0;
            rep->init_ms = 0.0;
            rep->loop_ms = ms;
        }
        return st;
    }

    Ctx *ctx_;
    eqs_ce_params p_{};
    CeDev d_{};
    void *arena_ = nullptr;
    size_t arena_bytes_ = 0;
    double *replay_buf_ = nullptr;
    size_t replay_cap_ = 0;
    CeCtrl *h_ctrl_ = nullptr;
    cudaEvent_t ev0_ = nullptr, ev1_ = nullptr;
    int sample_grid_ = 0;
    int64_t launches_ = 0, last_launches_ = 0;
    bool manual_ = false, bad_sigma0_ = false;
    int fused_ = CE_PHASED;
    bool small_path_ = false; // the last step() ran ce_small_kernel (elite list is in cost order)
};

} // namespace eqs

using namespace eqs;

struct eqs_ce {
    CeSolver s;
    explicit eqs_ce(Ctx *c) : s(c) {}
};

#define H_OR_FAIL(h)                                                                                             \
    if (!(h)) return EQS_ERR_INCORRECT_INPUT

extern "C" {

int eqs_ce_validate(const eqs_ce_params *p) { return validate_ce_params(p, nullptr); }

int eqs_ce_create(eqs_ctx *ctx, const eqs_ce_params *p, const double *x0, eqs_ce **out)
{
    EQS_BOUNDARY_TRY
    if (!ctx || !out) return EQS_ERR_INCORRECT_INPUT;
    *out = nullptr;
    eqs_ce *h = new (std::nothrow) eqs_ce(&ctx->c);
    if (!h) return ctx->c.fail(EQS_ERR_EXTERNAL, "out of host memory");
    const int st = h->s.create(p, x0);
    if (st != EQS_OK) {
        delete h;
        return st;
    }
    *out = h;
    return EQS_OK;
    EQS_BOUNDARY_CATCH
}

int eqs_ce_destroy(eqs_ce *h)
{
    EQS_BOUNDARY_TRY
    delete h;
    return EQS_OK;
    EQS_BOUNDARY_CATCH
}

int eqs_ce_step(eqs_ce *h, int64_t iters, const double *replay_normals)
{
    EQS_BOUNDARY_TRY
    H_OR_FAIL(h);
    if (iters < 0) return h->s.ctx_->fail(EQS_ERR_INCORRECT_INPUT, "negative iteration count");
    return h->s.step(iters, replay_normals);
    EQS_BOUNDARY_CATCH
}

int eqs_ce_sync(eqs_ce *h)
{
    EQS_BOUNDARY_TRY
    H_OR_FAIL(h);
    Ctx *c = h->s.ctx_;
    EQS_CUDA(c, cudaSetDevice(c->device));
    EQS_CUDA(c, cudaStreamSynchronize(c->stream));
    return EQS_OK;
    EQS_BOUNDARY_CATCH
}

int eqs_ce_last_step_ms(eqs_ce *h, double *ms, int64_t *kernel_launches)
{
    EQS_BOUNDARY_TRY
    H_OR_FAIL(h);
    Ctx *c = h->s.ctx_;
    EQS_CUDA(c, cudaEventSynchronize(h->s.ev1_));
    float f = 0.f;
    EQS_CUDA(c, cudaEventElapsedTime(&f, h->s.ev0_, h->s.ev1_));
    if (ms) *ms = f;
    if (kernel_launches) *kernel_launches = h->s.last_launches_;
    return EQS_OK;
    EQS_BOUNDARY_CATCH
}

int eqs_ce_read_state(eqs_ce *h, double *mu, double *sigma, int64_t *iters_run, int32_t *active)
{
    EQS_BOUNDARY_TRY
    H_OR_FAIL(h);
    return h->s.read_state(mu, sigma, iters_run, active);
    EQS_BOUNDARY_CATCH
}

int eqs_ce_read_elites(eqs_ce *h, int64_t *idx, int64_t *count)
{
    EQS_BOUNDARY_TRY
    H_OR_FAIL(h);
    CeSolver &s = h->s;
    Ctx *c = s.ctx_;
    EQS_CUDA(c, cudaSetDevice(c->device));
    EQS_TRY(s.fetch_ctrl());
    const long long ne = s.h_ctrl_->n_elite_local;
    if (count) *count = ne;
    if (idx && ne > 0) {
        EQS_CUDA(c, cudaMemcpyAsync(idx, s.d_.elite_idx, ne * 8, cudaMemcpyDeviceToHost, c->stream));
        EQS_CUDA(c, cudaStreamSynchronize(c->stream));
        for (long long j = 0; j < ne; ++j) idx[j] += s.d_.first; // local -> global
        if (s.small_path_) std::sort(idx, idx + ne);              // the contract is ascending sample order
    }
    return EQS_OK;
    EQS_BOUNDARY_CATCH
}

int eqs_ce_read_costs(eqs_ce *h, double *costs)
{
    EQS_BOUNDARY_TRY
    H_OR_FAIL(h);
    CeSolver &s = h->s;
    Ctx *c = s.ctx_;
    if (!costs) return EQS_ERR_INCORRECT_INPUT;
    EQS_CUDA(c, cudaSetDevice(c->device));
    if (s.d_.nloc > 0) EQS_CUDA(c, cudaMemcpyAsync(costs, s.d_.cost, s.d_.nloc * 8, cudaMemcpyDeviceToHost, c->stream));
    EQS_CUDA(c, cudaStreamSynchronize(c->stream));
    return EQS_OK;
    EQS_BOUNDARY_CATCH
}

int eqs_ce_local_range(eqs_ce *h, int64_t *first, int64_t *local)
{
    EQS_BOUNDARY_TRY
    H_OR_FAIL(h);
    if (first) *first = h->s.d_.first;
    if (local) *local = h->s.d_.nloc;
    return EQS_OK;
    EQS_BOUNDARY_CATCH
}

int eqs_ce_phase(eqs_ce *h, int32_t phase, const double *replay_normals)
{
    EQS_BOUNDARY_TRY
    H_OR_FAIL(h);
    CeSolver &s = h->s;
    Ctx *c = s.ctx_;
    if (phase < 0 || phase >= CE_PHASES) return c->fail(EQS_ERR_INCORRECT_INPUT, "phase out of range (0..%d)", CE_PHASES - 1);
    EQS_CUDA(c, cudaSetDevice(c->device));
    if (phase == 0) EQS_TRY(s.prepare_replay(replay_normals, 1));
    return s.launch_phase(phase);
    EQS_BOUNDARY_CATCH
}

int64_t eqs_ce_exchange_words(eqs_ce *h, int32_t phase)
{
    if (!h) return 0;
    const int kind = CeSolver::phase_exchange(phase);
    return kind == 1 ? CE_BINS : (kind == 2 ? h->s.d_.n : 0);
}

int eqs_ce_get_exchange(eqs_ce *h, int32_t phase, void *words)
{
    EQS_BOUNDARY_TRY
    H_OR_FAIL(h);
    CeSolver &s = h->s;
    Ctx *c = s.ctx_;
    const int kind = CeSolver::phase_exchange(phase);
    if (kind == 0 || !words) return c->fail(EQS_ERR_INCORRECT_INPUT, "phase has no exchange");
    EQS_CUDA(c, cudaSetDevice(c->device));
    const void *src = kind == 1 ? (const void *)s.d_.hsend : (const void *)s.d_.msend;
    const size_t bytes = (kind == 1 ? (size_t)CE_BINS : (size_t)s.d_.n) * 8;
    EQS_CUDA(c, cudaMemcpyAsync(words, src, bytes, cudaMemcpyDeviceToHost, c->stream));
    EQS_CUDA(c, cudaStreamSynchronize(c->stream));
    return EQS_OK;
    EQS_BOUNDARY_CATCH
}

int eqs_ce_set_exchange(eqs_ce *h, int32_t phase, const void *words)
{
    EQS_BOUNDARY_TRY
    H_OR_FAIL(h);
    CeSolver &s = h->s;
    Ctx *c = s.ctx_;
    const int kind = CeSolver::phase_exchange(phase);
    if (kind == 0 || !words) return c->fail(EQS_ERR_INCORRECT_INPUT, "phase has no exchange");
    EQS_CUDA(c, cudaSetDevice(c->device));
    void *dst = kind == 1 ? (void *)s.d_.hrecv : (void *)s.d_.mrecv;
    const size_t bytes = (kind == 1 ? (size_t)CE_BINS : (size_t)s.d_.n) * 8 * s.d_.world;
    EQS_CUDA(c, cudaMemcpyAsync(dst, words, bytes, cudaMemcpyHostToDevice, c->stream));
    EQS_CUDA(c, cudaStreamSynchronize(c->stream));
    return EQS_OK;
    EQS_BOUNDARY_CATCH
}

int eqs_ce_solve(eqs_ctx *ctx, const eqs_ce_params *p, const double *x0, double *out_mu, eqs_report *report)
{
    EQS_BOUNDARY_TRY
    if (!ctx) return EQS_ERR_INCORRECT_INPUT;
    if (!x0 || !out_mu) return ctx->c.fail(EQS_ERR_INCORRECT_INPUT, "null x0 / out_mu");
    CeSolver s(&ctx->c);
    EQS_TRY(s.create(p, x0));
    return s.solve(out_mu, report);
    EQS_BOUNDARY_CATCH
}

} // extern "C"
