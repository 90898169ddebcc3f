/*
 * eqsolver_b200.h -- C ABI of the B200-native population step of eqsolver's global optimisers.
 *
 * This is the drop-in boundary.  The reference (AzeezDa/eqsolver v0.4.1, pure Rust) has no FFI; the
 * interface it exposes for this path is the builder API of
 *   src/solvers/global_optimisers/particle_swarm.rs  (new :119-153, with_* :179-333, solve :356-431)
 *   src/solvers/global_optimisers/cross_entropy.rs   (new :75-85,  with_* :108-223, solve :245-306)
 * Every entry point below names the reference lines it replaces.  rust/eqsolver-cuda-sys binds these
 * symbols one-to-one and rust/eqsolver-b200 re-creates the builders on top.
 *
 * Conventions: every function returns an eqs_status (0 = ok) and never unwinds; all pointers are
 * plain host pointers to contiguous f64 / i64 owned by the caller; the context owns device memory,
 * streams, graphs and the NCCL communicator.  Arithmetic is IEEE f64 throughout.  There is no CPU
 * fallback: without a CUDA device every compute entry point fails with EQS_ERR_EXTERNAL.
 * Distinct handles may be used from distinct threads; calls on one handle must be serialised.
 */
#ifndef EQSOLVER_B200_H
#define EQSOLVER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define EQS_ABI_VERSION 1
#define EQS_STALL_WINDOW 50 /* DEFAULT_STALL_ITERATIONS, particle_swarm.rs:12 */

/* SolverError, reference src/lib.rs:19-44 (same order; 0 = Ok). */
typedef enum {
    EQS_OK = 0,
    EQS_ERR_MAX_ITER = 1,        /* MaxIterReached(usize)   -- never produced by this path */
    EQS_ERR_NAN = 2,             /* NotANumber: where the reference panics on a NaN cost / on sampling with a non-finite sigma */
    EQS_ERR_INCORRECT_INPUT = 3, /* IncorrectInput{details} */
    EQS_ERR_BAD_JACOBIAN = 4,    /* BadJacobian             -- never produced by this path */
    EQS_ERR_TYPE_CONVERSION = 5, /* TypeConversionError     -- MISER only (miser.rs:413-419) */
    EQS_ERR_EXTERNAL = 6         /* ExternalError(String): bad bounds (rand's message), CUDA / NCCL failures */
} eqs_status;

/* The user closure F: Fn(Vector) -> T (particle_swarm.rs:92, cross_entropy.rs:55) cannot run on the
 * device; objectives are ahead-of-time compiled device functors selected by ID. */
typedef enum {
    EQS_OBJ_SPHERE = 0,
    EQS_OBJ_ROSENBROCK = 1,
    EQS_OBJ_RASTRIGIN = 2,     /* examples/global_optimisation.rs:9-17, benchmarks/multi.rs:155-163 */
    EQS_OBJ_ACKLEY = 3,
    EQS_OBJ_THREE_CIRCLES = 4, /* examples/global_optimiser_as_solver.rs:22-42 (n = 2) */
    EQS_OBJ_HEAVY = 5,         /* benchmarks/multi.rs:106-115,137 */
    EQS_OBJ_USER = 6,          /* eqsolver_b200/csrc/user_objective.cuh, rebuilt by the user */
    EQS_OBJ_GAUSSIAN = 7,      /* exp(-sum x^2): the integrand of tests/integrators.rs:11-13 (n = 1 there) */
    EQS_OBJ_SINCOS = 8,        /* sin(cos(x1) + x0), n = 2: tests/integrators.rs:15-17 */
    EQS_OBJ_SERIES = 9,        /* sum_i (-1)^i x^i, i < 1000, summed over the dimensions: benchmarks/integrators.rs:69-75 */
    EQS_OBJ_COUNT = 10
} eqs_objective;

typedef enum {
    EQS_PSO_SEQUENTIAL_EXACT = 0, /* the reference: gbest moves inside the particle loop, :420-425 */
    EQS_PSO_SYNCHRONOUS = 1       /* one min-loc gbest update per iteration (throughput mode) */
} eqs_pso_mode;

typedef enum {
    EQS_CE_DIV_REFERENCE_TEN = 0, /* the literal 10 of cross_entropy.rs:287 */
    EQS_CE_DIV_ELITE_COUNT = 1    /* divide the elite sum by k */
} eqs_ce_divisor;

typedef enum {
    EQS_EXCHANGE_AUTO = 0,   /* none when world == 1, NCCL over NVLink when the context has a communicator */
    EQS_EXCHANGE_MANUAL = 1  /* split-phase: the caller moves the records (virtual ranks on one GPU) */
} eqs_exchange;

typedef struct eqs_ctx eqs_ctx;
typedef struct eqs_pso eqs_pso;
typedef struct eqs_ce eqs_ce;

/* ParticleSwarm fields (particle_swarm.rs:90-105) + what a closure-free, seedable, sharded build needs. */
typedef struct {
    int32_t n;                /* dimension D, 1..12800 (gbest tile in shared memory) */
    int32_t objective_id;     /* eqs_objective, replaces `f` */
    int32_t mode;             /* eqs_pso_mode */
    int32_t exchange;         /* eqs_exchange */
    int64_t particle_count;   /* GLOBAL particle count, with_particle_count :305-333 (default 100), at most 2^40 */
    int64_t iter_max;         /* with_iter_max (default 50, lib.rs:16) */
    double tol;               /* with_tol (default 1e-6, lib.rs:13) */
    double w, c1, c2;         /* inertia / cognitive / social, defaults 0.5 / 1 / 1 (:8-10) */
    const double *lower;      /* [n] */
    const double *upper;      /* [n] */
    uint64_t seed;            /* Philox4x32-10 key; the reference's rng() is unseedable */
    /* replay (parity mode): u in [0,1) in the reference's draw order, indexed by GLOBAL particle.
     * init [N][2n] = n position draws then n velocity draws (:363-373); NULL => device Philox. */
    const double *replay_init;
    /* manual sharding for EQS_EXCHANGE_MANUAL (virtual ranks); ignored otherwise */
    int32_t virtual_rank, virtual_world;
} eqs_pso_params;

/* CrossEntropy fields (cross_entropy.rs:51-65). */
typedef struct {
    int32_t n;
    int32_t objective_id;
    int32_t mean_divisor;     /* eqs_ce_divisor */
    int32_t exchange;         /* eqs_exchange */
    int64_t sample_size;      /* GLOBAL, with_sample_size (default 100, :8), at most 2^40 */
    int64_t elite_count;      /* with_importance_selection_size (default 10, :9); 2 <= k <= N */
    int64_t iter_max;         /* the loop runs at most iter_max-1 times (:253-255) */
    double tol;
    const double *std_dev;    /* [n] or NULL => ones (:247-251) */
    uint64_t seed;
    int32_t virtual_rank, virtual_world;
} eqs_ce_params;

typedef struct {
    int64_t iters_run;        /* outer iterations that executed */
    int64_t evals;            /* objective evaluations (global) */
    double best_cost;         /* PSO: gbest cost; CE: max |sigma| at exit */
    int32_t stalled;          /* PSO: left through the stall test :395-397; CE: sigma <= tol */
    int32_t world;            /* ranks that shared the population */
    int64_t kernel_launches;  /* kernels of this library launched by the call */
    double init_ms;           /* device time of the init phase (CUDA events) */
    double loop_ms;           /* device time of the iteration loop (CUDA events) */
} eqs_report;

/* ---- context ------------------------------------------------------------------------------- */
int eqs_abi_version(void);
int eqs_ctx_create(int device, eqs_ctx **out);
int eqs_ctx_destroy(eqs_ctx *ctx);
/* message behind the last non-zero status on this context (feeds SolverError::ExternalError(String));
 * ctx == NULL reads the message of the last failed eqs_ctx_create on this thread.  The pointer is a per-thread
 * copy, valid until this thread's next call of eqs_last_error: other threads may fail on the same context meanwhile.
 * A C++ exception stopped at the boundary by the calling thread's last entry (out of host memory ...) is reported here
 * too. */
const char *eqs_last_error(const eqs_ctx *ctx);
int eqs_ctx_sm_count(const eqs_ctx *ctx, int *out);

/* multi-GPU: one process per GPU.  Rank 0 calls eqs_comm_unique_id, ships the 128 bytes to the other
 * ranks (torch.distributed, MPI, a file ...), every rank calls eqs_ctx_comm_init. */
int eqs_comm_unique_id(void *out128);
int eqs_ctx_comm_init(eqs_ctx *ctx, int rank, int world, const void *id128);
int eqs_ctx_rank(const eqs_ctx *ctx, int *rank, int *world);
/* contiguous index blocks per rank (keeps the reference's particle order); pure host arithmetic */
int eqs_shard_range(int64_t count, int world, int rank, int64_t *first, int64_t *local);
/* doubles in one exchange record for dimension n: cost, index, pushes, pad, ring tail[50], the global indices of the
 * particles that pushed the tail[50], x[n] */
int64_t eqs_record_doubles(int32_t n);

/* ---- diagnostics --------------------------------------------------------------------------- */
/* the user closure evaluated on the device: x [count][n] -> out [count] (objective KATs) */
int eqs_objective_eval(eqs_ctx *ctx, int32_t objective_id, const double *x, int32_t n, int64_t count,
                       double *out);
/* raw Philox4x32-10 blocks from the device generator: out [count][4] */
int eqs_philox_blocks(eqs_ctx *ctx, const uint32_t *ctr /*[count][4]*/, const uint32_t key[2], int64_t count,
                      uint32_t *out);
/* draws exactly as the kernels consume them.  kind 0: PSO init (pos_u, vel_u), 1: PSO step (rp, rg),
 * 2: CE standard normals (a = z, b unused).  a, b: [count][n] for global indices first..first+count-1. */
int eqs_draws(eqs_ctx *ctx, int32_t kind, uint64_t seed, int64_t t, int64_t first, int64_t count, int32_t n,
              double *a, double *b);
/* the device cosine used by the Rastrigin / Ackley functors (accuracy tests): out[i] = cos(x[i]) */
int eqs_device_cos(eqs_ctx *ctx, const double *x, int64_t count, double *out);
/* measured machine ceilings for the roofline: DFMA issue rate (TFLOP/s) and copy bandwidth (GB/s) */
int eqs_measure_fp64(eqs_ctx *ctx, double *tflops);
int eqs_measure_copy(eqs_ctx *ctx, int64_t bytes, double *gbps);
/* the ParticleSwarm step's access pattern with the arithmetic removed (3 buffers read, 2 of them written in place,
 * 256-bit accesses in the [n/4][particles][4] layout; n a multiple of 8): GB/s over the 5 streams */
int eqs_measure_pattern(eqs_ctx *ctx, int32_t n, int64_t particles, double *gbps);

/* ---- ParticleSwarm ------------------------------------------------------------------------- */
/* ParticleSwarm::new validation, particle_swarm.rs:124-139: EQS_ERR_EXTERNAL when some lower > upper
 * or a bound is not finite (rand's Uniform::new_inclusive error), EQS_ERR_INCORRECT_INPUT for n <= 0. */
int eqs_pso_validate(const eqs_pso_params *p);
/* ParticleSwarm::solve(x0), :356-431.  out_x [n] receives the gbest position (always Ok in the reference). */
int eqs_pso_solve(eqs_ctx *ctx, const eqs_pso_params *p, const double *x0, double *out_x, eqs_report *report);

/* step-granular twins (parity, tracing, benchmarking with the population resident in HBM) */
int eqs_pso_create(eqs_ctx *ctx, const eqs_pso_params *p, eqs_pso **out);
int eqs_pso_destroy(eqs_pso *h);
int eqs_pso_init(eqs_pso *h, const double *x0);                   /* :357-389 */
/* runs up to `iters` passes of the outer loop :394-428 back to back on the device; passes after the
 * stall test fires are no-ops.  replay_step: NULL or u in [0,1) [iters][N][n][2] (rp, rg; :407-408). */
int eqs_pso_step(eqs_pso *h, int64_t iters, const double *replay_step);
int eqs_pso_sync(eqs_pso *h);
/* device milliseconds (CUDA events on the solver stream) of the last eqs_pso_step call */
int eqs_pso_last_step_ms(eqs_pso *h, double *ms, int64_t *kernel_launches);
int eqs_pso_read_gbest(eqs_pso *h, double *g /*[n] or NULL*/, double *cost, int64_t *iters_run, int32_t *stalled,
                       int64_t *improvements);
/* particle-major [local][n] like the reference's Vec<Particle>; any pointer may be NULL */
int eqs_pso_read_state(eqs_pso *h, double *x, double *v, double *pbest, double *pbest_cost);
int eqs_pso_read_ring(eqs_pso *h, double *ring /*[50]*/, int64_t *ring_i);
/* first global index and number of the particles this rank holds; they are the block first .. first + local - 1 except
 * in the tile-cyclic distribution described below, where only `local` is meaningful */
int eqs_pso_local_range(eqs_pso *h, int64_t *first, int64_t *local);
/* global index of every local particle, in the row order of eqs_pso_read_state: gidx [local].  Contiguous blocks
 * (first .. first + local - 1) except in sequential-exact mode over several ranks, which deals tiles of 256 consecutive
 * particles round-robin: global tile T lives on rank T % world (the windows of the reference's particle order,
 * particle_swarm.rs:399-427, then keep every rank busy). */
int eqs_pso_local_indices(eqs_pso *h, int64_t *gidx);
/* split-phase protocol for EQS_EXCHANGE_MANUAL (virtual ranks; also how the NCCL path is sequenced):
 *   init: init_local(phase 0) -> gather records -> init_local(phase 1) -> gather records -> init_apply
 *   step: step_local -> gather records -> step_apply
 * "gather records" = eqs_pso_get_record on every rank, eqs_pso_set_records with all of them on every rank. */
int eqs_pso_init_local(eqs_pso *h, const double *x0, int32_t phase);
int eqs_pso_init_apply(eqs_pso *h);
int eqs_pso_step_local(eqs_pso *h, const double *replay_step);
int eqs_pso_step_apply(eqs_pso *h);
int eqs_pso_get_record(eqs_pso *h, double *record);                 /* [eqs_record_doubles(n)] */
int eqs_pso_set_records(eqs_pso *h, const double *records);         /* [world][eqs_record_doubles(n)] */

/* ---- CrossEntropy -------------------------------------------------------------------------- */
int eqs_ce_validate(const eqs_ce_params *p);
/* CrossEntropy::solve(x0), cross_entropy.rs:245-306.  out_mu [n] receives the final mean. */
int eqs_ce_solve(eqs_ctx *ctx, const eqs_ce_params *p, const double *x0, double *out_mu, eqs_report *report);

int eqs_ce_create(eqs_ctx *ctx, const eqs_ce_params *p, const double *x0, eqs_ce **out);
int eqs_ce_destroy(eqs_ce *h);
/* up to `iters` passes of the while body :256-302; replay_normals: NULL or z [iters][N][n] (:264-271) */
int eqs_ce_step(eqs_ce *h, int64_t iters, const double *replay_normals);
int eqs_ce_sync(eqs_ce *h);
int eqs_ce_last_step_ms(eqs_ce *h, double *ms, int64_t *kernel_launches);
int eqs_ce_read_state(eqs_ce *h, double *mu, double *sigma, int64_t *iters_run, int32_t *active);
/* elites of the last step: global sample indices ascending [k_local]; costs of the local samples */
int eqs_ce_read_elites(eqs_ce *h, int64_t *idx, int64_t *count);
int eqs_ce_read_costs(eqs_ce *h, double *costs);
int eqs_ce_local_range(eqs_ce *h, int64_t *first, int64_t *local);
/* split-phase protocol (EQS_EXCHANGE_MANUAL; also how the pass is sequenced over NCCL when the ranks cannot map each
 * other's memory -- with peer mapping, and on one rank, the exchanges happen inside the kernels).  One pass of the while
 * body is phases 0..9; after phase p the caller gathers eqs_ce_exchange_words(h, p) 8-byte words from every rank (get)
 * and hands all ranks' words back in rank order (set): phase 0 the key range and NaN count of the pass (u64[2048], words
 * 0..2 used), phases 1-6 the histogram of one level of the radix select (u64[2048]), phases 7-8 the elite moment sums
 * (f64[n]), phase 9 nothing.  replay_normals (phase 0 only): NULL or z [N][n] for this pass. */
int eqs_ce_phase(eqs_ce *h, int32_t phase, const double *replay_normals);
int64_t eqs_ce_exchange_words(eqs_ce *h, int32_t phase);
int eqs_ce_get_exchange(eqs_ce *h, int32_t phase, void *words);
int eqs_ce_set_exchange(eqs_ce *h, int32_t phase, const void *words /*[world][words]*/);

/* ---- MonteCarlo (first row of SURVEY.md §8f after the optimisers) --------------------------------
 * MonteCarlo::new(f).with_sample_count(..).integrate_with_rng(from, to, rng), reference
 * src/integrators/monte_carlo.rs:85-212 with MeanVariance::from_iterator_using_welford (src/lib.rs:172-202):
 * sample_count points uniform in [from, to] (default 1000, integrators/mod.rs:18), integrand by ID, mean and
 * Bessel-corrected variance of the integrand values scaled by the volume / its square. */
typedef struct {
    int32_t n;
    int32_t integrand_id;     /* eqs_objective */
    int64_t sample_count;     /* GLOBAL; >= 2 */
    const double *from;       /* [n] */
    const double *to;         /* [n] */
    uint64_t seed;            /* Philox key (replaces the `rng` argument) */
    const double *replay_u;   /* NULL, or u in [0,1) [sample_count][n] in draw order (parity mode) */
} eqs_mc_params;

int eqs_mc_validate(const eqs_mc_params *p); /* EXTERNAL for from > to / non-finite, INCORRECT_INPUT for count < 2 */
int eqs_mc_integrate(eqs_ctx *ctx, const eqs_mc_params *p, double *mean, double *variance, eqs_report *report);

/* ---- MISER, recursive stratified sampling ------------------------------------------------
 * MISER::new(f).with_*(..).integrate_with_rng(from, to, rng), reference src/integrators/miser.rs:270-443
 * (miser_recurse :301-443).  Per region: leaf (plain Monte Carlo, population variance, :335-349) when the depth is
 * used up or fewer than minimum_dimensional_sample_count * n points are allotted; otherwise 2n variance probes of
 * sample_count points each (:362-400), bisection of the dimension with the smallest total variance at a dithered
 * mid-point, points split by variance^(1/(1+alpha)) (:404-420), recursion on both halves, results added (:442).
 * The recursion is evaluated level by level on one GPU: every probe and leaf of a level is one Welford batch of a
 * single kernel launch.  Defaults: miser.rs:13-23.  Not sharded over ranks (the levels are latency-bound). */
typedef struct {
    int32_t n;
    int32_t integrand_id;     /* eqs_objective */
    int64_t sample_count;     /* default 1000, integrators/mod.rs:18; at most 2^40 */
    int64_t minimum_dimensional_sample_count; /* default 32 */
    int32_t maximum_depth;    /* default 5; at most 30 */
    int32_t reserved;
    double alpha;             /* default 2 */
    double dither;            /* default 0.05 */
    const double *from;       /* [n] */
    const double *to;         /* [n] */
    uint64_t seed;            /* Philox key (replaces the `rng` argument) */
} eqs_miser_params;

/* INCORRECT_INPUT for n <= 0, an unknown / non-streaming integrand, depth or count out of range.  Everything the
 * reference reports while it runs comes from eqs_miser_integrate: EXTERNAL for from > to, non-finite bounds, a bad
 * dither (Uniform::new_inclusive, :296-299), TYPE_CONVERSION when a variance share is not a count (:413-419, e.g. a
 * constant integrand: 0/0), INCORRECT_INPUT when a leaf holds fewer than 2 points (lib.rs:190-193). */
int eqs_miser_validate(const eqs_miser_params *p);
int eqs_miser_integrate(eqs_ctx *ctx, const eqs_miser_params *p, double *mean, double *variance, eqs_report *report);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* EQSOLVER_B200_H */
