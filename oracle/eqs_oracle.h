/*
 * eqs_oracle.h -- CPU ORACLE for the population step of eqsolver's global optimisers.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under eqsolver_b200/ (the product) may include, link,
 * import or execute this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and there only as the checker or the reported CPU baseline.
 *
 * PARITY UNPINNED: the reference (AzeezDa/eqsolver v0.4.1, pure Rust) cannot be compiled in this
 * image (no cargo/rustc), draws all randomness from an unseedable thread-local `rand::rng()`
 * (particle_swarm.rs:367,373,407-408; cross_entropy.rs:270) and its own tests assert nothing about
 * ParticleSwarm / CrossEntropy results.  There are therefore no reference golden vectors for this
 * path.  What pins this restatement instead: the objective known-answer values and the Random123
 * Philox4x32-10 vectors in tests/golden/, the known optimum of the three-circles problem
 * (reference tests/multi.rs:70), and line-by-line citation of the Rust source below.
 *
 * Third-party arithmetic on the reference path that is NOT under /root/reference (Cargo.toml:21-26,
 * no Cargo.lock): rand 0.10.1 `rng()`, rand_distr 0.6.0 `Uniform::new_inclusive(lo,hi).sample` =
 * u*scale + lo and `Normal::new(mu,sigma).sample` = mu + sigma*z, nalgebra 0.35 `+=`, indexing,
 * `apply_norm(&UniformNorm)` = max |.|.  The oracle injects draws AFTER the bit generator
 * (u in [0,1) and z ~ N(0,1)), so the generators' internals drop out of parity.
 */
#ifndef EQS_ORACLE_H
#define EQS_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Objective IDs (shared numbering with include/eqsolver_b200.h; restated, not included). */
enum {
    ORC_OBJ_SPHERE = 0,
    ORC_OBJ_ROSENBROCK = 1,
    ORC_OBJ_RASTRIGIN = 2,      /* reference examples/global_optimisation.rs:9-17 */
    ORC_OBJ_ACKLEY = 3,
    ORC_OBJ_THREE_CIRCLES = 4,  /* reference examples/global_optimiser_as_solver.rs:22-42 (n = 2) */
    ORC_OBJ_HEAVY = 5,          /* reference benchmarks/multi.rs:106-115,137 */
    ORC_OBJ_USER = 6,           /* twin of the shipped user-hook example (Zakharov); not in the reference */
    ORC_OBJ_GAUSSIAN = 7,       /* exp(-sum x^2): reference tests/integrators.rs:11-13 for n = 1 */
    ORC_OBJ_SINCOS = 8,         /* sin(cos(x1) + x0), n = 2: reference tests/integrators.rs:15-17 */
    ORC_OBJ_SERIES = 9          /* sum_{i<1000} (-1)^i x^i per dimension: reference benchmarks/integrators.rs:69-75 */
};

/* Status codes = SolverError variants, reference src/lib.rs:19-44. */
enum {
    ORC_OK = 0,
    ORC_ERR_MAX_ITER = 1,
    ORC_ERR_NAN = 2,
    ORC_ERR_INCORRECT_INPUT = 3,
    ORC_ERR_BAD_JACOBIAN = 4,
    ORC_ERR_TYPE_CONVERSION = 5,
    ORC_ERR_EXTERNAL = 6
};

enum { ORC_PSO_SEQUENTIAL = 0, ORC_PSO_SYNCHRONOUS = 1 };
enum { ORC_CE_DIV_REFERENCE_TEN = 0, ORC_CE_DIV_ELITE_COUNT = 1 };

#define ORC_STALL_WINDOW 50 /* DEFAULT_STALL_ITERATIONS, particle_swarm.rs:12 */

/* ---- building blocks ---------------------------------------------------------------------- */
double orc_objective(int id, const double *x, int n);
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* u = ((w_hi mod 2^20) * 2^32 + w_lo) * 2^-52 in [0,1), the draw contract of DESIGN.md */
double orc_u01(uint32_t w_lo, uint32_t w_hi);
/* fill helpers used by tests to cross-check the device generator */
void orc_pso_init_draws(uint64_t seed, int64_t gidx, int n, double *pos_u, double *vel_u);
void orc_pso_step_draws(uint64_t seed, int64_t t, int64_t gidx, int n, double *rp, double *rg);
void orc_ce_normals(uint64_t seed, int64_t t, int64_t gidx, int n, double *z);

/* ---- ParticleSwarm ------------------------------------------------------------------------ */
typedef struct {
    int32_t n;
    int32_t objective_id;
    int32_t mode;            /* ORC_PSO_SEQUENTIAL = the reference; SYNCHRONOUS = one gbest update/iter */
    int32_t _pad;
    int64_t particle_count;  /* global N */
    int64_t first_index;     /* shard: global index of local particle 0 */
    int64_t local_count;     /* shard size; == particle_count when unsharded */
    int64_t iter_max;
    double tol, w, c1, c2;
    const double *lower, *upper;
    uint64_t seed;
    /* replay: u in [0,1).  init: [N][2n] (n position draws then n velocity draws per particle,
     * particle_swarm.rs:363-373); step: [iters][N][n][2] (rp, rg per dim, :407-408).  Indexed by
     * GLOBAL particle index.  NULL => host Philox4x32-10 keyed by seed. */
    const double *replay_init;
    const double *replay_step;
    int64_t replay_step_iters;
} orc_pso_params;

typedef struct orc_pso orc_pso;

int orc_pso_validate(const orc_pso_params *p); /* ParticleSwarm::new, particle_swarm.rs:119-153 */
orc_pso *orc_pso_create(const orc_pso_params *p);
void orc_pso_destroy(orc_pso *s);
/* unsharded, exact reference order: particle_swarm.rs:357-389 */
void orc_pso_init(orc_pso *s, const double *x0);
/* one pass of the outer loop body :395-427.  returns 1 if the stall test broke the loop (no-op). */
int orc_pso_step(orc_pso *s);
/* whole solve :356-431; out_x = gbest position.  returns status. */
int orc_pso_solve(const orc_pso_params *p, const double *x0, double *out_x, int64_t *iters_run,
                  double *gbest_cost);

/* sharded protocol pieces (SYNCHRONOUS mode; model of the multi-GPU exchange).
 * record layout (doubles): [0]=cost, [1]=global index (as double, exact < 2^53), [2..2+n) = x */
void orc_pso_init_local(orc_pso *s, const double *x0, double *record);
/* record-low scan of the local init costs against `seed_cost` (= min over x0 and all lower
 * ranks): writes the number of ring pushes and the last <=50 pushed values in order. */
void orc_pso_init_ring_local(orc_pso *s, double seed_cost, int64_t *pushes, double *last50);
/* merge: records [world][2+n], pushes[world], last50 [world][50] -> G, Gc, ring on this shard */
void orc_pso_init_apply(orc_pso *s, int world, const double *records, const int64_t *pushes,
                        const double *last50);
int orc_pso_step_local(orc_pso *s, double *record); /* returns 1 if stalled (no-op) */
void orc_pso_step_apply(orc_pso *s, int world, const double *records);

/* state readers, particle-major [i][d] like the reference's Vec<Particle> */
int64_t orc_pso_iters(const orc_pso *s);
double orc_pso_gbest_cost(const orc_pso *s);
void orc_pso_read_gbest(const orc_pso *s, double *g);
void orc_pso_read_state(const orc_pso *s, double *x, double *v, double *pb, double *pbc);
void orc_pso_read_ring(const orc_pso *s, double *ring, int64_t *ring_i);
int64_t orc_pso_last_improvements(const orc_pso *s); /* gbest improvements in the last step */

/* ---- CrossEntropy ------------------------------------------------------------------------- */
typedef struct {
    int32_t n;
    int32_t objective_id;
    int32_t mean_divisor;    /* REFERENCE_TEN = literal 10 of cross_entropy.rs:287; ELITE_COUNT = k */
    int32_t _pad;
    int64_t sample_size;     /* global N */
    int64_t first_index, local_count;
    int64_t elite_count;     /* importance_selection_size k */
    int64_t iter_max;
    double tol;
    const double *std_dev;   /* NULL => ones, cross_entropy.rs:247-251 */
    uint64_t seed;
    const double *replay_normals; /* z ~ N(0,1), [iters][N][n] in the reference's draw order :264-271 */
    int64_t replay_iters;
} orc_ce_params;

typedef struct orc_ce orc_ce;

int orc_ce_validate(const orc_ce_params *p);
orc_ce *orc_ce_create(const orc_ce_params *p, const double *x0);
void orc_ce_destroy(orc_ce *s);
/* one pass of the while body cross_entropy.rs:256-302.  returns 0 = ran, 1 = loop condition false
 * (no-op), <0 = -status where the reference would panic (NaN cost, non-finite sigma). */
int orc_ce_step(orc_ce *s);
int orc_ce_solve(const orc_ce_params *p, const double *x0, double *out_mu, int64_t *iters_run);
int64_t orc_ce_iters(const orc_ce *s);
void orc_ce_read_state(const orc_ce *s, double *mu, double *sigma);
/* elite global indices of the last step, ascending by (cost, index); and the costs of all local samples */
void orc_ce_read_elites(const orc_ce *s, int64_t *idx);
void orc_ce_read_costs(const orc_ce *s, double *costs);

/* sharded protocol pieces for CE */
void orc_ce_sample_local(orc_ce *s);                      /* sample + evaluate the local block */
/* apply a global elite decision: threshold cost T, strict-less all elite, `tie_take` of the local
 * samples with cost == T (lowest index first).  Accumulates local sum of x per dim. */
void orc_ce_elite_sum_local(orc_ce *s, double T, int64_t tie_take, double *sum_x);
void orc_ce_elite_sqsum_local(orc_ce *s, const double *mu_new, double *sum_sq);
void orc_ce_apply(orc_ce *s, const double *mu_new, const double *sigma_new);
int orc_ce_active(const orc_ce *s); /* loop condition :255 */

/* timing helper for bench.py's cpu_baseline: runs `iters` PSO steps, returns seconds */
/* ---- MonteCarlo -------------------------------------------------------
 * MonteCarlo::integrate_with_rng over [from, to], reference src/integrators/monte_carlo.rs:164-212:
 * `count` points uniform in the box (Uniform::new_inclusive per dimension, :191-195), integrand values
 * through MeanVariance::from_iterator_using_welford(.., bessel = true) (src/lib.rs:172-202), mean scaled
 * by the volume and variance by its square (:154-157, lib.rs:143-148).  replay_u: NULL (host Philox,
 * stream 3, one block per dimension pair) or u in [0,1) [count][n] in draw order.  returns a status. */
int orc_mc_integrate(int integrand_id, int n, int64_t count, const double *from, const double *to, uint64_t seed,
                     const double *replay_u, double *mean, double *variance);
void orc_mc_draws(uint64_t seed, int64_t index, int n, double *u);
/* one rank's share of a sharded MonteCarlo: moments = {count, mean, sum of squared deviations} of points [first, first+count) */
int orc_mc_moments(int integrand_id, int n, int64_t first, int64_t count, const double *from, const double *to,
                   uint64_t seed, double *moments);

/* ---- MISER (src/integrators/miser.rs:270-443), recursion exactly as miser_recurse :301-443; the draw counters are
 * described in eqs_oracle.c.  evals: integrand evaluations spent (may be NULL).  returns a status. */
int orc_miser_integrate(int integrand_id, int n, int64_t sample_count, int maximum_depth, int64_t min_dim_count,
                        double alpha, double dither, const double *from, const double *to, uint64_t seed,
                        double *mean, double *variance, int64_t *evals);

double orc_time_pso_steps(orc_pso *s, int64_t iters);
/* all-cores variant of a synchronous pass (threads over particles), bit-identical to orc_pso_step */
int orc_pso_step_mt(orc_pso *s, int threads);
double orc_time_pso_steps_mt(orc_pso *s, int64_t iters, int threads);
double orc_time_ce_steps(orc_ce *s, int64_t iters);

#ifdef __cplusplus
}
#endif
#endif
