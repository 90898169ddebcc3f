/*
 * eqs_oracle.c -- CPU restatement of eqsolver's ParticleSwarm / CrossEntropy population step.
 * TEST INFRASTRUCTURE ONLY; PARITY UNPINNED (see eqs_oracle.h for both statements in full).
 *
 * Follows /root/reference (AzeezDa/eqsolver v0.4.1):
 *   src/solvers/global_optimisers/particle_swarm.rs:119-153 (samplers), :24-46 (ring),
 *     :356-431 (solve), :433-441 (stall test)
 *   src/solvers/global_optimisers/cross_entropy.rs:245-306 (solve)
 *   constants particle_swarm.rs:8-12, cross_entropy.rs:8-9, lib.rs:13,16
 * Build with -O2 -ffp-contract=off: rustc never contracts a*b+c into an FMA.
 */
#include "eqs_oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ============================== objectives ================================================== */

static const double ORC_PI = 3.14159265358979323846; /* std::f64::consts::PI */

/* x.powi(2) lowers to x*x. */
static double sq(double x) { return x * x; }

/* f64::powi with a run-time exponent = compiler-rt's __powidf2: square-and-multiply from the low bit up */
static double powi(double a, int b)
{
    const int recip = b < 0;
    double r = 1.0;
    for (;;) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return recip ? 1.0 / r : r;
}

double orc_objective(int id, const double *x, int n)
{
    switch (id) {
    case ORC_OBJ_SPHERE: {
        double total = 0.0;
        for (int d = 0; d < n; ++d) total += x[d] * x[d];
        return total;
    }
    case ORC_OBJ_ROSENBROCK: {
        double total = 0.0;
        for (int d = 0; d + 1 < n; ++d) {
            double t = x[d + 1] - x[d] * x[d];
            double u = 1.0 - x[d];
            total += 100.0 * (t * t) + u * u;
        }
        return total;
    }
    case ORC_OBJ_RASTRIGIN: {
        /* examples/global_optimisation.rs:9-17: total = 10*SIZE; total += w*w - 10*cos(2*PI*w) */
        double total = 10.0 * (double)n;
        for (int d = 0; d < n; ++d) {
            double w = x[d];
            total += w * w - 10.0 * cos(2.0 * ORC_PI * w);
        }
        return total;
    }
    case ORC_OBJ_ACKLEY: {
        double s2 = 0.0, sc = 0.0;
        for (int d = 0; d < n; ++d) {
            s2 += x[d] * x[d];
            sc += cos(2.0 * ORC_PI * x[d]);
        }
        double a = -20.0 * exp(-0.2 * sqrt(s2 / (double)n));
        double b = exp(sc / (double)n);
        return a - b + 20.0 + 2.718281828459045235360287; /* std::f64::consts::E */
    }
    case ORC_OBJ_THREE_CIRCLES: {
        /* examples/global_optimiser_as_solver.rs:22-42; objective = ||F(v) - b||, n must be 2 */
        static const double c[3][3] = {{3., 5., 3.}, {1., 0., 4.}, {6., 2., 2.}};
        double acc = 0.0;
        for (int k = 0; k < 3; ++k) {
            double e = (sq(x[0] - c[k][0]) + sq(x[1] - c[k][1])) - sq(c[k][2]);
            acc += e * e;
        }
        return sqrt(acc);
    }
    case ORC_OBJ_HEAVY: {
        /* benchmarks/multi.rs:106-115: out_i = sum_{j>=i} x_j^2 (fresh forward sum per i); :137 norm */
        double acc = 0.0;
        for (int i = 0; i < n; ++i) {
            double o = 0.0;
            for (int j = i; j < n; ++j) o += x[j] * x[j];
            acc += o * o;
        }
        return sqrt(acc);
    }
    case ORC_OBJ_USER: {
        /* twin of the shipped user hook eqsolver_b200/csrc/user_objective.cuh (Zakharov); not in the reference */
        double s1 = 0.0, s2 = 0.0;
        for (int d = 0; d < n; ++d) {
            s1 += x[d] * x[d];
            s2 += (0.5 * (double)(d + 1)) * x[d];
        }
        double q = s2 * s2;
        return s1 + q + q * q;
    }
    case ORC_OBJ_GAUSSIAN: { /* tests/integrators.rs:11-13: (-x * x).exp(); summed over the dimensions for n > 1 */
        double s2 = 0.0;
        for (int d = 0; d < n; ++d) s2 += x[d] * x[d];
        return exp(-s2);
    }
    case ORC_OBJ_SINCOS: /* tests/integrators.rs:15-17: (v[1].cos() + v[0]).sin() */
        return sin(cos(x[1]) + x[0]);
    case ORC_OBJ_SERIES: { /* benchmarks/integrators.rs:69-75: sum += (-1f64).powi(i) * x.powi(i), i in 0..1000 */
        double total = 0.0;
        for (int d = 0; d < n; ++d) {
            double sum = 0.0;
            for (int i = 0; i < 1000; ++i) sum += powi(-1.0, i) * powi(x[d], i);
            total += sum;
        }
        return total;
    }
    default:
        return NAN;
    }
}

/* ============================== Philox4x32-10 =============================================== */
/* Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC'11); Random123 constants. */

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

double orc_u01(uint32_t w_lo, uint32_t w_hi)
{
    /* 52 random mantissa bits under exponent 0, minus one: what rand's Uniform<f64> does
     * (into_float_with_exponent(0) - 1.0), so u = k * 2^-52 with k = (w_hi mod 2^20) * 2^32 + w_lo */
    uint64_t k = ((uint64_t)(w_hi & 0xFFFFFu) << 32) | (uint64_t)w_lo; /* 52 bits */
    return (double)k * 0x1.0p-52;
}

/* Draw-index contract: key = (seed lo, seed hi);
 * ctr = (gidx lo, gidx hi, stream<<28 | dim-or-pair index, iteration). */
enum { STREAM_PSO_INIT = 0, STREAM_PSO_STEP = 1, STREAM_CE = 2, STREAM_MC = 3, STREAM_MISER = 4 };

static void philox_draw(uint64_t seed, int stream, int64_t t, int64_t gidx, uint32_t j, uint32_t out[4])
{
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t ctr[4] = {(uint32_t)(uint64_t)gidx, (uint32_t)((uint64_t)gidx >> 32),
                       ((uint32_t)stream << 28) | j, (uint32_t)t};
    orc_philox4x32_10(ctr, key, out);
}

void orc_pso_init_draws(uint64_t seed, int64_t gidx, int n, double *pos_u, double *vel_u)
{
    for (int d = 0; d < n; ++d) {
        uint32_t w[4];
        philox_draw(seed, STREAM_PSO_INIT, 0, gidx, (uint32_t)d, w);
        pos_u[d] = orc_u01(w[0], w[1]);
        vel_u[d] = orc_u01(w[2], w[3]);
    }
}

void orc_pso_step_draws(uint64_t seed, int64_t t, int64_t gidx, int n, double *rp, double *rg)
{
    for (int d = 0; d < n; ++d) {
        uint32_t w[4];
        philox_draw(seed, STREAM_PSO_STEP, t, gidx, (uint32_t)d, w);
        rp[d] = orc_u01(w[0], w[1]);
        rg[d] = orc_u01(w[2], w[3]);
    }
}

/* sin(pi*t), cos(pi*t) for t in [0,2): exact quadrant reduction, then libm on |r| <= 1/4. */
static void sincospi_host(double t, double *s, double *c)
{
    double k = nearbyint(2.0 * t);       /* 0..4 */
    double r = t - 0.5 * k;              /* exact, in [-1/4, 1/4] */
    double sr = sin(ORC_PI * r), cr = cos(ORC_PI * r);
    switch (((int)k) & 3) {
    case 0: *s = sr;  *c = cr;  break;
    case 1: *s = cr;  *c = -sr; break;
    case 2: *s = -sr; *c = -cr; break;
    default: *s = -cr; *c = sr; break;
    }
}

/* Box-Muller on one Philox block per PAIR of dimensions: z[2p] = r cos(2 pi u2), z[2p+1] = r sin(2 pi u2),
 * r = sqrt(-2 ln(1-u1)).  Stands in for rand_distr's StandardNormal (ziggurat), cross_entropy.rs:270. */
void orc_ce_normals(uint64_t seed, int64_t t, int64_t gidx, int n, double *z)
{
    for (int p = 0; 2 * p < n; ++p) {
        uint32_t w[4];
        philox_draw(seed, STREAM_CE, t, gidx, (uint32_t)p, w);
        double u1 = orc_u01(w[0], w[1]);
        double u2 = orc_u01(w[2], w[3]);
        double r = sqrt(-2.0 * log(1.0 - u1));
        double s, c;
        sincospi_host(2.0 * u2, &s, &c);
        z[2 * p] = r * c;
        if (2 * p + 1 < n) z[2 * p + 1] = r * s;
    }
}

/* ============================== ParticleSwarm =============================================== */

struct orc_pso {
    orc_pso_params p;
    double *lower, *upper;
    double *x, *v, *pb, *pbc; /* particle-major [i][d], like Vec<Particle> (particle_swarm.rs:14-22) */
    double *cost;             /* costs of the last evaluation pass */
    double *G;
    double Gc;
    double ring[ORC_STALL_WINDOW]; /* CircularArray<T,50>, particle_swarm.rs:24-46 */
    int64_t ring_i;
    int64_t iters;            /* outer-loop passes that ran */
    int64_t last_improvements;
    double *rp, *rg;          /* scratch draws for one particle */
};

static void ring_push(orc_pso *s, double v)
{
    s->ring[s->ring_i] = v;
    s->ring_i = (s->ring_i + 1) % ORC_STALL_WINDOW;
}

/* particle_swarm.rs:433-441 */
static int stalled_too_long(const orc_pso *s)
{
    for (int j = 0; j < ORC_STALL_WINDOW; ++j)
        if (!(fabs(s->ring[j] - s->Gc) < s->p.tol)) return 0;
    return 1;
}

int orc_pso_validate(const orc_pso_params *p)
{
    if (p->n <= 0 || p->particle_count < 0 || p->iter_max < 0) return ORC_ERR_INCORRECT_INPUT;
    /* Uniform::new_inclusive fails (=> SolverError::ExternalError) unless lo <= hi and both finite,
     * particle_swarm.rs:124-139 */
    for (int d = 0; d < p->n; ++d) {
        double lo = p->lower[d], hi = p->upper[d];
        if (!(lo <= hi) || !isfinite(lo) || !isfinite(hi) || !isfinite(hi - lo)) return ORC_ERR_EXTERNAL;
    }
    return ORC_OK;
}

orc_pso *orc_pso_create(const orc_pso_params *p)
{
    if (orc_pso_validate(p) != ORC_OK) return NULL;
    orc_pso *s = (orc_pso *)calloc(1, sizeof(orc_pso));
    s->p = *p;
    int n = p->n;
    int64_t L = p->local_count;
    s->lower = (double *)malloc(sizeof(double) * n);
    s->upper = (double *)malloc(sizeof(double) * n);
    memcpy(s->lower, p->lower, sizeof(double) * n);
    memcpy(s->upper, p->upper, sizeof(double) * n);
    s->p.lower = s->lower;
    s->p.upper = s->upper;
    size_t m = (size_t)(L > 0 ? L : 1) * n;
    s->x = (double *)malloc(sizeof(double) * m);
    s->v = (double *)malloc(sizeof(double) * m);
    s->pb = (double *)malloc(sizeof(double) * m);
    s->pbc = (double *)malloc(sizeof(double) * (size_t)(L > 0 ? L : 1));
    s->cost = (double *)malloc(sizeof(double) * (size_t)(L > 0 ? L : 1));
    s->G = (double *)malloc(sizeof(double) * n);
    s->rp = (double *)malloc(sizeof(double) * n);
    s->rg = (double *)malloc(sizeof(double) * n);
    for (int j = 0; j < ORC_STALL_WINDOW; ++j) s->ring[j] = INFINITY;
    return s;
}

void orc_pso_destroy(orc_pso *s)
{
    if (!s) return;
    free(s->lower); free(s->upper); free(s->x); free(s->v); free(s->pb); free(s->pbc);
    free(s->cost); free(s->G); free(s->rp); free(s->rg); free(s);
}

/* samplers of particle_swarm.rs:124-139 applied to u in [0,1): value = u*scale + low */
static void sample_particle(orc_pso *s, int64_t li)
{
    const orc_pso_params *p = &s->p;
    int n = p->n;
    int64_t gi = p->first_index + li;
    if (p->replay_init) {
        memcpy(s->rp, p->replay_init + (size_t)gi * 2 * n, sizeof(double) * n);
        memcpy(s->rg, p->replay_init + (size_t)gi * 2 * n + n, sizeof(double) * n);
    } else {
        orc_pso_init_draws(p->seed, gi, n, s->rp, s->rg);
    }
    double *x = s->x + (size_t)li * n, *v = s->v + (size_t)li * n;
    for (int d = 0; d < n; ++d) {
        double lo = s->lower[d], hi = s->upper[d];
        x[d] = s->rp[d] * (hi - lo) + lo;             /* :363-367 */
    }
    for (int d = 0; d < n; ++d) {
        double dist = fabs(s->upper[d] - s->lower[d]); /* :135 */
        v[d] = s->rg[d] * (dist + dist) + (-dist);     /* :369-373 */
    }
}

void orc_pso_init(orc_pso *s, const double *x0)
{
    const orc_pso_params *p = &s->p;
    int n = p->n;
    memcpy(s->G, x0, sizeof(double) * n);              /* :357 */
    s->Gc = orc_objective(p->objective_id, x0, n);     /* :358 */
    for (int j = 0; j < ORC_STALL_WINDOW; ++j) s->ring[j] = INFINITY; /* :359 */
    s->ring_i = 0;
    s->iters = 0;
    for (int64_t i = 0; i < p->local_count; ++i) {     /* :361-389 */
        sample_particle(s, i);
        double *x = s->x + (size_t)i * n;
        double cost = orc_objective(p->objective_id, x, n);
        if (cost < s->Gc) {                            /* :376-380 */
            ring_push(s, s->Gc);
            memcpy(s->G, x, sizeof(double) * n);
            s->Gc = cost;
        }
        memcpy(s->pb + (size_t)i * n, x, sizeof(double) * n);
        s->pbc[i] = cost;
        s->cost[i] = cost;
    }
}

static void fetch_step_draws(const orc_pso *s, int64_t li, double *rp, double *rg)
{
    const orc_pso_params *p = &s->p;
    int n = p->n;
    int64_t gi = p->first_index + li;
    if (p->replay_step) {
        const double *src = p->replay_step + ((size_t)s->iters * p->particle_count + gi) * 2 * n;
        for (int d = 0; d < n; ++d) { rp[d] = src[2 * d]; rg[d] = src[2 * d + 1]; }
    } else {
        orc_pso_step_draws(p->seed, s->iters, gi, n, rp, rg);
    }
}

/* velocity + position update and evaluation of one particle against gbest G, :406-415 */
static double move_particle_with(orc_pso *s, int64_t li, double *rps, double *rgs)
{
    const orc_pso_params *p = &s->p;
    int n = p->n;
    double *x = s->x + (size_t)li * n, *v = s->v + (size_t)li * n, *pb = s->pb + (size_t)li * n;
    fetch_step_draws(s, li, rps, rgs);
    for (int d = 0; d < n; ++d) {
        double rp = rps[d], rg = rgs[d];
        v[d] = p->w * v[d] + p->c1 * rp * (pb[d] - x[d]) + p->c2 * rg * (s->G[d] - x[d]); /* :409-411 */
    }
    for (int d = 0; d < n; ++d) x[d] += v[d];          /* :413 */
    return orc_objective(p->objective_id, x, n);       /* :415 */
}

static double move_particle(orc_pso *s, int64_t li) { return move_particle_with(s, li, s->rp, s->rg); }

static int step_sequential(orc_pso *s)
{
    const orc_pso_params *p = &s->p;
    int n = p->n;
    if (stalled_too_long(s)) return 1;                 /* :395-397 */
    s->last_improvements = 0;
    for (int64_t i = 0; i < p->local_count; ++i) {     /* :399-427 */
        double cost = move_particle(s, i);
        s->cost[i] = cost;
        if (cost < s->pbc[i]) {                        /* :417-419 */
            memcpy(s->pb + (size_t)i * n, s->x + (size_t)i * n, sizeof(double) * n);
            s->pbc[i] = cost;
            if (s->pbc[i] < s->Gc) {                   /* :420-425 */
                ring_push(s, s->Gc);
                memcpy(s->G, s->x + (size_t)i * n, sizeof(double) * n);
                s->Gc = s->pbc[i];
                s->last_improvements++;
            }
        }
    }
    s->iters++;
    return 0;
}

/* NaN never wins a `<`; give it +inf so min-loc reductions agree on every implementation */
static double sanitize(double c) { return (c != c) ? INFINITY : c; }

static void local_minloc(const orc_pso *s, double *record)
{
    int n = s->p.n;
    double best = INFINITY;
    int64_t bi = -1;
    for (int64_t i = 0; i < s->p.local_count; ++i) {
        double c = sanitize(s->cost[i]);
        if (bi < 0 || c < best) { best = c; bi = i; }
    }
    record[0] = best;
    record[1] = bi < 0 ? -1.0 : (double)(s->p.first_index + bi);
    for (int d = 0; d < n; ++d) record[2 + d] = bi < 0 ? 0.0 : s->x[(size_t)bi * n + d];
}

int orc_pso_step_local(orc_pso *s, double *record)
{
    const orc_pso_params *p = &s->p;
    int n = p->n;
    if (stalled_too_long(s)) return 1;
    for (int64_t i = 0; i < p->local_count; ++i) {
        double cost = move_particle(s, i);
        s->cost[i] = cost;
        if (cost < s->pbc[i]) {
            memcpy(s->pb + (size_t)i * n, s->x + (size_t)i * n, sizeof(double) * n);
            s->pbc[i] = cost;
        }
    }
    local_minloc(s, record);
    return 0;
}

static const double *best_record(int world, const double *records, int n)
{
    const double *best = NULL;
    for (int r = 0; r < world; ++r) {
        const double *rec = records + (size_t)r * (2 + n);
        if (rec[1] < 0) continue;
        if (!best || rec[0] < best[0] || (rec[0] == best[0] && rec[1] < best[1])) best = rec;
    }
    return best;
}

void orc_pso_step_apply(orc_pso *s, int world, const double *records)
{
    int n = s->p.n;
    const double *best = best_record(world, records, n);
    s->last_improvements = 0;
    if (best && best[0] < s->Gc) {
        ring_push(s, s->Gc);
        memcpy(s->G, best + 2, sizeof(double) * n);
        s->Gc = best[0];
        s->last_improvements = 1;
    }
    s->iters++;
}

int orc_pso_step(orc_pso *s)
{
    if (s->p.mode == ORC_PSO_SEQUENTIAL) return step_sequential(s);
    double *record = (double *)malloc(sizeof(double) * (2 + s->p.n));
    int r = orc_pso_step_local(s, record);
    if (!r) orc_pso_step_apply(s, 1, record);
    free(record);
    return r;
}

int orc_pso_solve(const orc_pso_params *p, const double *x0, double *out_x, int64_t *iters_run,
                  double *gbest_cost)
{
    int st = orc_pso_validate(p);
    if (st != ORC_OK) return st;
    orc_pso *s = orc_pso_create(p);
    orc_pso_init(s, x0);
    for (int64_t it = 0; it < p->iter_max; ++it)       /* :394 */
        if (orc_pso_step(s)) break;
    memcpy(out_x, s->G, sizeof(double) * p->n);        /* :430 */
    if (iters_run) *iters_run = s->iters;
    if (gbest_cost) *gbest_cost = s->Gc;
    orc_pso_destroy(s);
    return ORC_OK;
}

void orc_pso_init_local(orc_pso *s, const double *x0, double *record)
{
    const orc_pso_params *p = &s->p;
    int n = p->n;
    memcpy(s->G, x0, sizeof(double) * n);
    s->Gc = orc_objective(p->objective_id, x0, n);
    for (int j = 0; j < ORC_STALL_WINDOW; ++j) s->ring[j] = INFINITY;
    s->ring_i = 0;
    s->iters = 0;
    for (int64_t i = 0; i < p->local_count; ++i) {
        sample_particle(s, i);
        double cost = orc_objective(p->objective_id, s->x + (size_t)i * n, n);
        memcpy(s->pb + (size_t)i * n, s->x + (size_t)i * n, sizeof(double) * n);
        s->pbc[i] = cost;
        s->cost[i] = cost;
    }
    local_minloc(s, record);
}

void orc_pso_init_ring_local(orc_pso *s, double seed_cost, int64_t *pushes, double *last50)
{
    double running = seed_cost;
    int64_t k = 0;
    double tmp[ORC_STALL_WINDOW];
    for (int64_t i = 0; i < s->p.local_count; ++i) {
        double c = s->cost[i];
        if (c < running) { tmp[k % ORC_STALL_WINDOW] = running; running = c; ++k; }
    }
    int64_t kept = k < ORC_STALL_WINDOW ? k : ORC_STALL_WINDOW;
    for (int64_t j = 0; j < kept; ++j) last50[j] = tmp[(k - kept + j) % ORC_STALL_WINDOW];
    for (int64_t j = kept; j < ORC_STALL_WINDOW; ++j) last50[j] = INFINITY;
    *pushes = k;
}

void orc_pso_init_apply(orc_pso *s, int world, const double *records, const int64_t *pushes,
                        const double *last50)
{
    int n = s->p.n;
    const double *best = best_record(world, records, n);
    if (best && best[0] < s->Gc) {
        memcpy(s->G, best + 2, sizeof(double) * n);
        s->Gc = best[0];
    }
    int64_t g = 0;
    for (int r = 0; r < world; ++r) {
        int64_t kept = pushes[r] < ORC_STALL_WINDOW ? pushes[r] : ORC_STALL_WINDOW;
        g += pushes[r] - kept;
        for (int64_t j = 0; j < kept; ++j, ++g)
            s->ring[g % ORC_STALL_WINDOW] = last50[(size_t)r * ORC_STALL_WINDOW + j];
    }
    s->ring_i = g % ORC_STALL_WINDOW;
}

int64_t orc_pso_iters(const orc_pso *s) { return s->iters; }
double orc_pso_gbest_cost(const orc_pso *s) { return s->Gc; }
void orc_pso_read_gbest(const orc_pso *s, double *g) { memcpy(g, s->G, sizeof(double) * s->p.n); }
int64_t orc_pso_last_improvements(const orc_pso *s) { return s->last_improvements; }

void orc_pso_read_state(const orc_pso *s, double *x, double *v, double *pb, double *pbc)
{
    size_t m = (size_t)s->p.local_count * s->p.n;
    if (x) memcpy(x, s->x, sizeof(double) * m);
    if (v) memcpy(v, s->v, sizeof(double) * m);
    if (pb) memcpy(pb, s->pb, sizeof(double) * m);
    if (pbc) memcpy(pbc, s->pbc, sizeof(double) * (size_t)s->p.local_count);
}

void orc_pso_read_ring(const orc_pso *s, double *ring, int64_t *ring_i)
{
    memcpy(ring, s->ring, sizeof(s->ring));
    if (ring_i) *ring_i = s->ring_i;
}

/* ============================== CrossEntropy ================================================ */

struct orc_ce {
    orc_ce_params p;
    double *mu, *sigma;
    double *xs;     /* [local][n] samples of the last step (the reference materialises them too) */
    double *costs;  /* [local] */
    int64_t *order; /* local sample order by (cost, index) */
    int64_t *elite; /* [k] global indices of the elites of the last unsharded step */
    int64_t iter;   /* the reference's `iter`, starts at 1 (cross_entropy.rs:253) */
    int64_t local_elites; /* sharded protocol: elites held by this shard after elite_sum_local */
    double *z;
};

int orc_ce_validate(const orc_ce_params *p)
{
    if (p->n <= 0 || p->sample_size <= 0) return ORC_ERR_INCORRECT_INPUT;
    /* the reference does not validate k; k<2 divides by zero (:295), k>N silently takes N (:284).
     * Both implementations reject those instead. */
    if (p->elite_count < 2 || p->elite_count > p->sample_size) return ORC_ERR_INCORRECT_INPUT;
    return ORC_OK;
}

orc_ce *orc_ce_create(const orc_ce_params *p, const double *x0)
{
    if (orc_ce_validate(p) != ORC_OK) return NULL;
    orc_ce *s = (orc_ce *)calloc(1, sizeof(orc_ce));
    s->p = *p;
    int n = p->n;
    size_t L = (size_t)(p->local_count > 0 ? p->local_count : 1);
    s->mu = (double *)malloc(sizeof(double) * n);
    s->sigma = (double *)malloc(sizeof(double) * n);
    memcpy(s->mu, x0, sizeof(double) * n);                          /* :246 */
    for (int d = 0; d < n; ++d) s->sigma[d] = p->std_dev ? p->std_dev[d] : 1.0; /* :247-251 */
    s->p.std_dev = NULL;
    s->xs = (double *)malloc(sizeof(double) * L * n);
    s->costs = (double *)malloc(sizeof(double) * L);
    s->order = (int64_t *)malloc(sizeof(int64_t) * L);
    s->elite = (int64_t *)malloc(sizeof(int64_t) * (size_t)p->elite_count);
    s->z = (double *)malloc(sizeof(double) * (n + 1));
    s->iter = 1;                                                    /* :253 */
    return s;
}

void orc_ce_destroy(orc_ce *s)
{
    if (!s) return;
    free(s->mu); free(s->sigma); free(s->xs); free(s->costs); free(s->order); free(s->elite);
    free(s->z); free(s);
}

static double norm_inf(const double *v, int n)
{
    double m = 0.0;
    for (int d = 0; d < n; ++d) { double a = fabs(v[d]); if (a > m) m = a; }
    return m;
}

int orc_ce_active(const orc_ce *s)
{
    return norm_inf(s->sigma, s->p.n) > s->p.tol && s->iter < s->p.iter_max; /* :255 */
}

static int sigma_finite(const orc_ce *s)
{
    for (int d = 0; d < s->p.n; ++d) if (!isfinite(s->sigma[d])) return 0;
    return 1;
}

void orc_ce_sample_local(orc_ce *s)
{
    const orc_ce_params *p = &s->p;
    int n = p->n;
    for (int64_t i = 0; i < p->local_count; ++i) {                   /* :264-274 */
        int64_t gi = p->first_index + i;
        const double *z;
        if (p->replay_normals) {
            z = p->replay_normals + ((size_t)(s->iter - 1) * p->sample_size + gi) * n;
        } else {
            orc_ce_normals(p->seed, s->iter - 1, gi, n, s->z);
            z = s->z;
        }
        double *x = s->xs + (size_t)i * n;
        for (int d = 0; d < n; ++d) x[d] = s->mu[d] + s->sigma[d] * z[d]; /* Normal::sample = mean + std*z */
        s->costs[i] = orc_objective(p->objective_id, x, n);
    }
}

static const double *g_sort_costs;
static int cmp_cost_index(const void *a, const void *b)
{
    int64_t ia = *(const int64_t *)a, ib = *(const int64_t *)b;
    double ca = g_sort_costs[ia], cb = g_sort_costs[ib];
    if (ca < cb) return -1;
    if (ca > cb) return 1;
    return (ia > ib) - (ia < ib);
}

int orc_ce_step(orc_ce *s)
{
    const orc_ce_params *p = &s->p;
    int n = p->n;
    int64_t N = p->local_count, k = p->elite_count;
    if (!orc_ce_active(s)) return 1;            /* the loop condition comes first (:255): a non-finite sigma only ... */
    if (!sigma_finite(s)) return -ORC_ERR_NAN;  /* ... panics when another pass samples with it: Normal::new(..).unwrap(), :259 */
    orc_ce_sample_local(s);
    for (int64_t i = 0; i < N; ++i) {
        if (s->costs[i] != s->costs[i]) return -ORC_ERR_NAN; /* partial_cmp().unwrap() panics, :276 */
        s->order[i] = i;
    }
    g_sort_costs = s->costs;
    qsort(s->order, (size_t)N, sizeof(int64_t), cmp_cost_index);     /* :276 (ties: lowest index first) */
    for (int64_t j = 0; j < k; ++j) s->elite[j] = p->first_index + s->order[j];
    double div = p->mean_divisor == ORC_CE_DIV_REFERENCE_TEN ? 10.0 : (double)k; /* :287 */
    for (int d = 0; d < n; ++d) {                                    /* :281-297 */
        double mu = 0.0;
        for (int64_t j = 0; j < k; ++j) mu += s->xs[(size_t)s->order[j] * n + d];
        mu /= div;
        double sg = 0.0;
        for (int64_t j = 0; j < k; ++j) {
            double e = s->xs[(size_t)s->order[j] * n + d] - mu;
            sg += e * e;
        }
        sg /= (double)(k - 1);
        s->mu[d] = mu;       /* safe to overwrite: dimension d is not read again */
        s->sigma[d] = sqrt(sg);
    }
    s->iter++;                                                       /* :302 */
    return 0;
}

int orc_ce_solve(const orc_ce_params *p, const double *x0, double *out_mu, int64_t *iters_run)
{
    int st = orc_ce_validate(p);
    if (st != ORC_OK) return st;
    orc_ce *s = orc_ce_create(p, x0);
    int r = 0;
    while ((r = orc_ce_step(s)) == 0) {}
    memcpy(out_mu, s->mu, sizeof(double) * p->n);                    /* :305 */
    if (iters_run) *iters_run = s->iter - 1;
    orc_ce_destroy(s);
    return r < 0 ? -r : ORC_OK;
}

int64_t orc_ce_iters(const orc_ce *s) { return s->iter - 1; }

void orc_ce_read_state(const orc_ce *s, double *mu, double *sigma)
{
    if (mu) memcpy(mu, s->mu, sizeof(double) * s->p.n);
    if (sigma) memcpy(sigma, s->sigma, sizeof(double) * s->p.n);
}

void orc_ce_read_elites(const orc_ce *s, int64_t *idx)
{
    memcpy(idx, s->elite, sizeof(int64_t) * (size_t)s->p.elite_count);
}

void orc_ce_read_costs(const orc_ce *s, double *costs)
{
    memcpy(costs, s->costs, sizeof(double) * (size_t)s->p.local_count);
}

void orc_ce_elite_sum_local(orc_ce *s, double T, int64_t tie_take, double *sum_x)
{
    int n = s->p.n;
    for (int d = 0; d < n; ++d) sum_x[d] = 0.0;
    int64_t e = 0, ties = 0;
    for (int64_t i = 0; i < s->p.local_count; ++i) {
        double c = s->costs[i];
        int take = c < T || (c == T && ties++ < tie_take);
        if (!take) continue;
        s->order[e++] = i;
        for (int d = 0; d < n; ++d) sum_x[d] += s->xs[(size_t)i * n + d];
    }
    s->local_elites = e;
}

void orc_ce_elite_sqsum_local(orc_ce *s, const double *mu_new, double *sum_sq)
{
    int n = s->p.n;
    int64_t e = s->local_elites;
    for (int d = 0; d < n; ++d) sum_sq[d] = 0.0;
    for (int64_t j = 0; j < e; ++j) {
        const double *x = s->xs + (size_t)s->order[j] * n;
        for (int d = 0; d < n; ++d) { double t = x[d] - mu_new[d]; sum_sq[d] += t * t; }
    }
}

void orc_ce_apply(orc_ce *s, const double *mu_new, const double *sigma_new)
{
    memcpy(s->mu, mu_new, sizeof(double) * s->p.n);
    memcpy(s->sigma, sigma_new, sizeof(double) * s->p.n);
    s->iter++;
}

/* ============================== MonteCarlo ================================================== */

void orc_mc_draws(uint64_t seed, int64_t index, int n, double *u)
{
    for (int p = 0; 2 * p < n; ++p) {
        uint32_t w[4];
        philox_draw(seed, STREAM_MC, 0, index, (uint32_t)p, w);
        u[2 * p] = orc_u01(w[0], w[1]);
        if (2 * p + 1 < n) u[2 * p + 1] = orc_u01(w[2], w[3]);
    }
}

int orc_mc_integrate(int integrand_id, int n, int64_t count, const double *from, const double *to, uint64_t seed,
                     const double *replay_u, double *mean_out, double *variance_out)
{
    if (n <= 0) return ORC_ERR_INCORRECT_INPUT;
    double volume = 1.0;
    for (int d = 0; d < n; ++d) {
        /* Uniform::new_inclusive(x, y) -> ExternalError, monte_carlo.rs:171-172,191-195 */
        if (!(from[d] <= to[d]) || !isfinite(from[d]) || !isfinite(to[d]) || !isfinite(to[d] - from[d])) return ORC_ERR_EXTERNAL;
        volume *= to[d] - from[d];                      /* (to - &from).product(), :197 */
    }
    if (volume < 0.0) return ORC_ERR_INCORRECT_INPUT;   /* :148-152 */
    double *x = (double *)malloc(sizeof(double) * (size_t)n), *u = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    double mean = 0.0, sum_of_squares = 0.0;             /* lib.rs:176-178 */
    int64_t total = 0;
    for (int64_t i = 0; i < count; ++i) {
        if (replay_u) memcpy(u, replay_u + (size_t)i * n, sizeof(double) * (size_t)n);
        else orc_mc_draws(seed, i, n, u);
        for (int d = 0; d < n; ++d) x[d] = u[d] * (to[d] - from[d]) + from[d];
        const double sample = orc_objective(integrand_id, x, n);
        const double c = (double)(i + 1);
        const double delta_mean = sample - mean;         /* lib.rs:184-186 */
        mean = mean + delta_mean / c;
        sum_of_squares = sum_of_squares + delta_mean * (sample - mean);
        total += 1;
    }
    free(x); free(u);
    if (total < 2) return ORC_ERR_INCORRECT_INPUT;      /* lib.rs:190-193 */
    const double variance = sum_of_squares / (double)(total - 1); /* bessel, lib.rs:195-199 */
    *mean_out = mean * volume;                           /* scale_mean, lib.rs:143-148 */
    *variance_out = variance * volume * volume;
    return ORC_OK;
}


/* Welford moments (count, mean, sum of squared deviations) of the points [first, first + count) only: one rank's
 * share of a sharded MonteCarlo (the ranks' moments are then merged pairwise in rank order, csrc/mc.cu) */
int orc_mc_moments(int integrand_id, int n, int64_t first, int64_t count, const double *from, const double *to,
                   uint64_t seed, double *moments)
{
    double *x = (double *)malloc(sizeof(double) * (size_t)n), *u = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    double mean = 0.0, sum_of_squares = 0.0;
    for (int64_t i = 0; i < count; ++i) {
        orc_mc_draws(seed, first + i, n, u);
        for (int d = 0; d < n; ++d) x[d] = u[d] * (to[d] - from[d]) + from[d];
        const double sample = orc_objective(integrand_id, x, n);
        const double delta_mean = sample - mean;
        mean = mean + delta_mean / (double)(i + 1);
        sum_of_squares = sum_of_squares + delta_mean * (sample - mean);
    }
    free(x); free(u);
    moments[0] = (double)count;
    moments[1] = mean;
    moments[2] = sum_of_squares;
    return ORC_OK;
}

/* ============================== MISER ======================================================= */
/* miser_recurse, src/integrators/miser.rs:301-443.  The reference threads one rng through the recursion; here
 * every draw has a counter: stream 4, word 3 = node id (root 1, children 2 id, 2 id + 1), index =
 * (batch << 40) | point, batch = 2 d + half for the variance probes, 2 n for the node's own scalars (point 0: the
 * fallback dimension, point 1 + d: the dither of dimension d), 2 n + 1 for the leaf's Monte Carlo. */

typedef struct {
    int integrand_id, n;
    int64_t min_dim_count;
    double alpha, dither;
    uint64_t seed;
    int64_t evals;
} miser_cfg;

static int uniform_ok(double lo, double hi) { return lo <= hi && isfinite(lo) && isfinite(hi) && isfinite(hi - lo); }

/* from_iterator_using_welford over `count` points of the box [lo, hi], lib.rs:172-202 */
static int miser_batch(miser_cfg *m, const double *lo, const double *hi, int64_t count, uint32_t node, int64_t gbase,
                       int bessel, double *mean_out, double *var_out)
{
    int n = m->n;
    double *x = (double *)malloc(sizeof(double) * (size_t)n);
    double mean = 0.0, sum_of_squares = 0.0;
    int64_t total = 0;
    for (int64_t i = 0; i < count; ++i) {
        for (int p = 0; 2 * p < n; ++p) {
            uint32_t w[4];
            philox_draw(m->seed, STREAM_MISER, node, gbase | i, (uint32_t)p, w);
            x[2 * p] = orc_u01(w[0], w[1]) * (hi[2 * p] - lo[2 * p]) + lo[2 * p];
            if (2 * p + 1 < n) x[2 * p + 1] = orc_u01(w[2], w[3]) * (hi[2 * p + 1] - lo[2 * p + 1]) + lo[2 * p + 1];
        }
        const double sample = orc_objective(m->integrand_id, x, n);
        const double c = (double)(i + 1);
        const double delta_mean = sample - mean;
        mean = mean + delta_mean / c;
        sum_of_squares = sum_of_squares + delta_mean * (sample - mean);
        total += 1;
    }
    free(x);
    m->evals += count;
    if (total < 2) return ORC_ERR_INCORRECT_INPUT;
    if (bessel) total -= 1;
    *mean_out = mean;
    *var_out = sum_of_squares / (double)total;
    return ORC_OK;
}

static int to_usize(double v, int64_t *out) /* num_traits ToPrimitive::to_usize for f64 */
{
    if (!(v > -1.0 && v < 18446744073709551616.0)) return 0;
    *out = v < 0.0 ? 0 : (int64_t)(v < 9.0e18 ? v : 9.0e18);
    return 1;
}

static int miser_recurse(miser_cfg *m, double *from, double *to, uint32_t node, int depth_remaining, int64_t sample_count,
                         double *mean_out, double *var_out)
{
    int n = m->n;
    const int64_t minimum_sample_count = m->min_dim_count * n;         /* :319 */
    const int64_t ctl = (int64_t)(2 * n) << 40;
    for (int d = 0; d < n; ++d)
        if (!uniform_ok(from[d], to[d])) return ORC_ERR_EXTERNAL;       /* samplers, :320-324 */
    if (depth_remaining == 0 || sample_count < minimum_sample_count) { /* :335-349 */
        const int64_t count = sample_count > minimum_sample_count ? sample_count : minimum_sample_count;
        double volume = 1.0, mean, var;
        for (int d = 0; d < n; ++d) volume = volume * (to[d] - from[d]);
        int st = miser_batch(m, from, to, count, node, (int64_t)(2 * n + 1) << 40, 0, &mean, &var);
        if (st) return st;
        *mean_out = mean * volume;                                     /* scale_mean */
        *var_out = var * volume * volume;
        return ORC_OK;
    }
    if (sample_count < 2) sample_count = 2;                            /* :351 */
    double best_variance = DBL_MAX, best_lower_variance = DBL_MAX, best_upper_variance = DBL_MAX; /* :356-358 */
    uint32_t w[4];
    philox_draw(m->seed, STREAM_MISER, node, ctl, 0u, w);
    int best_dimension = (int)(orc_u01(w[0], w[1]) * (double)n);       /* uniform(0, n - 1), :359 */
    if (best_dimension > n - 1) best_dimension = n - 1;
    double best_mid = from[best_dimension] + (to[best_dimension] - from[best_dimension]) * 0.5; /* :360 */
    if (!uniform_ok(-m->dither, m->dither)) return ORC_ERR_EXTERNAL;   /* :361 */
    /* the GPU build draws every dither of a node before the first probe runs; without a shared rng stream the
     * order is immaterial, but an invalid bisection of ANY dimension fails the node before probing */
    double *mids = (double *)malloc(sizeof(double) * (size_t)n);
    for (int d = 0; d < n; ++d) {
        philox_draw(m->seed, STREAM_MISER, node, ctl | (int64_t)(1 + d), 0u, w);
        const double r = orc_u01(w[0], w[1]) * (m->dither - (-m->dither)) + (-m->dither);
        const double delta = to[d] - from[d];
        mids[d] = from[d] + delta * (0.5 + r);                         /* :364-365 */
        if (!uniform_ok(from[d], mids[d]) || !uniform_ok(mids[d], to[d])) { free(mids); return ORC_ERR_EXTERNAL; }
    }
    for (int d = 0; d < n; ++d) {
        const double mid = mids[d], old_from = from[d], old_to = to[d];
        double mean, variance_lower, variance_upper;
        to[d] = mid;                                                   /* lower_d_sampler, :366 */
        int st = miser_batch(m, from, to, sample_count, node, (int64_t)(2 * d) << 40, 1, &mean, &variance_lower);
        to[d] = old_to;
        if (!st) {
            from[d] = mid;                                             /* upper_d_sampler, :367 */
            st = miser_batch(m, from, to, sample_count, node, (int64_t)(2 * d + 1) << 40, 1, &mean, &variance_upper);
            from[d] = old_from;
        }
        if (st) { free(mids); return st; }
        const double total_variance = variance_lower + variance_upper;
        if (total_variance < best_variance) {                          /* :396-402 */
            best_variance = total_variance;
            best_lower_variance = variance_lower;
            best_upper_variance = variance_upper;
            best_dimension = d;
            best_mid = mid;
        }
    }
    free(mids);
    const double old_from = from[best_dimension], old_to = to[best_dimension];
    const double beta = 1.0 / (1.0 + m->alpha);                        /* :407 */
    const double la = pow(best_lower_variance, beta), ua = pow(best_upper_variance, beta);
    const double sum = la + ua;
    int64_t lower_count, upper_count;
    if (!to_usize((double)sample_count * (la / sum), &lower_count)) return ORC_ERR_TYPE_CONVERSION; /* :413-415 */
    if (!to_usize((double)sample_count * (ua / sum), &upper_count)) return ORC_ERR_TYPE_CONVERSION; /* :417-419 */
    double lm, lv, um, uv;
    to[best_dimension] = best_mid;                                     /* :422-431 */
    int st = miser_recurse(m, from, to, 2 * node, depth_remaining - 1, lower_count, &lm, &lv);
    to[best_dimension] = old_to;
    if (st) return st;
    from[best_dimension] = best_mid;                                   /* :433-440 */
    st = miser_recurse(m, from, to, 2 * node + 1, depth_remaining - 1, upper_count, &um, &uv);
    from[best_dimension] = old_from;
    if (st) return st;
    *mean_out = lm + um;                                               /* MeanVariance::add, :442 */
    *var_out = lv + uv;
    return ORC_OK;
}

int orc_miser_integrate(int integrand_id, int n, int64_t sample_count, int maximum_depth, int64_t min_dim_count,
                        double alpha, double dither, const double *from, const double *to, uint64_t seed,
                        double *mean_out, double *variance_out, int64_t *evals_out)
{
    if (n <= 0 || maximum_depth < 0 || maximum_depth > 30) return ORC_ERR_INCORRECT_INPUT;
    miser_cfg m = {integrand_id, n, min_dim_count, alpha, dither, seed, 0};
    double *f = (double *)malloc(sizeof(double) * (size_t)n), *t = (double *)malloc(sizeof(double) * (size_t)n);
    memcpy(f, from, sizeof(double) * (size_t)n);
    memcpy(t, to, sizeof(double) * (size_t)n);
    int st = miser_recurse(&m, f, t, 1u, maximum_depth, sample_count, mean_out, variance_out);
    if (evals_out) *evals_out = m.evals;
    free(f); free(t);
    return st;
}

/* ============================== all-cores variant =============================== */
/* Synchronous mode only: the particles of one pass are independent, so T threads each take a contiguous slice
 * (own draw scratch); the min-loc and the gbest update stay on the calling thread.  Bit-identical to orc_pso_step. */

typedef struct {
    orc_pso *s;
    int64_t lo, hi;
} mt_job;

static void *mt_worker(void *arg)
{
    mt_job *j = (mt_job *)arg;
    orc_pso *s = j->s;
    int n = s->p.n;
    double *scratch = (double *)malloc(sizeof(double) * 2 * (size_t)n);
    for (int64_t i = j->lo; i < j->hi; ++i) {
        double cost = move_particle_with(s, i, scratch, scratch + n);
        s->cost[i] = cost;
        if (cost < s->pbc[i]) {
            memcpy(s->pb + (size_t)i * n, s->x + (size_t)i * n, sizeof(double) * n);
            s->pbc[i] = cost;
        }
    }
    free(scratch);
    return NULL;
}

int orc_pso_step_mt(orc_pso *s, int threads)
{
    if (s->p.mode == ORC_PSO_SEQUENTIAL) return -1;
    if (stalled_too_long(s)) return 1;
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t tid[256];
    mt_job job[256];
    int64_t N = s->p.local_count;
    for (int t = 0; t < threads; ++t) {
        job[t].s = s;
        job[t].lo = N * t / threads;
        job[t].hi = N * (t + 1) / threads;
        pthread_create(&tid[t], NULL, mt_worker, &job[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(tid[t], NULL);
    double *record = (double *)malloc(sizeof(double) * (2 + s->p.n));
    local_minloc(s, record);
    orc_pso_step_apply(s, 1, record);
    free(record);
    return 0;
}

/* ============================== timing helpers ============================================== */

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

double orc_time_pso_steps(orc_pso *s, int64_t iters)
{
    double t0 = now_s();
    for (int64_t i = 0; i < iters; ++i) orc_pso_step(s);
    return now_s() - t0;
}

double orc_time_pso_steps_mt(orc_pso *s, int64_t iters, int threads)
{
    double t0 = now_s();
    for (int64_t i = 0; i < iters; ++i) orc_pso_step_mt(s, threads);
    return now_s() - t0;
}

double orc_time_ce_steps(orc_ce *s, int64_t iters)
{
    double t0 = now_s();
    for (int64_t i = 0; i < iters; ++i) orc_ce_step(s);
    return now_s() - t0;
}
