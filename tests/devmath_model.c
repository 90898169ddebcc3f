/* devmath_model.c -- CPU model of the formulas in eqsolver_b200/csrc/devmath.cuh and philox.cuh (normal_pair).
 *
 * Test infrastructure, not product code: the device functions cannot run on the host (MUFU seeds, __constant__ tables), so
 * this file restates their arithmetic operation for operation with libm's exact fma() and a 20-bit truncation standing in
 * for rcp.approx / rsqrt.approx, and compares with glibc -- the same libm the oracle (and Rust on linux-gnu) uses.  It backs
 * the accuracy figures quoted in DESIGN.md section 4; the device code itself is checked on the GPU
 * (tests/test_pso_gpu.py::test_device_cos_accuracy, ::test_box_muller_normals_stay_within_ulps_of_the_oracle,
 * tests/test_ce_gpu.py).  If a formula in devmath.cuh changes, change it here too.
 *
 *   gcc -O2 -ffp-contract=off -o devmath_model devmath_model.c -lm && ./devmath_model [samples]
 *
 * Prints one JSON object: max error of the logarithms in ulp and of the cosine in absolute terms, mismatches of the square root, and for the Box-Muller normals
 * the share that equals the oracle's bit for bit and the largest deviation in units of max(ulp(z), ulp(1)). */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static const double S[6] = {1.58969099521155010221e-10,  -2.50507602534068634195e-08, 2.75573137070700676789e-06,
                            -1.98412698298579493134e-04, 8.33333333332248946124e-03,  -1.66666666666666324348e-01};
static const double C[6] = {-1.13596475577881948265e-11, 2.08757232129817482790e-09,  -2.75573143513906633035e-07,
                            2.48015872894767294178e-05,  -1.38888888888741095749e-03, 4.16666666666666019037e-02};
static const double L[7] = {6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01, 2.222219843214978396e-01,
                            1.818357216161805012e-01, 1.531383769920937332e-01, 1.479819860511658591e-01};
static const double MAGIC = 6755399441055744.0, TWO_PI = 6.28318530717958647692;
#define __device__
#include "../eqsolver_b200/csrc/devmath_tables.inc" /* g_logtab, g_ktab: the tables log_unit_tab reads from shared memory */

static double approx20(double x) /* what MUFU.RCP64H / RSQ64H hand out: about 20 good bits, low word zero */
{
    uint64_t u;
    memcpy(&u, &x, 8);
    u &= 0xffffffff00000000ULL;
    memcpy(&x, &u, 8);
    return x;
}

static int lo_word(double t)
{
    uint64_t u;
    memcpy(&u, &t, 8);
    return (int)(uint32_t)u;
}

/* sincos_kernel */
static void sincos_kernel(double y, double *sn, double *cs)
{
    const double y2 = y * y;
    double ps = fma(y2, S[0], S[1]), pc = fma(y2, C[0], C[1]);
    for (int i = 2; i < 6; ++i) ps = fma(ps, y2, S[i]);
    for (int i = 2; i < 6; ++i) pc = fma(pc, y2, C[i]);
    *sn = fma(y * y2, ps, y);
    *cs = fma(y2, fma(pc, y2, -0.5), 1.0);
}

/* round 1's eqs_cos_t (reduction by pi/2, both kernels): kept for comparison */
static double cos_pio2_model(double y)
{
    const double t = fma(y, 0.63661977236758134308, MAGIC);
    const int k = lo_word(t);
    const double kf = t - MAGIC;
    double r = fma(kf, -1.57079632679489655800e+00, y), sn, cs;
    r = fma(kf, -6.12323399573676603587e-17, r);
    sincos_kernel(r, &sn, &cs);
    const double res = (k & 1) ? sn : cs;
    return ((k + 1) & 2) ? -res : res;
}

/* eqs_cos_t, |y| < 8e5: reduction by pi, ONE even polynomial on [-pi/2, pi/2] (13 FP64 operations instead of 19) */
static const double CP[8] = {4.627675699925956e-14, -1.1464689862855446e-11, 2.0876630866706838e-09, -2.7557317767309485e-07,
                             2.4801587292446125e-05, -0.0013888888888860709, 0.04166666666666634, -0.5};
static double cos_model(double y)
{
    const double t = fma(y, 0.31830988618379067154, MAGIC);
    const int n = lo_word(t);
    const double kf = t - MAGIC;
    double r = fma(kf, -3.14159265358979311600e+00, y);
    r = fma(kf, -1.22464679914735317723e-16, r);
    const double r2 = r * r;
    double p = fma(r2, CP[0], CP[1]);
    for (int i = 2; i < 8; ++i) p = fma(p, r2, CP[i]);
    const double res = fma(r2, p, 1.0);
    return (n & 1) ? -res : res;
}

/* log_unit, a in (0, 1] */
static double log_model(double a)
{
    uint64_t u;
    memcpy(&u, &a, 8);
    int hx = (int)(u >> 32), k = (hx >> 20) - 1023;
    const uint32_t lx = (uint32_t)u;
    hx &= 0x000fffff;
    const int i = (hx + 0x95f64) & 0x100000;
    const uint64_t mu = ((uint64_t)(uint32_t)(hx | (i ^ 0x3ff00000)) << 32) | lx;
    double m;
    memcpy(&m, &mu, 8);
    k += i >> 20;
    const double f = m - 1.0, y = 2.0 + f;
    double rc = approx20(1.0 / y), e = fma(-y, rc, 1.0);
    rc = fma(rc, e, rc);
    e = fma(-y, rc, 1.0);
    rc = fma(rc, e, rc);
    const double s = f * rc, z = s * s, w = z * z;
    const double t1 = w * fma(w, fma(w, L[5], L[3]), L[1]);
    const double R = fma(z, fma(w, fma(w, fma(w, L[6], L[4]), L[2]), L[0]), t1);
    const double hfsq = 0.5 * f * f, dk = (double)k;
    return fma(dk, 6.93147180369123816490e-01, f - (hfsq - fma(s, hfsq + R, dk * 1.90821492927058770002e-10)));
}

/* log_unit_tab, a in (0, 1]: table-driven, 10 FP64 operations */
static double log_tab_model(double a)
{
    uint64_t u;
    memcpy(&u, &a, 8);
    int hx = (int)(u >> 32), k = (hx >> 20) - 1023;
    const uint32_t lx = (uint32_t)u;
    hx &= 0x000fffff;
    const int i = (hx + 0x95f64) & 0x100000;
    const int hm = hx | (i ^ 0x3ff00000);
    const uint64_t mu = ((uint64_t)(uint32_t)hm << 32) | lx;
    double m;
    memcpy(&m, &mu, 8);
    k += i >> 20;
    const double *te = g_logtab[((hm + 0x800) >> 12) - 0x3fe6a], *kt = g_ktab[-k];
    const double r = fma(m, te[0], -1.0), r2 = r * r;
    double q = fma(r, -1.0 / 6.0, 0.2);
    q = fma(q, r, -0.25);
    q = fma(q, r, 1.0 / 3.0);
    q = fma(q, r, -0.5);
    const double c = fma(r2, q, r);
    return kt[0] + (te[1] + (c + kt[1]));
}

/* sqrt_boxmuller */
static double sqrt_model(double a)
{
    uint64_t au;
    memcpy(&au, &a, 8);
    if ((int)((au >> 32) & 0x7fffffff) < 0x03500000) return a;
    const double y0 = approx20(1.0 / sqrt(a));
    const double e = fma(a, -(y0 * y0), 1.0);
    const double y1 = fma(fma(e, 0.375, 0.5), y0 * e, y0);
    const double g = a * y1, h = 0.5 * y1;
    return fma(fma(g, -g, a), h, g);
}

/* normal_pair on the two mantissa doubles m1, m2 in [1, 2) */
static void normal_pair_model(double m1, double m2, double *z0, double *z1)
{
    const double r = sqrt_model(-2.0 * log_tab_model(2.0 - m1));
    const double tt = fma(m2, 4.0, MAGIC - 4.0);
    const int k = lo_word(tt);
    const double q = fma(tt - (MAGIC - 4.0), -0.25, m2);
    double s, c;
    sincos_kernel(q * TWO_PI, &s, &c);
    const double a = (k & 1) ? c : s, b = (k & 1) ? s : c;
    *z0 = r * (((k + 1) & 2) ? -b : b);
    *z1 = r * ((k & 2) ? -a : a);
}

/* the oracle's normals: oracle/eqs_oracle.c orc_ce_normals + sincospi_host (glibc log, sqrt, sin, cos) */
static void normal_pair_oracle(double m1, double m2, double *z0, double *z1)
{
    const double u1 = m1 - 1.0, u2 = m2 - 1.0, r = sqrt(-2.0 * log(1.0 - u1)), t = 2.0 * u2;
    const double k = nearbyint(2.0 * t), rr = t - 0.5 * k, sr = sin(M_PI * rr), cr = cos(M_PI * rr);
    double s, c;
    switch (((int)k) & 3) {
    case 0: s = sr; c = cr; break;
    case 1: s = cr; c = -sr; break;
    case 2: s = -sr; c = -cr; break;
    default: s = -cr; c = sr; break;
    }
    *z0 = r * c;
    *z1 = r * s;
}

static double ulp_of(double x) { return nextafter(fabs(x), INFINITY) - fabs(x); }
static double mantissa_double(void) { return 1.0 + floor(drand48() * 4503599627370496.0) / 4503599627370496.0; }

int main(int argc, char **argv)
{
    const long n = argc > 1 ? atol(argv[1]) : 2000000;
    srand48(12345);
    double cospi_abs = 0.0;
    double log_ulp = 0.0, logtab_ulp = 0.0, cos_ulp = 0.0, cos_abs = 0.0, normal_units = 0.0;
    long sqrt_mismatch = 0, normal_equal = 0;
    for (long i = 0; i < n; ++i) {
        /* log on (0, 1]: half uniform, half spread over the 52 binades the uniforms can reach */
        double a = (i & 1) ? drand48() : ldexp(0.5 + 0.5 * drand48(), -(int)(drand48() * 52));
        if (!(a > 0.0 && a <= 1.0)) a = 0.5;
        const double wl = log(a), el = fabs(log_model(a) - wl) / ulp_of(wl), et = fabs(log_tab_model(a) - wl) / ulp_of(wl);
        if (wl != 0.0 && el > log_ulp) log_ulp = el;
        if (wl != 0.0 && et > logtab_ulp) logtab_ulp = et;
        /* close to 1, where ln(a) is tiny and only the c = 1 interval's relative accuracy saves it */
        const double a1 = 1.0 - ldexp(drand48(), -(int)(drand48() * 52)), wl1 = log(a1);
        if (a1 > 0.0 && wl1 != 0.0 && fabs(log_tab_model(a1) - wl1) / ulp_of(wl1) > logtab_ulp) logtab_ulp = fabs(log_tab_model(a1) - wl1) / ulp_of(wl1);
        const double q = -2.0 * wl;
        if (sqrt_model(q) != sqrt(q)) ++sqrt_mismatch;
        /* cos on the arguments the objectives produce */
        const double y = (i % 3 == 0) ? (2 * drand48() - 1) * 700.0 : (i % 3 == 1) ? (2 * drand48() - 1) * 8e5 : TWO_PI * (2 * drand48() - 1) * 5.12;
        const double wc = cos(y), dc = fabs(cos_model(y) - wc);
        if (dc > cos_abs) cos_abs = dc;
        const double dp = fabs(cos_pio2_model(y) - wc);
        if (dp > cospi_abs) cospi_abs = dp;
        if (dc > 2.8e-16 && dc / ulp_of(wc) > cos_ulp) cos_ulp = dc / ulp_of(wc); /* the bound is absolute: 1.25 ulp of 1.0 */
        /* normals */
        const double m1 = mantissa_double(), m2 = mantissa_double();
        double g0, g1, w0, w1;
        normal_pair_model(m1, m2, &g0, &g1);
        normal_pair_oracle(m1, m2, &w0, &w1);
        normal_equal += (g0 == w0) + (g1 == w1);
        const double e0 = fabs(g0 - w0) / fmax(ulp_of(w0), 2.220446049250313e-16), e1 = fabs(g1 - w1) / fmax(ulp_of(w1), 2.220446049250313e-16);
        if (e0 > normal_units) normal_units = e0;
        if (e1 > normal_units) normal_units = e1;
    }
    printf("{\"samples\": %ld, \"log_max_ulp\": %.3f, \"log_tab_max_ulp\": %.3f, \"cos_max_abs\": %.3e, \"cos_round1_max_abs\": %.3e, \"cos_max_ulp_beyond_2.8e-16\": %.3f, \"sqrt_mismatches\": %ld, "
           "\"normals_equal_share\": %.4f, \"normals_max_units\": %.2f, \"sqrt_zero\": %g, \"sqrt_minus_zero_is_negative\": %d, \"log_one\": %g}\n",
           n, log_ulp, logtab_ulp, cos_abs, cospi_abs, cos_ulp, sqrt_mismatch, (double)normal_equal / (2.0 * n), normal_units, sqrt_model(0.0),
           signbit(sqrt_model(-0.0)) ? 1 : 0, log_tab_model(1.0));
    return 0;
}
