/*
 * mcp_oracle.c — CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of MCPower's per-simulation hot loop, following the
 * reference's operation order.  File:line citations are relative to
 * /root/reference/crates/engine-core/src/.  Build: oracle/Makefile
 * (gcc -O2 -ffp-contract=off — the f32 data plane must not be FMA-contracted,
 * rng.rs:199-202,327-336; explicit fma() is used only where noted).
 *
 * Third-party arithmetic that is NOT in the reference tree and is restated
 * here from its published algorithm:
 *   rand_philox 0.1.0 — Philox4x32-10 (Random123), Mix13 splitmix64 finaliser,
 *                       23-bit open-interval u32→f32.   KAT-pinned.
 *   glmm 0.1.2        — OLS sufficient statistics + normal-equation solve,
 *                       IRLS, exp/ln kernels ("<=1-2 ULP of libm").  Value-pinned
 *                       (1e-11 OLS / 1e-4 GLM R goldens); summation order and
 *                       last-ulp behaviour are UNPINNED, so this oracle fixes its
 *                       own order (sequential rows, fma accumulate) and its own
 *                       exp/log (orc_exp/orc_log below, <1 ULP) and the CUDA path
 *                       mirrors that order op-for-op.
 *   faer 0.24.0       — LLT of the k×k correlation (bit-pinned only for k<=2 by
 *                       the validation fixtures), self-adjoint EVD (unpinned).
 */
#include "mcp_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static __thread char g_err[256];
const char* orc_last_error(void) { return g_err; }
static int fail(int code, const char* msg) {
    snprintf(g_err, sizeof g_err, "%s", msg);
    return code;
}

#define FLOAT_NEAR_ZERO 1e-30 /* lib.rs:46 */
#define EPS_RANK 1e-12        /* batch.rs:41 */

/* ======================================================================= */
/* RNG primitives                                                          */
/* ======================================================================= */

/* rand_philox::splitmix64 — Mix13 finaliser only, no Weyl add
 * (tests/rng_byte_identity.rs:36-42). */
uint64_t orc_splitmix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

/* rng.rs:36-38 */
uint64_t orc_pcg_mix64(uint64_t base_seed, uint64_t sim_id) {
    uint64_t r = (sim_id << 32) | (sim_id >> 32);
    return orc_splitmix64(base_seed ^ r);
}

/* Random123 Philox4x32-10 (KATs tests/rng_byte_identity.rs:17-33). */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
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

/* rand_philox::u32_to_unit_f32: (2*(w>>9)+1) / 2^24 — open interval, exact. */
float orc_u32_to_unit_f32(uint32_t w) {
    return (float)(2u * (w >> 9) + 1u) * (1.0f / 16777216.0f);
}

/* rng.rs:290-325 — frozen coefficients. */
static const float NI_BP = 4.757499993e-01f;
static const float NI_FLOOR = 2.980232239e-08f;
static const float NI_LN2 = 6.931471825e-01f;
static const float NI_LOG2[6] = {-2.786813021e+00f, 5.046875954e+00f, -3.492494345e+00f,
                                 1.593901396e+00f,  -4.048671722e-01f, 4.342890903e-02f};
static const float NI_CENTRAL[11] = {2.506664753e+00f,  2.586458445e+00f,  1.238584900e+01f, -4.216123047e+02f,
                                     1.463396387e+04f,  -2.779154375e+05f, 3.215264750e+06f, -2.294391600e+07f,
                                     9.886485600e+07f,  -2.359132160e+08f, 2.401336800e+08f};
static const float NI_TAIL[10] = {-2.150734663e+00f, 2.428994656e+00f,  -7.639000416e-01f, 2.937270701e-01f,
                                  -8.018484712e-02f, 1.541402005e-02f,  -2.041375730e-03f, 1.774382981e-04f,
                                  -9.117987247e-06f, 2.100489240e-07f};

/* rng.rs:329-336 — ascending coefficients, separate mul and add. */
static inline float horner_f32(const float* c, int n, float x) {
    float acc = c[n - 1];
    for (int i = n - 2; i >= 0; i--) acc = acc * x + c[i];
    return acc;
}

/* rng.rs:342-348 */
static inline float ln_f32(float t) {
    uint32_t bits;
    memcpy(&bits, &t, 4);
    float e = (float)((int32_t)((bits >> 23) & 0xff) - 127);
    uint32_t mb = (bits & 0x007fffffu) | 0x3f800000u;
    float m;
    memcpy(&m, &mb, 4);
    return (e + horner_f32(NI_LOG2, 6, m)) * NI_LN2;
}

/* rng.rs:354-364 */
float orc_norm_inv_cdf_f32(float u) {
    float v = u - 0.5f;
    float a = fabsf(v);
    if (a <= NI_BP) {
        return v * horner_f32(NI_CENTRAL, 11, v * v);
    } else {
        float t = 0.5f - a;
        if (!(t > NI_FLOOR)) t = NI_FLOOR; /* f32::max */
        float q = sqrtf(-2.0f * ln_f32(t));
        return copysignf(horner_f32(NI_TAIL, 10, q), v);
    }
}

/* rng.rs:51-128 — sequential per-sim stream, counter [c_lo, c_hi, 0, 0]. */
typedef struct {
    uint32_t key[2];
    uint64_t counter;
    uint32_t buf[4];
    int pos;
} SeqRng;

static void seq_init(SeqRng* r, uint64_t base_seed, uint64_t sim_id) {
    uint64_t k = orc_pcg_mix64(base_seed, sim_id);
    r->key[0] = (uint32_t)k;
    r->key[1] = (uint32_t)(k >> 32);
    r->counter = 0;
    r->pos = 4;
}
static inline uint32_t seq_u32(SeqRng* r) {
    if (r->pos == 4) {
        uint32_t c[4] = {(uint32_t)r->counter, (uint32_t)(r->counter >> 32), 0, 0};
        orc_philox4x32_10(c, r->key, r->buf);
        r->counter++;
        r->pos = 0;
    }
    return r->buf[r->pos++];
}
static inline float seq_uniform(SeqRng* r) { return orc_u32_to_unit_f32(seq_u32(r)); }
static inline float seq_normal(SeqRng* r) { return orc_norm_inv_cdf_f32(seq_uniform(r)); }
/* rng.rs:112-127 */
static int seq_categorical(SeqRng* r, const double* probs, int n) {
    double u = (double)seq_uniform(r);
    double acc = 0.0;
    for (int i = 0; i < n; i++) {
        acc += probs[i];
        if (u < acc) return i;
    }
    return n - 1;
}

/* rng.rs:154-178 — planar word(class, col, row) */
static void fill_words(const uint32_t key[2], uint32_t cls, uint32_t col, int n, uint32_t* out) {
    uint32_t w[4];
    for (int i = 0; i < n; i += 4) {
        uint32_t c[4] = {(uint32_t)(i >> 2), col, cls, 0};
        orc_philox4x32_10(c, key, w);
        int take = n - i < 4 ? n - i : 4;
        for (int t = 0; t < take; t++) out[i + t] = w[t];
    }
}
/* rng.rs:236-261 (bit-identical to the scalar kernel by construction) */
static void fill_normal_column(const uint32_t key[2], uint32_t cls, uint32_t col, int n, uint32_t* words, float* out) {
    fill_words(key, cls, col, n, words);
    for (int i = 0; i < n; i++) out[i] = orc_norm_inv_cdf_f32(orc_u32_to_unit_f32(words[i]));
}
static void fill_uniform_column(const uint32_t key[2], uint32_t cls, uint32_t col, int n, uint32_t* words, float* out) {
    fill_words(key, cls, col, n, words);
    for (int i = 0; i < n; i++) out[i] = orc_u32_to_unit_f32(words[i]);
}

static inline uint32_t f32_bits(float f) {
    uint32_t b;
    memcpy(&b, &f, 4);
    return b;
}

void orc_seq_stream(uint64_t base_seed, uint64_t sim_id, int kind, int n, uint32_t* out_bits) {
    SeqRng r;
    seq_init(&r, base_seed, sim_id);
    for (int i = 0; i < n; i++) out_bits[i] = f32_bits(kind == 0 ? seq_normal(&r) : seq_uniform(&r));
}
void orc_seq_categorical(uint64_t base_seed, uint64_t sim_id, const double* probs, int n_probs, int n, int32_t* out) {
    SeqRng r;
    seq_init(&r, base_seed, sim_id);
    for (int i = 0; i < n; i++) out[i] = seq_categorical(&r, probs, n_probs);
}
void orc_planar_normals(uint64_t base_seed, uint64_t sim_id, uint32_t cls, uint32_t col, int n, uint32_t* out_bits) {
    uint64_t k = orc_pcg_mix64(base_seed, sim_id);
    uint32_t key[2] = {(uint32_t)k, (uint32_t)(k >> 32)};
    uint32_t* words = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(n + 4));
    float* z = (float*)malloc(sizeof(float) * (size_t)(n + 4));
    fill_normal_column(key, cls, col, n, words, z);
    for (int i = 0; i < n; i++) out_bits[i] = f32_bits(z[i]);
    free(words);
    free(z);
}

/* ======================================================================= */
/* Owned exp / log (stand-ins for glmm::simd_transcendental, "<=1-2 ULP")   */
/* Pure IEEE f64 +,*,fma so the CUDA twin is bit-identical.                 */
/* ======================================================================= */

static inline double bits_f64(uint64_t b) {
    double d;
    memcpy(&d, &b, 8);
    return d;
}
static inline uint64_t f64_bits(double d) {
    uint64_t b;
    memcpy(&b, &d, 8);
    return b;
}

double orc_exp(double x) {
    if (x != x) return x;
    if (x > 709.782712893384) return INFINITY;
    if (x < -745.1332191019412) return 0.0;
    const double INV_LN2 = 1.4426950408889634;
    const double LN2_HI = 6.93147180369123816490e-01;
    const double LN2_LO = 1.90821492927058770002e-10;
    double kf = rint(x * INV_LN2);
    double r = fma(-kf, LN2_HI, x);
    r = fma(-kf, LN2_LO, r);
    /* Taylor to r^13, |r| <= 0.3466: truncation < 2^-57 */
    double p = 1.0 / 6227020800.0;
    p = fma(p, r, 1.0 / 479001600.0);
    p = fma(p, r, 1.0 / 39916800.0);
    p = fma(p, r, 1.0 / 3628800.0);
    p = fma(p, r, 1.0 / 362880.0);
    p = fma(p, r, 1.0 / 40320.0);
    p = fma(p, r, 1.0 / 5040.0);
    p = fma(p, r, 1.0 / 720.0);
    p = fma(p, r, 1.0 / 120.0);
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    int k = (int)kf;
    int k1 = k / 2, k2 = k - k1;
    double s1 = bits_f64((uint64_t)(k1 + 1023) << 52);
    double s2 = bits_f64((uint64_t)(k2 + 1023) << 52);
    return p * s1 * s2;
}

/* fdlibm-style log: x = 2^k (1+f), s = f/(2+f), log(1+f) = f - hfsq + s(hfsq+R(s^2)). */
double orc_log(double x) {
    if (x != x) return x;
    if (x < 0.0) return NAN;
    if (x == 0.0) return -INFINITY;
    if (x == INFINITY) return x;
    const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
    const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
                 Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                 Lg7 = 1.479819860511658591e-01;
    int k = 0;
    uint64_t b = f64_bits(x);
    if ((b >> 52) == 0) { /* subnormal */
        x *= 18014398509481984.0; /* 2^54 */
        b = f64_bits(x);
        k -= 54;
    }
    k += (int)(b >> 52) - 1023;
    uint64_t mant = b & 0x000fffffffffffffULL;
    /* normalise m into [sqrt(1/2), sqrt(2)) */
    if (mant >= 0x6a09e667f3bcdULL) {
        b = mant | 0x3fe0000000000000ULL;
        k += 1;
    } else {
        b = mant | 0x3ff0000000000000ULL;
    }
    double f = bits_f64(b) - 1.0;
    double s = f / (2.0 + f);
    double z = s * s;
    double w = z * z;
    double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
    double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
    double R = t2 + t1;
    double hfsq = 0.5 * f * f;
    double dk = (double)k;
    return dk * LN2_HI - ((hfsq - (s * (hfsq + R) + dk * LN2_LO)) - f);
}

/* ======================================================================= */
/* Transforms                                                              */
/* ======================================================================= */

/* distributions.rs:38-56 — Abramowitz–Stegun 7.1.26 */
static double as_erfc(double x) {
    const double A1 = 0.254829592, A2 = -0.284496736, A3 = 1.421413741, A4 = -1.453152027, A5 = 1.061405429,
                 P = 0.3275911;
    double sign = x < 0.0 ? -1.0 : 1.0;
    double ax = fabs(x);
    double t = 1.0 / (1.0 + P * ax);
    double y = 1.0 - (((((A5 * t + A4) * t) + A3) * t + A2) * t + A1) * t * orc_exp(-ax * ax);
    return 1.0 - sign * y;
}
/* distributions.rs:32-34 */
double orc_phi(double z) { return 0.5 * as_erfc(-z * 0.70710678118654752440); }

/* data_gen.rs:39-42 */
#define EXP_CAP 6.95925599364711
#define EXP_CENSORED_MEAN 0.9990501970288289
#define EXP_CENSORED_STD 0.9933676327697134
#define SQRT3 1.7320508075688772

/* ---- critvals needed by the t3 table come first ---- */

/* critvals.rs:33-87 — Acklam, no refinement */
double orc_norm_ppf(double p) {
    static const double A[6] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
                                1.38357751867269e2,     -3.066479806614716e+01, 2.506628277459239e+00};
    static const double B[5] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
                                6.680131188771972e+01, -1.328068155288572e+01};
    static const double C[6] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
                                -2.549732539343734e+00, 4.374664141464968e+00,  2.938163982698783e+00};
    static const double D[4] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00,
                                3.754408661907416e+00};
    const double P_LOW = 0.02425, P_HIGH = 1.0 - 0.02425;
    if (p <= 0.0) return -INFINITY;
    if (p >= 1.0) return INFINITY;
    if (p < P_LOW) {
        double q = sqrt(-2.0 * log(p));
        return (((((C[0] * q + C[1]) * q + C[2]) * q + C[3]) * q + C[4]) * q + C[5]) /
               ((((D[0] * q + D[1]) * q + D[2]) * q + D[3]) * q + 1.0);
    } else if (p <= P_HIGH) {
        double q = p - 0.5, r = q * q;
        return (((((A[0] * r + A[1]) * r + A[2]) * r + A[3]) * r + A[4]) * r + A[5]) * q /
               (((((B[0] * r + B[1]) * r + B[2]) * r + B[3]) * r + B[4]) * r + 1.0);
    } else {
        double q = sqrt(-2.0 * log(1.0 - p));
        return -(((((C[0] * q + C[1]) * q + C[2]) * q + C[3]) * q + C[4]) * q + C[5]) /
               ((((D[0] * q + D[1]) * q + D[2]) * q + D[3]) * q + 1.0);
    }
}

/* critvals.rs:112-135 — Lanczos g=7, n=9 */
static double ln_gamma(double x) {
    static const double P[9] = {0.9999999999998099,  676.5203681218851,    -1259.1392167224028,
                                771.3234287776531,   -176.6150291621406,   12.507343278686905,
                                -0.13857109526572012, 9.984369578019572e-6, 1.5056327351493116e-7};
    if (x < 0.5) return log(M_PI / sin(M_PI * x)) - ln_gamma(1.0 - x);
    x -= 1.0;
    double a = P[0];
    double t = x + 7.0 + 0.5;
    for (int i = 1; i < 9; i++) a += P[i] / (x + (double)i);
    return 0.5 * log(2.0 * M_PI) + (x + 0.5) * log(t) - t + log(a);
}
static double ln_beta(double a, double b) { return ln_gamma(a) + ln_gamma(b) - ln_gamma(a + b); }

/* critvals.rs:138-186 — modified Lentz */
static double betacf(double a, double b, double x) {
    const int MAXIT = 200;
    const double EPS = 3e-15, FPMIN = 1.0e-300;
    double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0, d = 1.0 - qab * x / qap;
    if (fabs(d) < FPMIN) d = FPMIN;
    d = 1.0 / d;
    double h = d;
    for (int m = 1; m <= MAXIT; m++) {
        double mf = (double)m, m2 = 2.0 * mf;
        double aa = mf * (b - mf) * x / ((qam + m2) * (a + m2));
        d = 1.0 + aa * d;
        if (fabs(d) < FPMIN) d = FPMIN;
        c = 1.0 + aa / c;
        if (fabs(c) < FPMIN) c = FPMIN;
        d = 1.0 / d;
        h *= d * c;
        aa = -(a + mf) * (qab + mf) * x / ((a + m2) * (qap + m2));
        d = 1.0 + aa * d;
        if (fabs(d) < FPMIN) d = FPMIN;
        c = 1.0 + aa / c;
        if (fabs(c) < FPMIN) c = FPMIN;
        d = 1.0 / d;
        double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < EPS) return h;
    }
    return h;
}
/* critvals.rs:91-105 */
static double betai(double a, double b, double x) {
    if (x <= 0.0) return 0.0;
    if (x >= 1.0) return 1.0;
    double bt = exp(a * log(x) + b * log(1.0 - x) - ln_beta(a, b));
    if (x < (a + 1.0) / (a + b + 2.0)) return bt * betacf(a, b, x) / a;
    return 1.0 - bt * betacf(b, a, 1.0 - x) / b;
}
/* critvals.rs:189-200 */
static double t_cdf(double x, double df) {
    if (!isfinite(x)) return x > 0.0 ? 1.0 : 0.0;
    double z = df / (df + x * x);
    double half = 0.5 * betai(0.5 * df, 0.5, z);
    return x >= 0.0 ? 1.0 - half : half;
}
/* critvals.rs:211-284 */
double orc_t_ppf(double p, double df) {
    if (!isfinite(df) || df <= 0.0) return NAN;
    if (p <= 0.0) return -INFINITY;
    if (p >= 1.0) return INFINITY;
    if (p == 0.5) return 0.0;
    int upper = p > 0.5;
    double q = upper ? p : 1.0 - p;
    double x;
    if (df > 200.0) {
        double z = orc_norm_ppf(q), z2 = z * z;
        x = z + (z * (z2 + 1.0)) / (4.0 * df) + (z * (5.0 * z2 * z2 + 16.0 * z2 + 3.0)) / (96.0 * df * df);
    } else if (fabs(df - 1.0) < 1e-9) {
        x = fabs(tan(M_PI * (q - 0.5)));
    } else if (fabs(df - 2.0) < 1e-9) {
        double alpha = 4.0 * q * (1.0 - q);
        x = sqrt(2.0 / alpha - 2.0) * (q > 0.5 ? 1.0 : -1.0);
    } else {
        x = orc_norm_ppf(q);
    }
    if (df > 1.0e5) return upper ? x : -x;
    double log_norm_const = ln_gamma(0.5 * (df + 1.0)) - ln_gamma(0.5 * df) - 0.5 * log(df * M_PI);
    for (int it = 0; it < 80; it++) {
        double cdf = t_cdf(x, df);
        double pdf = exp(log_norm_const - 0.5 * (df + 1.0) * log(1.0 + x * x / df));
        if (pdf <= 0.0 || !isfinite(pdf)) break;
        double step = (cdf - q) / pdf;
        double nx = x - step;
        if (!isfinite(nx)) break;
        double rel = fabs(nx - x) / (1.0 + fabs(x));
        x = nx;
        if (rel < 1e-14) break;
    }
    return upper ? x : -x;
}

/* critvals.rs:294-345 */
static double gamma_inc_p(double a, double x) {
    if (x <= 0.0 || a <= 0.0) return 0.0;
    const int MAXIT = 200;
    const double EPS = 3e-15, FPMIN = 1.0e-300;
    if (x < a + 1.0) {
        double ap = a, sum = 1.0 / a, term = sum;
        for (int i = 0; i < MAXIT; i++) {
            ap += 1.0;
            term *= x / ap;
            sum += term;
            if (fabs(term) < fabs(sum) * EPS) return sum * exp(-x + a * log(x) - ln_gamma(a));
        }
        return sum * exp(-x + a * log(x) - ln_gamma(a));
    } else {
        double b = x + 1.0 - a, c = 1.0 / FPMIN, d = 1.0 / b, h = d;
        for (int i = 1; i < MAXIT; i++) {
            double an = -(double)i * ((double)i - a);
            b += 2.0;
            d = an * d + b;
            if (fabs(d) < FPMIN) d = FPMIN;
            c = b + an / c;
            if (fabs(c) < FPMIN) c = FPMIN;
            d = 1.0 / d;
            double del = d * c;
            h *= del;
            if (fabs(del - 1.0) < EPS) return 1.0 - exp(-x + a * log(x) - ln_gamma(a)) * h;
        }
        return 1.0 - exp(-x + a * log(x) - ln_gamma(a)) * h;
    }
}
static double powi3(double v) { return v * v * v; } /* f64::powi(3) = x*x*x */

/* critvals.rs:356-406 */
double orc_chi2_ppf(double p, double k) {
    if (!isfinite(k) || k <= 0.0) return NAN;
    if (p <= 0.0) return 0.0;
    if (p >= 1.0) return INFINITY;
    double z = orc_norm_ppf(p);
    double h = 2.0 / (9.0 * k);
    double x = k * powi3(1.0 - h + z * sqrt(h));
    if (!isfinite(x) || x <= 0.0) x = k > 1e-6 ? k : 1e-6;
    double a = 0.5 * k;
    for (int it = 0; it < 80; it++) {
        double cdf = gamma_inc_p(a, 0.5 * x);
        double pdf = exp((a - 1.0) * log(0.5 * x) - 0.5 * x - M_LN2 - ln_gamma(a));
        if (!isfinite(pdf) || pdf <= 0.0) break;
        double step = (cdf - p) / pdf;
        double nx = x - step;
        if (!isfinite(nx) || nx <= 0.0) nx = 0.5 * x;
        double rel = fabs(nx - x) / (1.0 + fabs(x));
        x = nx;
        if (rel < 1e-13) break;
    }
    return x;
}
/* critvals.rs:416-441 */
static double f_pdf(double x, double dfn, double dfd) {
    if (x <= 0.0) return 0.0;
    double a = 0.5 * dfn, b = 0.5 * dfd;
    return exp(a * log(dfn / dfd) + (a - 1.0) * log(x) - (a + b) * log(1.0 + dfn / dfd * x) - ln_beta(a, b));
}
static double f_cdf(double x, double dfn, double dfd) {
    if (x <= 0.0) return 0.0;
    if (!isfinite(x)) return 1.0;
    double z = dfd / (dfd + dfn * x);
    return 1.0 - betai(0.5 * dfd, 0.5 * dfn, z);
}
/* critvals.rs:450-485 */
double orc_f_ppf(double p, double dfn, double dfd) {
    if (!isfinite(dfn) || !isfinite(dfd) || dfn <= 0.0 || dfd <= 0.0) return NAN;
    if (p <= 0.0) return 0.0;
    if (p >= 1.0) return INFINITY;
    double z = orc_norm_ppf(p);
    double h = 2.0 / (9.0 * dfn);
    double chi_seed = dfn * powi3(1.0 - h + z * sqrt(h));
    double x = chi_seed / dfn;
    if (!(x > 1e-6)) x = 1e-6;
    for (int it = 0; it < 80; it++) {
        double cdf = f_cdf(x, dfn, dfd);
        double pdf = f_pdf(x, dfn, dfd);
        if (!isfinite(pdf) || pdf <= 0.0) break;
        double step = (cdf - p) / pdf;
        double nx = x - step;
        if (!isfinite(nx) || nx <= 0.0) nx = 0.5 * x;
        double rel = fabs(nx - x) / (1.0 + fabs(x));
        x = nx;
        if (rel < 1e-13) break;
    }
    return x;
}

/* critvals.rs:493-746 — studentized range via composite 32-node Gauss–Legendre */
static const double GL32_X[16] = {0.04830766568773832, 0.1444719615827965, 0.23928736225213707, 0.33186860228212765,
                                  0.42135127613063535, 0.5068999089322294, 0.5877157572407623,  0.6630442669302152,
                                  0.7321821187402897,  0.7944837959679424, 0.84936761373257,    0.8963211557660521,
                                  0.9349060759377397,  0.9647622555875064, 0.9856115115452683,  0.9972638618494816};
static const double GL32_W[16] = {0.0965400885147278,  0.09563872007927486, 0.09384439908080457, 0.09117387869576389,
                                  0.08765209300440381, 0.08331192422694676, 0.07819389578707031, 0.0723457941088485,
                                  0.06582222277636185, 0.05868409347853555, 0.05099805926237618, 0.04283589802222669,
                                  0.03427386291302143, 0.02539206530926206, 0.01627439473090567, 0.007018610009470117};
static double norm_pdf(double z) { return exp(-0.5 * z * z) / sqrt(2.0 * M_PI); }
/* critvals.rs — phi here is the engine's A&S phi (critvals.rs:1 `use crate::distributions::phi`). */
static double range_integrand(double z, double w, double k) {
    double inner = orc_phi(z) - orc_phi(z - w);
    return inner > 0.0 ? norm_pdf(z) * pow(inner, k - 1.0) : 0.0;
}
static double range_cdf(double w, double k) {
    if (w <= 0.0) return 0.0;
    double lo = -8.0, hi = 8.0;
    int kk = (int)k;
    int panels = 4 + (kk > 0 ? kk - 1 : 0);
    double step = (hi - lo) / (double)panels, half = 0.5 * step, acc = 0.0;
    for (int p = 0; p < panels; p++) {
        double mid = lo + ((double)p + 0.5) * step;
        for (int i = 0; i < 16; i++) {
            double dx = half * GL32_X[i];
            acc += GL32_W[i] * (range_integrand(mid + dx, w, k) + range_integrand(mid - dx, w, k));
        }
    }
    double r = k * (half * acc);
    return r < 1.0 ? r : 1.0;
}
static double srange_g(double u, double q, double k, double half_df, double ln_c) {
    if (u <= 0.0) return 0.0;
    double dens = exp(ln_c + (half_df - 1.0) * log(u) - half_df * u);
    return dens * range_cdf(q * sqrt(u), k);
}
static double studentized_range_cdf(double q, double k, double df) {
    if (q <= 0.0) return 0.0;
    if (!isfinite(df) || df > 2.0e4) return range_cdf(q, k);
    double half_df = 0.5 * df;
    double ln_c = half_df * log(half_df) - ln_gamma(half_df);
    double u_max = 1.0 + 12.0 * sqrt(2.0 / df);
    int panels = df < 6.0 ? 96 : 24;
    double grade = df < 6.0 ? 3.0 : 2.0;
    double acc = 0.0, a = 0.0;
    for (int p = 1; p <= panels; p++) {
        double b = u_max * pow((double)p / (double)panels, grade);
        double half = 0.5 * (b - a), mid = 0.5 * (b + a);
        for (int i = 0; i < 16; i++) {
            double dx = half * GL32_X[i];
            acc += half * GL32_W[i] * (srange_g(mid + dx, q, k, half_df, ln_c) + srange_g(mid - dx, q, k, half_df, ln_c));
        }
        a = b;
    }
    return acc < 1.0 ? acc : 1.0;
}
double orc_q_tukey_ppf(double p, double k, double df) {
    if (!isfinite(k) || k <= 1.0 || !isfinite(df) || df < 2.0 || !isfinite(p)) return NAN;
    if (p <= 0.0) return 0.0;
    if (p >= 1.0) return INFINITY;
    double m = k * (k - 1.0);
    double t_q = orc_t_ppf(1.0 - (1.0 - p) / m, df);
    double q = (isfinite(t_q) && t_q > 0.0) ? M_SQRT2 * t_q : 3.0;
    double lo = 0.0, hi = (q > 1.0 ? q : 1.0) * 2.0 + 4.0;
    while (studentized_range_cdf(hi, k, df) < p && hi < 1.0e3) hi *= 1.5;
    const double h = 1e-4;
    for (int it = 0; it < 60; it++) {
        double f = studentized_range_cdf(q, k, df) - p;
        if (fabs(f) < 1e-9) return q;
        if (f < 0.0) lo = q; else hi = q;
        double deriv = (studentized_range_cdf(q + h, k, df) - studentized_range_cdf(q - h, k, df)) / (2.0 * h);
        double next = (isfinite(deriv) && deriv > 1e-12) ? q - f / deriv : NAN;
        if (!isfinite(next) || next <= lo || next >= hi) next = 0.5 * (lo + hi);
        if (fabs(next - q) < 1e-10) return next;
        q = next;
    }
    return q;
}

/* marginals/t3.rs:23-93 — censored t(3) PPF table, unit variance */
#define T3_RES 2048
static double g_t3[T3_RES];
static int g_t3_ready = 0;
static const double T3_PMIN = 0.00121, T3_PMAX = 0.99879;
static void t3_build(void) {
#pragma omp critical(orc_t3)
    {
        if (!g_t3_ready) {
            double denom = (double)(T3_RES - 1);
            for (int i = 0; i < T3_RES; i++) {
                double p = T3_PMIN + (T3_PMAX - T3_PMIN) * (double)i / denom;
                g_t3[i] = orc_t_ppf(p, 3.0);
            }
            double du = (T3_PMAX - T3_PMIN) / denom;
            double m1 = T3_PMIN * g_t3[0] + (1.0 - T3_PMAX) * g_t3[T3_RES - 1];
            double m2 = T3_PMIN * g_t3[0] * g_t3[0] + (1.0 - T3_PMAX) * g_t3[T3_RES - 1] * g_t3[T3_RES - 1];
            for (int i = 0; i + 1 < T3_RES; i++) {
                double a = g_t3[i], b = g_t3[i + 1];
                m1 += du * (a + b) / 2.0;
                m2 += du * (a * a + a * b + b * b) / 3.0;
            }
            double sd = sqrt(m2 - m1 * m1);
            for (int i = 0; i < T3_RES; i++) g_t3[i] /= sd;
            g_t3_ready = 1;
        }
    }
}
double orc_t3_lookup(double percentile) {
    if (!g_t3_ready) t3_build();
    double p = percentile < T3_PMIN ? T3_PMIN : (percentile > T3_PMAX ? T3_PMAX : percentile);
    double scale = (double)(T3_RES - 1) / (T3_PMAX - T3_PMIN);
    double idx = (p - T3_PMIN) * scale;
    int ii = (int)floor(idx);
    double frac = idx - (double)ii;
    if (ii + 1 >= T3_RES) return g_t3[T3_RES - 1];
    return g_t3[ii] * (1.0 - frac) + g_t3[ii + 1] * frac;
}

/* data_gen.rs:1106-1182 (scalar contract form; ln_owned has a 2^-11 input floor,
 * data_gen.rs:1137-1141 — irrelevant after min(EXP_CAP) but kept). */
static double ln_floor(double u) {
    const double FL = 1.0 / 2048.0;
    return orc_log(u < FL ? FL : u);
}
double orc_apply_marginal(double z, int dist, double param) {
    switch (dist) {
        case ORC_DIST_NORMAL: return z;
        case ORC_DIST_BINARY: return orc_phi(z) < 1.0 - param ? 0.0 : 1.0;
        case ORC_DIST_RIGHT_SKEWED: {
            double e = -ln_floor(orc_phi(-z));
            if (e > EXP_CAP) e = EXP_CAP;
            return (e - EXP_CENSORED_MEAN) / EXP_CENSORED_STD;
        }
        case ORC_DIST_LEFT_SKEWED: {
            double e = -ln_floor(orc_phi(z));
            if (e > EXP_CAP) e = EXP_CAP;
            return (EXP_CENSORED_MEAN - e) / EXP_CENSORED_STD;
        }
        case ORC_DIST_HIGH_KURTOSIS: return orc_t3_lookup(orc_phi(z));
        case ORC_DIST_UNIFORM: return -SQRT3 + (SQRT3 - (-SQRT3)) * orc_phi(z);
        case ORC_DIST_UPLOADED_BINARY: return orc_phi(z) < 1.0 - param ? 0.0 : 1.0; /* data_gen.rs:1162-1175 */
        default: return z;
    }
}

/* data_gen.rs:1262-1277 — type-7 linear-interpolation quantile of column `col` of the sorted (rows x cols) table */
static double empirical_quantile(const double* table, int rows, int cols, int col, double u) {
    if (rows <= 0) return 0.0;
    const double pos = u * (double)(rows - 1);
    const long lo = (long)floor(pos);
    const long hi = lo + 1 < rows - 1 ? lo + 1 : rows - 1;
    const double frac = pos - (double)lo;
    const double val_lo = table[(size_t)lo * cols + col], val_hi = table[(size_t)hi * cols + col];
    return val_lo + frac * (val_hi - val_lo);
}
double orc_apply_marginal_upload(double z, int dist, double param, const double* table, int rows, int cols, int col) {
    if (dist == ORC_DIST_UPLOADED_DATA) { /* data_gen.rs:1156-1161: u = clamp(phi(z), 0, 1) */
        double u = orc_phi(z);
        u = u < 0.0 ? 0.0 : (u > 1.0 ? 1.0 : u);
        return empirical_quantile(table, rows, cols, col, u);
    }
    return orc_apply_marginal(z, dist, param);
}

/* rng.rs:399-419 */
double orc_poisson_quantile(double z, double lambda) {
    if (!(lambda > 0.0)) return 0.0;
    if (lambda <= 20.0) {
        double u = orc_phi(z);
        double p = orc_exp(-lambda), cdf = p;
        uint32_t k = 0;
        while (cdf < u && k < 200) {
            k++;
            p *= lambda / (double)k;
            cdf += p;
        }
        return (double)k;
    }
    double v = floor(lambda + z * sqrt(lambda) + 0.5);
    return v > 0.0 ? v : 0.0;
}

/* distributions.rs:78-97 */
static double sample_t(SeqRng* r, double df) {
    if (!isfinite(df) || df <= 0.0) return (double)seq_normal(r);
    double z = (double)seq_normal(r);
    long df_int = (long)round(df); /* f64::round — ties away from zero */
    if (df_int < 1) df_int = 1;
    double chi2 = 0.0;
    for (long i = 0; i < df_int; i++) {
        double g = (double)seq_normal(r);
        chi2 += g * g;
    }
    double denom = sqrt(chi2 / df);
    return denom <= 0.0 ? z : z / denom;
}
void orc_sample_t(uint64_t base_seed, uint64_t sim_id, double df, int n, double* out) {
    SeqRng r;
    seq_init(&r, base_seed, sim_id);
    for (int i = 0; i < n; i++) out[i] = sample_t(&r, df);
}

/* data_gen.rs:1289-1320 */
static double draw_residual(SeqRng* r, int dist, double df_in) {
    switch (dist) {
        case ORC_RES_HIGH_KURTOSIS: {
            double df = df_in > 3.0 ? df_in : 3.0;
            double scale = 1.0 / sqrt(df / (df - 2.0));
            return sample_t(r, df) * scale;
        }
        case ORC_RES_RIGHT_SKEWED:
        case ORC_RES_LEFT_SKEWED: {
            double df = df_in > 3.0 ? df_in : 3.0;
            long df_int = (long)round(df);
            if (df_int < 1) df_int = 1;
            double chi2 = 0.0;
            for (long i = 0; i < df_int; i++) {
                double g = (double)seq_normal(r);
                chi2 += g * g;
            }
            double centered = (chi2 - df) * (1.0 / sqrt(2.0 * df));
            return dist == ORC_RES_LEFT_SKEWED ? -centered : centered;
        }
        case ORC_RES_UNIFORM: return (2.0 * (double)seq_uniform(r) - 1.0) * SQRT3;
        default: return (double)seq_normal(r);
    }
}

void orc_draw_residual(uint64_t base_seed, uint64_t sim_id, int dist, double df, int n, double* out) {
    SeqRng r;
    seq_init(&r, base_seed, sim_id);
    for (int i = 0; i < n; i++) out[i] = draw_residual(&r, dist, df);
}

/* ======================================================================= */
/* Critical-value table (critvals.rs:827-1012, batch.rs:2334-2407)          */
/* ======================================================================= */

static double tukey_k_for_col(const OrcSpec* s, int col) { /* batch.rs:2233-2250 */
    int base = 1 + s->n_non_factor;
    if (col < base) return NAN;
    int cursor = base;
    for (int f = 0; f < s->n_factors; f++) {
        int l = s->factor_n_levels[f] > 0 ? s->factor_n_levels[f] : 0;
        int width = l > 0 ? l - 1 : 0;
        if (col < cursor + width) return (double)l;
        cursor += width;
    }
    return NAN;
}

int orc_build_crit_table(const OrcSpec* s, const uint32_t* sample_sizes, int n_ss, int n_predictors, double* t_unc_sq,
                         double* corr_sq, double* posthoc_sq, double* overall) {
    double alpha = s->alpha;
    if (!isfinite(alpha) || alpha <= 0.0 || alpha >= 1.0) return fail(ORC_ERR_INVALID_SPEC, "alpha must lie in (0, 1)");
    double pa = isnan(s->posthoc_alpha) ? alpha : s->posthoc_alpha;
    if (!isfinite(pa) || pa <= 0.0 || pa >= 1.0) return fail(ORC_ERR_INVALID_SPEC, "posthoc_alpha must lie in (0, 1)");
    int use_t = s->estimator == ORC_EST_OLS; /* EstimatorSpec::uses_student_t */
    int m = s->n_marginal_targets + s->n_contrasts;
    for (int i = 0; i < n_ss; i++) {
        long df_signed = (long)sample_sizes[i] - (long)n_predictors;
        if (df_signed < 1) return fail(ORC_ERR_INVALID_SPEC, "df = N - P < 1");
        double df = (double)df_signed;
#define QUANT(p) (use_t ? orc_t_ppf((p), df) : orc_norm_ppf((p)))
        double t_unc = QUANT(1.0 - alpha / 2.0);
        t_unc_sq[i] = t_unc * t_unc;
        double* row = corr_sq ? corr_sq + (size_t)i * m : NULL;
        if (row) {
            switch (s->correction_method) {
                case ORC_CORR_NONE:
                    for (int k = 0; k < m; k++) row[k] = t_unc * t_unc;
                    break;
                case ORC_CORR_BONFERRONI: {
                    double b = QUANT(1.0 - alpha / (2.0 * (double)m));
                    for (int k = 0; k < m; k++) row[k] = b * b;
                    break;
                }
                case ORC_CORR_HOLM:
                    for (int k = 0; k < m; k++) {
                        double eff = alpha / (2.0 * (double)(m - k));
                        double t = eff < 1e-12 ? INFINITY : QUANT(1.0 - eff);
                        row[k] = t * t;
                    }
                    break;
                case ORC_CORR_BH:
                    for (int k = 0; k < m; k++) {
                        double eff = (double)(k + 1) / (double)m * alpha / 2.0;
                        double t = eff < 1e-12 ? INFINITY : QUANT(1.0 - eff);
                        row[k] = t * t;
                    }
                    break;
                case ORC_CORR_TUKEY: {  /* one studentized-range quantile per distinct k in the row (~0.2 s each) */
                    double last_k = -1.0, last_q = NAN;
                    for (int k = 0; k < m; k++) {
                        int col = k < s->n_marginal_targets ? s->target_indices[k]
                                                            : s->contrast_pairs[2 * (k - s->n_marginal_targets)];
                        double kk = tukey_k_for_col(s, col);
                        if (!(kk == last_k)) { last_q = orc_q_tukey_ppf(1.0 - alpha, kk, df); last_k = kk; }
                        double t = last_q / M_SQRT2;
                        row[k] = t * t;
                    }
                    break;
                }
                default: return fail(ORC_ERR_INVALID_SPEC, "unknown correction method");
            }
        }
        if (posthoc_sq) {
            double tp = QUANT(1.0 - pa / 2.0);
            posthoc_sq[i] = tp * tp;
        }
#undef QUANT
        if (overall) {
            double npni = (double)(n_predictors - 1 > 0 ? n_predictors - 1 : 0);
            if (npni < 1.0) overall[i] = INFINITY;
            else if (s->estimator == ORC_EST_OLS) overall[i] = orc_f_ppf(1.0 - alpha, npni, df);
            else if (s->estimator == ORC_EST_GLM) overall[i] = orc_chi2_ppf(1.0 - alpha, npni);
            else overall[i] = INFINITY;
        }
    }
    return ORC_OK;
}

int orc_posthoc_corr_rows(const OrcSpec* s, const uint32_t* sample_sizes, int n_ss, int n_predictors, double* out) {
    double pa = isnan(s->posthoc_alpha) ? s->alpha : s->posthoc_alpha;
    if (!isfinite(pa) || pa <= 0.0 || pa >= 1.0) return fail(ORC_ERR_INVALID_SPEC, "posthoc_alpha must lie in (0, 1)");
    int use_t = s->estimator == ORC_EST_OLS;
    for (int i = 0; i < n_ss; i++) {
        long df_signed = (long)sample_sizes[i] - (long)n_predictors;
        if (df_signed < 1) return fail(ORC_ERR_INVALID_SPEC, "df = N - P < 1");
        double df = (double)df_signed;
        double* row = out + (size_t)i * s->n_posthoc;
        int o = 0;
#define QUANT(p) (use_t ? orc_t_ppf((p), df) : orc_norm_ppf((p)))
        for (int b = 0; b < s->n_posthoc_blocks; b++) {
            int m = s->posthoc_block_ncontrasts[b];
            double t_tukey = s->correction_method == ORC_CORR_TUKEY
                                 ? orc_q_tukey_ppf(1.0 - pa, (double)s->posthoc_block_levels[b], df) / M_SQRT2 : 0.0;
            for (int k = 0; k < m; k++) {
                double t;
                switch (s->correction_method) {
                    case ORC_CORR_NONE: t = QUANT(1.0 - pa / 2.0); break;
                    case ORC_CORR_BONFERRONI: t = QUANT(1.0 - pa / (2.0 * (double)m)); break;
                    case ORC_CORR_HOLM: {
                        double eff = pa / (2.0 * (double)(m - k));
                        t = eff < 1e-12 ? INFINITY : QUANT(1.0 - eff);
                        break;
                    }
                    case ORC_CORR_BH: {
                        double eff = (double)(k + 1) / (double)m * pa / 2.0;
                        t = eff < 1e-12 ? INFINITY : QUANT(1.0 - eff);
                        break;
                    }
                    case ORC_CORR_TUKEY: t = t_tukey; break;
                    default: return fail(ORC_ERR_INVALID_SPEC, "unknown correction method");
                }
                row[o++] = t * t;
            }
        }
#undef QUANT
    }
    return ORC_OK;
}

/* ======================================================================= */
/* Corrections (correction.rs:16-115)                                       */
/* ======================================================================= */

static void sort_indices_desc(const double* t_sq, int n, int* idx) {
    for (int i = 0; i < n; i++) idx[i] = i;
    /* stable insertion sort: desc by value, NaN last, ties by ascending index (== total_cmp on non-NaN) */
    for (int i = 1; i < n; i++) {
        int v = idx[i], j = i - 1;
        while (j >= 0) {
            double a = t_sq[idx[j]], b = t_sq[v];
            int a_nan = isnan(a), b_nan = isnan(b);
            int move; /* does idx[j] sort AFTER v? */
            if (a_nan && b_nan) move = 0;
            else if (a_nan) move = 1;
            else if (b_nan) move = 0;
            else move = a < b || (a == b && signbit(a) && !signbit(b)); /* total_cmp: -0 < +0 */
            if (!move) break;
            idx[j + 1] = idx[j];
            j--;
        }
        idx[j + 1] = v;
    }
}

void orc_apply_correction(int method, const double* t_sq, int t, double crit_unc, const double* crit_sq, uint8_t* out) {
    memset(out, 0, (size_t)t);
    if (t == 0) return;
    int order[256];
    switch (method) {
        case ORC_CORR_NONE:
            for (int i = 0; i < t; i++) out[i] = !isnan(t_sq[i]) && t_sq[i] > crit_unc;
            break;
        case ORC_CORR_BONFERRONI:
        case ORC_CORR_TUKEY:
            for (int i = 0; i < t; i++) out[i] = !isnan(t_sq[i]) && t_sq[i] > crit_sq[i];
            break;
        case ORC_CORR_HOLM:
            sort_indices_desc(t_sq, t, order);
            for (int k = 0; k < t; k++) {
                double v = t_sq[order[k]];
                if (!isnan(v) && v > crit_sq[k]) out[order[k]] = 1;
                else break;
            }
            break;
        case ORC_CORR_BH: {
            sort_indices_desc(t_sq, t, order);
            int last = -1;
            for (int k = 0; k < t; k++) {
                double v = t_sq[order[k]];
                if (!isnan(v) && v > crit_sq[k]) last = k;
            }
            for (int k = 0; k <= last; k++) out[order[k]] = 1;
            break;
        }
    }
}

/* ======================================================================= */
/* Small dense linear algebra                                              */
/* ======================================================================= */

/* LLT of a k×k symmetric matrix, column-major in/out, lower triangle read
 * (correlation.rs:48-74 → faer llt; restated as the left-looking column
 * algorithm of data_gen.rs:55-75).  Fails on a non-positive pivot. */
static int llt_colmajor(double* a, int n) {
    for (int j = 0; j < n; j++) {
        double diag = a[j * n + j];
        for (int k = 0; k < j; k++) diag -= a[k * n + j] * a[k * n + j];
        if (!(diag > 0.0) || !isfinite(diag)) return -1;
        double ljj = sqrt(diag);
        a[j * n + j] = ljj;
        for (int i = j + 1; i < n; i++) {
            double s = a[j * n + i];
            for (int k = 0; k < j; k++) s -= a[k * n + i] * a[k * n + j];
            a[j * n + i] = s / ljj;
        }
    }
    for (int j = 0; j < n; j++)
        for (int i = 0; i < j; i++) a[j * n + i] = 0.0;
    return 0;
}

/* Cyclic Jacobi eigen-decomposition of a symmetric n×n matrix (row-major a,
 * destroyed); eigenvalues in w, eigenvectors as columns of v.  Stand-in for
 * faer's self_adjoint_eigen (correlation.rs:120-127) — UNPINNED at bit level. */
static void jacobi_eigh(double* a, int n, double* w, double* v) {
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) v[i * n + j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0.0;
        for (int i = 0; i < n; i++)
            for (int j = i + 1; j < n; j++) off += a[i * n + j] * a[i * n + j];
        if (off < 1e-300) break;
        for (int p = 0; p < n; p++)
            for (int q = p + 1; q < n; q++) {
                double apq = a[p * n + q];
                if (fabs(apq) < 1e-300) continue;
                double theta = (a[q * n + q] - a[p * n + p]) / (2.0 * apq);
                double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; k++) {
                    double akp = a[k * n + p], akq = a[k * n + q];
                    a[k * n + p] = c * akp - s * akq;
                    a[k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; k++) {
                    double apk = a[p * n + k], aqk = a[q * n + k];
                    a[p * n + k] = c * apk - s * aqk;
                    a[q * n + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; k++) {
                    double vkp = v[k * n + p], vkq = v[k * n + q];
                    v[k * n + p] = c * vkp - s * vkq;
                    v[k * n + q] = s * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < n; i++) w[i] = a[i * n + i];
}

/* correlation.rs:82-189 */
static int psd_repair_and_factor(double* a /* col-major n×n */, int n) {
    double* work = (double*)malloc(sizeof(double) * (size_t)n * n * 4 + sizeof(double) * (size_t)n * 2);
    double* copy = work;
    memcpy(copy, a, sizeof(double) * (size_t)n * n);
    if (llt_colmajor(copy, n) == 0) {
        memcpy(a, copy, sizeof(double) * (size_t)n * n);
        free(work);
        return 0;
    }
    double* sym = work + n * n;
    double* vec = sym + n * n;
    double* rep = vec + n * n;
    double* lam = rep + n * n;
    double* d = lam + n;
    for (int j = 0; j < n; j++)
        for (int i = 0; i < n; i++) sym[i * n + j] = i >= j ? a[j * n + i] : a[i * n + j];
    jacobi_eigh(sym, n, lam, vec);
    for (int k = 0; k < n; k++)
        if (!(lam[k] > 1e-10)) lam[k] = 1e-10;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            double acc = 0.0;
            for (int k = 0; k < n; k++) acc += vec[i * n + k] * lam[k] * vec[j * n + k];
            rep[i * n + j] = acc;
        }
    for (int i = 0; i < n; i++) d[i] = rep[i * n + i] > 0.0 ? sqrt(rep[i * n + i]) : 1.0;
    for (int j = 0; j < n; j++)
        for (int i = 0; i < n; i++) a[j * n + i] = rep[i * n + j] / (d[i] * d[j]);
    for (int i = 0; i < n; i++) a[i * n + i] = 1.0;
    int rc = llt_colmajor(a, n);
    free(work);
    return rc;
}

/* ======================================================================= */
/* Data generation (data_gen.rs:87-925)                                     */
/* ======================================================================= */

typedef struct {
    int max_n, P, k;
    float* x;          /* col-major max_n × P */
    float* y;
    float* z_planar;   /* k × max_n */
    float* residuals;  /* max_n */
    float* resid_g;
    uint32_t* words;
    double* mix_acc;
    double* eta_acc;
    double* l_design;  /* k×k col-major */
    int* var_types_design;
    int residual_dist_design;
    double residual_df_design;
    double tau_squared_design;
    float* cluster_u;
    float* cluster_u1; /* random-slope effects */
    float* extra_u[2]; /* extra-grouping random intercepts */
    double extra_tau_sq_design[2];
    int32_t* cluster_ids;
    int n_cluster_cap;
    uint32_t* alloc_counts;
    double* beta_eff;
    /* OLS fit scratch */
    double* xtx; /* packed lower, row-major: idx(i,j)=i(i+1)/2+j */
    double* xty;
    uint32_t* prefix_counts;
    uint8_t* excl_flags;
    /* reduced-model thresholds memoised per (N, p_red) like reduced_crit_entry (batch.rs:2418-2437) */
    struct RedCrit { uint32_t n; int p_red; double unc, ov, ph; double corr[256], phc[64]; } * red_memo;
    int n_red_memo;
} Ws;

static Ws* ws_new(const OrcSpec* s, int max_n) {
    Ws* w = (Ws*)calloc(1, sizeof(Ws));
    int k = s->n_non_factor, P = 1 + k + s->n_factor_dummies + s->n_interactions;
    w->max_n = max_n; w->P = P; w->k = k;
    w->x = (float*)calloc((size_t)max_n * P, sizeof(float));
    w->y = (float*)calloc((size_t)max_n, sizeof(float));
    w->z_planar = (float*)calloc((size_t)max_n * (k > 0 ? k : 1), sizeof(float));
    w->residuals = (float*)calloc((size_t)max_n, sizeof(float));
    w->resid_g = (float*)calloc((size_t)max_n, sizeof(float));
    w->words = (uint32_t*)calloc((size_t)max_n + 4, sizeof(uint32_t));
    w->mix_acc = (double*)calloc((size_t)max_n, sizeof(double));
    w->eta_acc = (double*)calloc((size_t)max_n, sizeof(double));
    w->l_design = (double*)calloc((size_t)(k > 0 ? k * k : 1), sizeof(double));
    w->var_types_design = (int*)calloc((size_t)(k > 0 ? k : 1), sizeof(int));
    int nprops = 0;
    for (int f = 0; f < s->n_factors; f++) nprops += s->factor_n_levels[f];
    w->alloc_counts = (uint32_t*)calloc((size_t)(nprops > 0 ? nprops : 1), sizeof(uint32_t));
    w->prefix_counts = (uint32_t*)calloc((size_t)(nprops > 0 ? nprops : 1), sizeof(uint32_t));
    w->excl_flags = (uint8_t*)calloc((size_t)(s->n_factors > 0 ? s->n_factors : 1), 1);
    w->beta_eff = (double*)calloc((size_t)P, sizeof(double));
    w->xtx = (double*)calloc((size_t)P * (P + 1) / 2, sizeof(double));
    w->xty = (double*)calloc((size_t)P, sizeof(double));
    if (s->has_cluster) {
        int nc = s->sizing_fixed_size ? max_n / (s->sizing_value > 0 ? s->sizing_value : 1) + 1
                                      : (s->sizing_value > 0 ? s->sizing_value : 1);
        w->n_cluster_cap = nc;
        w->cluster_u = (float*)calloc((size_t)nc, sizeof(float));
        w->cluster_u1 = (float*)calloc((size_t)nc * 2, sizeof(float));
        for (int g = 0; g < s->n_extra; g++) w->extra_u[g] = (float*)calloc((size_t)orc_extra_levels_at(s, g, max_n) + 1, sizeof(float));
        w->cluster_ids = (int32_t*)calloc((size_t)max_n, sizeof(int32_t));
        for (int i = 0; i < max_n; i++) /* generation.rs cluster_of_row */
            w->cluster_ids[i] = s->sizing_fixed_size ? i / (s->sizing_value > 0 ? s->sizing_value : 1)
                                                     : i % (s->sizing_value > 0 ? s->sizing_value : 1);
    }
    return w;
}
static void ws_free(Ws* w) {
    if (!w) return;
    free(w->x); free(w->y); free(w->z_planar); free(w->residuals); free(w->resid_g); free(w->words);
    free(w->mix_acc); free(w->eta_acc); free(w->l_design); free(w->var_types_design); free(w->alloc_counts);
    free(w->prefix_counts); free(w->excl_flags); free(w->beta_eff); free(w->xtx); free(w->xty);
    free(w->cluster_u); free(w->cluster_u1); free(w->extra_u[0]); free(w->extra_u[1]); free(w->cluster_ids); free(w->red_memo);
    free(w);
}

static int is_optimistic(const OrcSpec* s) { /* spec.rs:229-236 */
    return s->heterogeneity == 0.0 && s->hsk_ratio == 1.0 && s->corr_noise_sd == 0.0 && s->dist_change_prob == 0.0 &&
           s->resid_change_prob == 0.0 && !s->has_lme;
}

/* generation.rs:370-430: one lexicographic stride chain, primary fastest, extras in declaration order */
int orc_extra_level_of_row(const OrcSpec* s, int g, int i) {
    const int bl = s->extra_count[g] > 0 ? s->extra_count[g] : 1;
    const int sv = s->sizing_value > 0 ? s->sizing_value : 1;
    if (!s->sizing_fixed_size) {
        long stride = sv;
        for (int h = 0; h < g; h++) stride *= s->extra_count[h] > 0 ? s->extra_count[h] : 1;
        const int within = (int)(((long)i / stride) % bl);
        return s->extra_nested[g] ? (i % sv) * bl + within : within;
    }
    return (i / sv) * bl + (i % sv) % bl; /* FixedSize: nested only (validate() rejects crossed) */
}
int orc_extra_levels_at(const OrcSpec* s, int g, int n) { /* generation.rs:301-320 */
    const int bl = s->extra_count[g] > 0 ? s->extra_count[g] : 1;
    if (!s->extra_nested[g]) return bl;
    const int sv = s->sizing_value > 0 ? s->sizing_value : 1;
    return (s->sizing_fixed_size ? (n + sv - 1) / sv : sv) * bl;
}

/* data_gen.rs:990-1093 + scenarios.rs:67-186 */
static int populate_design(const OrcSpec* s, uint64_t base_seed, uint64_t sim_id, Ws* w) {
    int k = w->k;
    if (is_optimistic(s)) {
        memcpy(w->l_design, s->correlation, sizeof(double) * (size_t)k * k);
        if (k > 0 && llt_colmajor(w->l_design, k) != 0) return fail(ORC_ERR_NOT_PSD, "correlation matrix is not positive semi-definite");
        for (int j = 0; j < k; j++) w->var_types_design[j] = s->var_types[j];
        w->residual_dist_design = s->residual_dist;
        w->residual_df_design = s->residual_df;
        w->tau_squared_design = s->has_cluster ? s->tau_squared : 0.0;
        for (int g = 0; g < s->n_extra; g++) w->extra_tau_sq_design[g] = s->extra_tau_squared[g];
        return ORC_OK;
    }
    SeqRng sr;
    seq_init(&sr, base_seed, sim_id ^ 0x5C5C5C5C5C5C5C5CULL); /* scenarios.rs:19,50 */
    /* perturb_correlation (scenarios.rs:67-120) */
    if (s->corr_noise_sd == 0.0) {
        memcpy(w->l_design, s->correlation, sizeof(double) * (size_t)k * k);
    } else {
        double* noise = (double*)malloc(sizeof(double) * (size_t)(k * k + 1));
        for (int j = 0; j < k; j++)
            for (int i = 0; i < k; i++) noise[j * k + i] = (double)seq_normal(&sr) * s->corr_noise_sd;
        for (int j = 0; j < k; j++)
            for (int i = 0; i < k; i++) {
                double sym = 0.5 * (noise[j * k + i] + noise[i * k + j]);
                double v = s->correlation[j * k + i] + sym;
                w->l_design[j * k + i] = v < -0.8 ? -0.8 : (v > 0.8 ? 0.8 : v);
            }
        for (int i = 0; i < k; i++) w->l_design[i * k + i] = 1.0;
        free(noise);
    }
    if (k > 0 && psd_repair_and_factor(w->l_design, k) != 0) return fail(ORC_ERR_NOT_PSD, "correlation matrix is not positive semi-definite");
    /* perturb_var_types (scenarios.rs:127-156) */
    for (int j = 0; j < k; j++) w->var_types_design[j] = s->var_types[j];
    if (s->dist_change_prob > 0.0 && s->n_new_dists > 0) {
        for (int j = 0; j < k; j++) {
            double u = (double)seq_uniform(&sr);
            int pinned = s->var_pinned ? s->var_pinned[j] : 0;
            if (w->var_types_design[j] == ORC_DIST_NORMAL && !pinned && u < s->dist_change_prob) {
                int kk = (int)((double)seq_uniform(&sr) * (double)s->n_new_dists);
                if (kk > s->n_new_dists - 1) kk = s->n_new_dists - 1;
                w->var_types_design[j] = s->new_dists[kk];
            }
        }
    }
    /* pick_residual (scenarios.rs:163-186) */
    w->residual_dist_design = s->residual_dist;
    if (s->resid_change_prob > 0.0 && s->n_resid_dists > 0) {
        double u = (double)seq_uniform(&sr);
        if (u < s->resid_change_prob && s->residual_dist == ORC_RES_NORMAL && !s->residual_pinned) {
            int kk = (int)((double)seq_uniform(&sr) * (double)s->n_resid_dists);
            if (kk > s->n_resid_dists - 1) kk = s->n_resid_dists - 1;
            w->residual_dist_design = s->resid_dists[kk];
        }
    }
    w->residual_df_design = s->residual_df;
    double base_tau = s->has_cluster ? s->tau_squared : 0.0;
    if (s->has_lme && s->icc_noise_sd > 0.0) {
        double jit = (double)seq_normal(&sr) * s->icc_noise_sd;
        double t = base_tau + jit;
        w->tau_squared_design = t > 0.0 ? t : 0.0;
    } else {
        w->tau_squared_design = base_tau;
    }
    for (int g = 0; g < s->n_extra; g++) { /* one independent jitter draw per extra grouping, declaration order (data_gen.rs:1077-1091) */
        if (s->has_lme && s->icc_noise_sd > 0.0) {
            double t = s->extra_tau_squared[g] + (double)seq_normal(&sr) * s->icc_noise_sd;
            w->extra_tau_sq_design[g] = t > 0.0 ? t : 0.0;
        } else {
            w->extra_tau_sq_design[g] = s->extra_tau_squared[g];
        }
    }
    return ORC_OK;
}

/* data_gen.rs:943-961 */
static int fixed_level_next(const double* probs, int n, uint32_t* counts) {
    double total = 0.0;
    for (int g = 0; g < n; g++) total += probs[g];
    uint32_t assigned = 0;
    for (int g = 0; g < n; g++) assigned += counts[g];
    double t = (double)(assigned + 1);
    int best = 0;
    double best_def = -INFINITY;
    for (int g = 0; g < n; g++) {
        double def = (probs[g] / total) * t - (double)counts[g];
        if (def > best_def) { best_def = def; best = g; }
    }
    counts[best]++;
    return best;
}

/* spec.rs:436-515 */
static void het_coeffs(const OrcSpec* s, double* lp_mean, double* lp_std, double* col_mean, double* col_std) {
    int k = s->n_non_factor, fd = s->n_factor_dummies, p = 1 + k + fd;
    col_mean[0] = 1.0; col_std[0] = 0.0;
    for (int j = 0; j < k; j++) {
        if (s->var_types[j] == ORC_DIST_BINARY) {
            double pp = s->var_params[j];
            col_mean[1 + j] = pp;
            double v = pp * (1.0 - pp);
            col_std[1 + j] = sqrt(v > 0.0 ? v : 0.0);
        } else { col_mean[1 + j] = 0.0; col_std[1 + j] = 1.0; }
    }
    int col = 1 + k, off = 0;
    for (int f = 0; f < s->n_factors; f++) {
        int nl = s->factor_n_levels[f] > 0 ? s->factor_n_levels[f] : 0;
        for (int d = 0; d + 1 < nl; d++) {
            double pp = s->factor_proportions[off + d + 1];
            col_mean[col + d] = pp;
            double v = pp * (1.0 - pp);
            col_std[col + d] = sqrt(v > 0.0 ? v : 0.0);
        }
        col += nl > 0 ? nl - 1 : 0;
        off += nl;
    }
    double m = 0.0, var = 0.0;
    for (int j = 0; j < p; j++) m += s->effect_sizes[j] * col_mean[j];
    for (int j = 0; j < p; j++) var += s->effect_sizes[j] * s->effect_sizes[j] * col_std[j] * col_std[j];
    for (int j = 0; j < k; j++)
        for (int kk = j + 1; kk < k; kk++)
            var += 2.0 * s->effect_sizes[1 + j] * s->effect_sizes[1 + kk] * s->correlation[kk * k + j] * col_std[1 + j] * col_std[1 + kk];
    *lp_mean = m;
    *lp_std = sqrt(var > 0.0 ? var : 0.0);
}

static int gen_sim(const OrcSpec* s, uint64_t base_seed, uint64_t sim_id, Ws* w) {
    const int max_n = w->max_n, P = w->P, k = w->k;
    int rc = populate_design(s, base_seed, sim_id, w);
    if (rc) return rc;

    SeqRng rng;
    seq_init(&rng, base_seed, sim_id);
    /* 2a. cluster random intercepts (data_gen.rs:146-161) */
    if (s->has_cluster) {
        double tau = sqrt(w->tau_squared_design > 0.0 ? w->tau_squared_design : 0.0);
        int re_dist = s->has_lme ? s->re_dist : ORC_RES_NORMAL;
        double re_df = s->has_lme ? s->re_df : 0.0;
        int nc = s->sizing_fixed_size ? max_n / (s->sizing_value > 0 ? s->sizing_value : 1)
                                      : (s->sizing_value > 0 ? s->sizing_value : 1);
        for (int c = 0; c < nc; c++) w->cluster_u[c] = (float)(draw_residual(&rng, re_dist, re_df) * tau);
        /* 2a'. primary random slope (data_gen.rs:163-213): per cluster z_1 ~ draw_residual AFTER all the intercept draws; z_0 is
         * recovered from the f32 intercept effect; (u_0, u_1) = chol(D) z with D = diag(tau) R diag(tau) */
        if (s->n_slopes > 0) {
            const int nsl = s->n_slopes, q = 1 + nsl;
            double tv[3], R[9], D[9], L[9], z[3];
            tv[0] = tau;
            for (int k = 0; k < nsl; k++) tv[1 + k] = sqrt(s->slope_variance[k] > 0.0 ? s->slope_variance[k] : 0.0);
            for (int i = 0; i < q * q; i++) R[i] = 0.0;
            for (int d = 0; d < q; d++) R[d * q + d] = 1.0;
            for (int k = 0; k < nsl; k++) R[k + 1] = R[(k + 1) * q] = s->slope_corr[k];
            if (nsl == 2) R[1 * q + 2] = R[2 * q + 1] = s->slope_corr12;
            for (int i = 0; i < q; i++)
                for (int j = 0; j < q; j++) D[i * q + j] = tv[i] * R[i * q + j] * tv[j];
            for (int i = 0; i < q * q; i++) L[i] = 0.0;
            for (int j = 0; j < q; j++) { /* chol_lower (data_gen.rs:55-75) */
                double diag = D[j * q + j];
                for (int k = 0; k < j; k++) diag -= L[j * q + k] * L[j * q + k];
                const double ljj = sqrt(diag > 0.0 ? diag : 0.0);
                L[j * q + j] = ljj;
                for (int i = j + 1; i < q; i++)
                    if (ljj > 0.0) {
                        double sacc = D[i * q + j];
                        for (int k = 0; k < j; k++) sacc -= L[i * q + k] * L[j * q + k];
                        L[i * q + j] = sacc / ljj;
                    }
            }
            for (int c = 0; c < nc; c++) {
                z[0] = tau > 0.0 ? (double)w->cluster_u[c] / tau : 0.0;
                for (int j = 1; j < q; j++) z[j] = draw_residual(&rng, re_dist, re_df);
                for (int k = 1; k < q; k++) {
                    double uk = 0.0;
                    for (int j = 0; j <= k; j++) uk += L[k * q + j] * z[j];
                    w->cluster_u1[(size_t)c * nsl + (k - 1)] = (float)uk;
                }
            }
        }
        /* 2b. extra-grouping intercepts, one block per grouping behind the primary's draws (data_gen.rs:214-236) */
        for (int g = 0; g < s->n_extra; g++) {
            const double tau_g = sqrt(w->extra_tau_sq_design[g] > 0.0 ? w->extra_tau_sq_design[g] : 0.0);
            const int nl = orc_extra_levels_at(s, g, max_n);
            for (int c = 0; c < nl; c++) w->extra_u[g][c] = (float)(draw_residual(&rng, re_dist, re_df) * tau_g);
        }
    }
    for (int i = 0; i < max_n; i++) w->x[i] = 1.0f;

    int any_exact = 0;
    for (int f = 0; f < s->n_factors; f++) {
        int fs = s->factor_sampled ? s->factor_sampled[f] : -1;
        int sampled = fs < 0 ? s->sampled_factor_proportions : fs;
        if (!sampled) any_exact = 1;
    }
    if (any_exact) {
        int nprops = 0;
        for (int f = 0; f < s->n_factors; f++) nprops += s->factor_n_levels[f];
        memset(w->alloc_counts, 0, sizeof(uint32_t) * (size_t)nprops);
    }

    /* (a) continuous block (data_gen.rs:316-363) */
    for (int j = 0; j < k; j++) fill_normal_column(rng.key, 1, (uint32_t)j, max_n, w->words, w->z_planar + (size_t)j * max_n);
    for (int j = 0; j < k; j++) {
        for (int i = 0; i < max_n; i++) w->mix_acc[i] = 0.0;
        for (int m = 0; m <= j; m++) {
            double l = w->l_design[m * k + j]; /* L[j][m], col-major */
            const float* zk = w->z_planar + (size_t)m * max_n;
            for (int i = 0; i < max_n; i++) w->mix_acc[i] += l * (double)zk[i];
        }
        float* xc = w->x + (size_t)(1 + j) * max_n;
        int vt = w->var_types_design[j];
        if (vt == ORC_DIST_NORMAL) {
            for (int i = 0; i < max_n; i++) xc[i] = (float)w->mix_acc[i];
        } else {
            double param = s->var_params[j];
            for (int i = 0; i < max_n; i++)
                xc[i] = (float)orc_apply_marginal_upload((double)(float)w->mix_acc[i], vt, param, s->upload_normal, s->un_rows, s->un_cols, j);
        }
    }

    /* (c) residual columns (data_gen.rs:375-503) */
    if (s->outcome_kind == ORC_OUT_BINARY && !s->link_probit) {
        fill_uniform_column(rng.key, 2, 0, max_n, w->words, w->residuals);
    } else if (s->outcome_kind != ORC_OUT_CONTINUOUS) {
        fill_normal_column(rng.key, 2, 0, max_n, w->words, w->residuals);
    } else {
        switch (w->residual_dist_design) {
            case ORC_RES_NORMAL: fill_normal_column(rng.key, 2, 0, max_n, w->words, w->residuals); break;
            case ORC_RES_HIGH_KURTOSIS: {
                double df = w->residual_df_design > 3.0 ? w->residual_df_design : 3.0;
                long dfi = (long)round(df);
                if (dfi < 1) dfi = 1;
                double scale = 1.0 / sqrt(df / (df - 2.0));
                fill_normal_column(rng.key, 2, 0, max_n, w->words, w->residuals);
                for (int i = 0; i < max_n; i++) w->mix_acc[i] = 0.0;
                for (long sl = 0; sl < dfi; sl++) {
                    fill_normal_column(rng.key, 2, (uint32_t)(1 + sl), max_n, w->words, w->resid_g);
                    for (int i = 0; i < max_n; i++) { double g = (double)w->resid_g[i]; w->mix_acc[i] += g * g; }
                }
                for (int i = 0; i < max_n; i++) {
                    double z = (double)w->residuals[i];
                    double denom = sqrt(w->mix_acc[i] / df);
                    double t = denom <= 0.0 ? z : z / denom;
                    w->residuals[i] = (float)(t * scale);
                }
                break;
            }
            case ORC_RES_RIGHT_SKEWED:
            case ORC_RES_LEFT_SKEWED: {
                double df = w->residual_df_design > 3.0 ? w->residual_df_design : 3.0;
                long dfi = (long)round(df);
                if (dfi < 1) dfi = 1;
                double scale = 1.0 / sqrt(2.0 * df);
                double sign = w->residual_dist_design == ORC_RES_LEFT_SKEWED ? -1.0 : 1.0;
                for (int i = 0; i < max_n; i++) w->mix_acc[i] = 0.0;
                for (long sl = 0; sl < dfi; sl++) {
                    fill_normal_column(rng.key, 2, (uint32_t)sl, max_n, w->words, w->resid_g);
                    for (int i = 0; i < max_n; i++) { double g = (double)w->resid_g[i]; w->mix_acc[i] += g * g; }
                }
                for (int i = 0; i < max_n; i++) w->residuals[i] = (float)(sign * (w->mix_acc[i] - df) * scale);
                break;
            }
            case ORC_RES_UNIFORM:
                fill_uniform_column(rng.key, 2, 0, max_n, w->words, w->residuals);
                for (int i = 0; i < max_n; i++) w->residuals[i] = (float)((2.0 * (double)w->residuals[i] - 1.0) * SQRT3);
                break;
        }
    }

    /* 3. row loop: factors, interactions (data_gen.rs:509-650) */
    const int inter_base = 1 + k + s->n_factor_dummies;
    for (int i = 0; i < max_n; i++) {
        int col = 1 + k, off = 0;
        for (int f = 0; f < s->n_factors; f++) {
            int nl = s->factor_n_levels[f] > 0 ? s->factor_n_levels[f] : 0;
            if (nl == 0) continue;
            int fs = s->factor_sampled ? s->factor_sampled[f] : -1;
            int sampled = fs < 0 ? s->sampled_factor_proportions : fs;
            int lvl = !sampled ? fixed_level_next(s->factor_proportions + off, nl, w->alloc_counts + off)
                               : seq_categorical(&rng, s->factor_proportions + off, nl);
            for (int d = 0; d + 1 < nl; d++) w->x[(size_t)(col + d) * max_n + i] = 0.0f;
            if (lvl > 0 && lvl - 1 < nl - 1) w->x[(size_t)(col + lvl - 1) * max_n + i] = 1.0f;
            off += nl;
            col += nl - 1;
        }
        /* (b.4) strict bootstrap (data_gen.rs:553-604): ONE source row per design row, drawn after the factor draws */
        if (s->n_frame_map > 0) {
            const int u_rows = s->ud_rows, nc = s->ud_cols;
            size_t r = (size_t)((double)seq_uniform(&rng) * (double)u_rows);
            if (r > (size_t)(u_rows > 0 ? u_rows - 1 : 0)) r = (size_t)(u_rows > 0 ? u_rows - 1 : 0);
            for (int j = 0; j < k; j++) {
                const int fc = s->frame_map[j];
                if (fc < 0) continue;
                const double v = s->upload_data[r * (size_t)nc + fc];
                double xv = v;
                if (s->var_types[j] == ORC_DIST_UPLOADED_BINARY) { /* the ORIGINAL type, not the scenario-perturbed one */
                    const double p = j < s->n_var_params ? s->var_params[j] : 0.0;
                    xv = v + p >= 0.5 ? 1.0 : 0.0;
                }
                w->x[(size_t)(1 + j) * max_n + i] = (float)xv;
            }
            int start = 1 + k;
            for (int f = 0; f < s->n_factors; f++) {
                const int nd = s->factor_n_levels[f] > 1 ? s->factor_n_levels[f] - 1 : 0;
                const int fc = s->frame_map[k + f];
                if (fc >= 0) {
                    const double lv = round(s->upload_data[r * (size_t)nc + fc]); /* f64::round, then a saturating `as usize` */
                    const long level = lv != lv || lv <= 0.0 ? 0 : (lv > 1e9 ? 1000000000L : (long)lv);
                    for (int d = 0; d < nd; d++) w->x[(size_t)(start + d) * max_n + i] = 0.0f;
                    if (level > 0 && level - 1 < nd) w->x[(size_t)(start + level - 1) * max_n + i] = 1.0f;
                }
                start += nd;
            }
        }
        /* (b.45) cluster-level broadcast (data_gen.rs:605-640): after the factor / bootstrap block, before the interactions;
         * the representative row precedes row i in both layouts, so its value is final and resident */
        if (s->n_between > 0 && s->has_cluster) {
            const int sv = s->sizing_value > 0 ? s->sizing_value : 1;
            const int rep = s->sizing_fixed_size ? i - (i % sv) : i % sv;
            if (i != rep)
                for (int b = 0; b < s->n_between; b++)
                    w->x[(size_t)s->between_cols[b] * max_n + i] = w->x[(size_t)s->between_cols[b] * max_n + rep];
        }
        for (int j = 0; j < s->n_interactions; j++) {
            double prod = 1.0;
            for (int c = s->inter_ptr[j]; c < s->inter_ptr[j + 1]; c++) prod *= (double)w->x[(size_t)s->inter_cols[c] * max_n + i];
            w->x[(size_t)(inter_base + j) * max_n + i] = (float)prod;
        }
    }

    /* 4. y assembly (data_gen.rs:709-922) */
    double het_total = s->heterogeneity > 0.0 ? s->heterogeneity : 0.0;
    int use_het = het_total > FLOAT_NEAR_ZERO;
    double het_int_sd = s->outcome_kind == ORC_OUT_CONTINUOUS ? 0.0 : het_total;
    if (use_het) {
        SeqRng hr;
        seq_init(&hr, base_seed, sim_id ^ 0x4848484848484848ULL); /* scenarios.rs:28 */
        for (int j = 0; j < P; j++) {
            double b = s->effect_sizes[j];
            double sj = j == 0 ? het_int_sd : het_total * fabs(b);
            double raw = b + (double)seq_normal(&hr) * sj;
            if (j == 0 || b == 0.0) w->beta_eff[j] = raw;
            else if (b > 0.0) w->beta_eff[j] = raw > 0.0 ? raw : 0.0;
            else w->beta_eff[j] = raw < 0.0 ? raw : 0.0;
        }
    }
    double lambda = s->hsk_ratio > 1.0 ? s->hsk_ratio : 1.0;
    double gamma = log(lambda) / 4.0;
    double hsk_center = 0.0, hsk_std = 0.0;
    {
        int pf = 1 + k + s->n_factor_dummies;
        double* cm = (double*)malloc(sizeof(double) * (size_t)pf * 2);
        double lpm, lps;
        het_coeffs(s, &lpm, &lps, cm, cm + pf);
        if (s->hsk_driver < 0) { hsk_center = lpm; hsk_std = lps; }
        else if (s->hsk_driver < pf) { hsk_center = cm[s->hsk_driver]; hsk_std = cm[pf + s->hsk_driver]; }
        free(cm);
    }
    int apply_hsk = s->outcome_kind == ORC_OUT_CONTINUOUS && gamma > FLOAT_NEAR_ZERO && hsk_std > FLOAT_NEAR_ZERO;
    double hsk_inv_std = apply_hsk ? 1.0 / hsk_std : 0.0;
    double hsk_inv_norm = orc_exp(-gamma * gamma / 2.0);

    for (int i = 0; i < max_n; i++) {
        double lp_clean = 0.0, lp_eff = 0.0;
        for (int j = 0; j < P; j++) {
            double xij = (double)w->x[(size_t)j * max_n + i];
            lp_clean += xij * s->effect_sizes[j];
            if (use_het) lp_eff += xij * w->beta_eff[j];
        }
        double lp = use_het ? lp_eff : lp_clean;
        if (apply_hsk) {
            double driver = s->hsk_driver < 0 ? lp_clean : (double)w->x[(size_t)s->hsk_driver * max_n + i];
            double z = (driver - hsk_center) * hsk_inv_std;
            w->mix_acc[i] = gamma * z;
        }
        double u_re = s->has_cluster ? (double)w->cluster_u[w->cluster_ids[i]] : 0.0;
        if (s->has_cluster) /* extras first (data_gen.rs:818-835), then the primary slopes (836-842) */
            for (int g = 0; g < s->n_extra; g++) u_re += (double)w->extra_u[g][orc_extra_level_of_row(s, g, i)];
        if (s->has_cluster)
            for (int k = 0; k < s->n_slopes; k++)
                u_re += (double)w->cluster_u1[(size_t)w->cluster_ids[i] * s->n_slopes + k] * (double)w->x[(size_t)s->slope_col[k] * max_n + i];
        if (s->outcome_kind == ORC_OUT_CONTINUOUS && !apply_hsk) w->y[i] = (float)(lp + (double)w->residuals[i] + u_re);
        else w->eta_acc[i] = lp + u_re;
    }
    if (apply_hsk) {
        for (int i = 0; i < max_n; i++) {
            double mult = orc_exp(w->mix_acc[i]) * hsk_inv_norm;
            w->residuals[i] = (float)((double)w->residuals[i] * sqrt(mult));
        }
        for (int i = 0; i < max_n; i++) w->y[i] = (float)(w->eta_acc[i] + (double)w->residuals[i]);
    }
    if (s->outcome_kind == ORC_OUT_BINARY) {
        for (int i = 0; i < max_n; i++) {
            double thr = w->eta_acc[i];
            if (!s->link_probit) { /* stable sigmoid (glmm sigmoid_fill, <=2 ULP of libm form) */
                double e = thr;
                if (e >= 0.0) thr = 1.0 / (1.0 + orc_exp(-e));
                else { double ex = orc_exp(e); thr = ex / (1.0 + ex); }
            }
            w->y[i] = (double)w->residuals[i] < thr ? 1.0f : 0.0f;
        }
    } else if (s->outcome_kind == ORC_OUT_COUNT) {
        for (int i = 0; i < max_n; i++) w->y[i] = (float)orc_poisson_quantile((double)w->residuals[i], orc_exp(w->eta_acc[i]));
    }
    return ORC_OK;
}

/* The lower Cholesky factor populate_design leaves for one study (column-major k x k; data_gen.rs:990-1093): test hook for the
 * PSD-repair cross-check against an independent eigensolver (tests/test_oracle_golden.py). */
int orc_design_factor(const OrcSpec* s, uint64_t base_seed, uint64_t sim_id, double* l_colmajor) {
    Ws* w = ws_new(s, 4);
    int rc = populate_design(s, base_seed, sim_id, w);
    if (!rc) memcpy(l_colmajor, w->l_design, sizeof(double) * (size_t)w->k * w->k);
    ws_free(w);
    return rc;
}

static int validate_spec(const OrcSpec* s, const uint32_t* ss, int n_ss) { /* batch.rs:208-330 */
    if (s->estimator == ORC_EST_MLE && !s->has_cluster) return fail(ORC_ERR_INVALID_SPEC, "estimator Mle requires ClusterSpec");
    if (ss) {
        if (n_ss <= 0) return fail(ORC_ERR_INVALID_SPEC, "sample_sizes must be non-empty");
        for (int i = 1; i < n_ss; i++)
            if (ss[i] <= ss[i - 1]) return fail(ORC_ERR_INVALID_SPEC, "sample_sizes must be strictly increasing");
        if (ss[0] == 0) return fail(ORC_ERR_INVALID_SPEC, "sample_sizes[0] must be > 0");
    }
    int P = 1 + s->n_non_factor + s->n_factor_dummies + s->n_interactions;
    int nd = 0;
    for (int f = 0; f < s->n_factors; f++) {
        if (s->factor_n_levels[f] < 2) return fail(ORC_ERR_INVALID_SPEC, "factor_n_levels entries must be >= 2");
        nd += s->factor_n_levels[f] - 1;
    }
    if (nd != s->n_factor_dummies) return fail(ORC_ERR_INVALID_SPEC, "n_factor_dummies does not match factor_n_levels");
    for (int t = 0; t < s->n_marginal_targets; t++)
        if (s->target_indices[t] < 0 || s->target_indices[t] >= P) return fail(ORC_ERR_INVALID_SPEC, "target_indices entry out of range");
    for (int c = 0; c < 2 * s->n_contrasts; c++)
        if (s->contrast_pairs[c] < 0 || s->contrast_pairs[c] >= P) return fail(ORC_ERR_INVALID_SPEC, "contrast_pairs entry out of range");
    int k = s->n_non_factor;
    if (s->n_slopes > 0) {
        if (!s->has_cluster || s->n_slopes > 2) return fail(ORC_ERR_INVALID_SPEC, "random slopes need a cluster (two slopes at most)");
        for (int q = 0; q < s->n_slopes; q++)
            if (s->slope_col[q] < 1 || s->slope_col[q] >= 1 + s->n_non_factor + s->n_factor_dummies + s->n_interactions)
                return fail(ORC_ERR_INVALID_SPEC, "cluster_slope_design_cols entry out of range");
        if (!(s->tau_squared > 0.0)) return fail(ORC_ERR_INVALID_SPEC, "random slopes need tau_squared > 0"); /* validate.rs invariant */
    }
    if (s->n_extra > 0) { /* validate.rs:720-745 */
        if (!s->has_cluster || s->n_extra > 2) return fail(ORC_ERR_INVALID_SPEC, "extra groupings need a cluster (two at most)");
        int n_nested = 0;
        for (int g = 0; g < s->n_extra; g++) {
            if (s->extra_count[g] < 1) return fail(ORC_ERR_INVALID_SPEC, "extra grouping needs at least one level");
            if (!(s->extra_tau_squared[g] >= 0.0)) return fail(ORC_ERR_INVALID_SPEC, "extra grouping tau_squared must be >= 0");
            if (!s->extra_nested[g] && s->sizing_fixed_size) return fail(ORC_ERR_INVALID_SPEC, "GroupingRelation::Crossed requires ClusterSizing::FixedClusters");
            if (s->extra_nested[g]) {
                n_nested++;
                if (s->sizing_fixed_size && s->sizing_value % s->extra_count[g] != 0)
                    return fail(ORC_ERR_INVALID_SPEC, "NestedWithin requires cluster_size % n_per_parent == 0");
            }
        }
        if (n_nested > 1) return fail(ORC_ERR_INVALID_SPEC, "at most one NestedWithin grouping is supported");
    }
    for (int b = 0; b < s->n_between; b++) /* the reference indexes x_full unchecked (a malformed spec panics); the backends reject it */
        if (s->between_cols[b] < 1 || s->between_cols[b] >= 1 + s->n_non_factor + s->n_factor_dummies)
            return fail(ORC_ERR_INVALID_SPEC, "between_var_indices entry out of range");
    /* uploaded-data tables: the reference indexes them unchecked (a malformed spec panics); the backends reject it */
    if (s->n_frame_map > 0) {
        if (s->n_frame_map != k + s->n_factors) return fail(ORC_ERR_INVALID_SPEC, "bootstrap_frame_map length must equal n_non_factor + n_factors");
        if (s->ud_rows <= 0 || s->ud_cols <= 0) return fail(ORC_ERR_INVALID_SPEC, "bootstrap mode requires a non-empty frame");
        for (int c = 0; c < s->n_frame_map; c++)
            if (s->frame_map[c] >= s->ud_cols) return fail(ORC_ERR_INVALID_SPEC, "bootstrap_frame_map entry out of range");
    }
    for (int j = 0; j < k; j++)
        if (s->var_types[j] == ORC_DIST_UPLOADED_DATA && s->un_rows > 0 && j >= s->un_cols)
            return fail(ORC_ERR_INVALID_SPEC, "upload_normal has no column for an uploaded_data variable");
    if (k > 0) {
        double* tmp = (double*)malloc(sizeof(double) * (size_t)k * k);
        memcpy(tmp, s->correlation, sizeof(double) * (size_t)k * k);
        int rc = llt_colmajor(tmp, k);
        free(tmp);
        if (rc) return fail(ORC_ERR_NOT_PSD, "correlation matrix is not positive semi-definite");
    }
    return ORC_OK;
}

int orc_generate_sim(const OrcSpec* s, uint64_t base_seed, uint64_t sim_id, int max_n, float* x, float* y, int32_t* cluster_ids) {
    int rc = validate_spec(s, NULL, 0);
    if (rc) return rc;
    Ws* w = ws_new(s, max_n);
    rc = gen_sim(s, base_seed, sim_id, w);
    if (!rc) {
        memcpy(x, w->x, sizeof(float) * (size_t)max_n * w->P);
        memcpy(y, w->y, sizeof(float) * (size_t)max_n);
        if (cluster_ids) {
            for (int i = 0; i < max_n; i++) cluster_ids[i] = w->cluster_ids ? w->cluster_ids[i] : 0;
        }
    }
    ws_free(w);
    return rc;
}

/* ======================================================================= */
/* OLS: sufficient statistics + normal-equation solve                       */
/* (glmm OlsSuffStats::add_rows / fit_suff_stats_t_sq as documented at       */
/*  workspace.rs:113-171, posthoc.rs:3-8, batch.rs:1015-1140 — restated)     */
/* ======================================================================= */

#define TRI(i, j) ((size_t)(i) * ((size_t)(i) + 1) / 2 + (size_t)(j))

typedef struct {
    double yty, sum_y;
    uint32_t n_rows;
} SuffScalars;

/* rows [from, to): sequential in row order, one fma per term (order fixed by this oracle). */
static void suff_add_rows(const float* x, const float* y, int ld, int P, int from, int to, double* xtx, double* xty, SuffScalars* sc) {
    double xr[64];
    for (int i = from; i < to; i++) {
        for (int j = 0; j < P; j++) xr[j] = (double)x[(size_t)j * ld + i];
        double yi = (double)y[i];
        size_t t = 0;
        for (int a = 0; a < P; a++) {
            for (int b = 0; b <= a; b++) { xtx[t] = fma(xr[a], xr[b], xtx[t]); t++; }
            xty[a] = fma(xr[a], yi, xty[a]);
        }
        sc->yty = fma(yi, yi, sc->yty);
        sc->sum_y += yi;
        sc->n_rows++;
    }
}

typedef struct {
    int converged;
    double sigma_sq, sst, rss, df_resid;
} OlsFit;

/* Solve on packed lower A (p×p), b; L (packed), inv_d, w, beta are caller scratch of size p(p+1)/2 / p. */
static void ols_solve(const double* A, const double* b, double yty, double sum_y, uint32_t n_rows, int p, double* L,
                      double* inv_d, double* wv, double* beta, OlsFit* fit) {
    int ok = 1;
    double dmin = INFINITY, dmax = 0.0;
    for (int i = 0; i < p; i++) {
        for (int j = 0; j <= i; j++) {
            double s = A[TRI(i, j)];
            for (int k = 0; k < j; k++) s = fma(-L[TRI(i, k)], L[TRI(j, k)], s);
            if (i == j) {
                if (!(s > 0.0) || !isfinite(s)) { ok = 0; s = 1.0; }
                double d = sqrt(s);
                L[TRI(i, i)] = d;
                inv_d[i] = 1.0 / d;
                if (d < dmin) dmin = d;
                if (d > dmax) dmax = d;
            } else {
                L[TRI(i, j)] = s * inv_d[j];
            }
        }
    }
    if (ok && !(dmin / dmax >= EPS_RANK)) ok = 0; /* batch.rs:40-41 */
    for (int i = 0; i < p; i++) {
        double s = b[i];
        for (int k = 0; k < i; k++) s = fma(-L[TRI(i, k)], wv[k], s);
        wv[i] = s * inv_d[i];
    }
    for (int i = p - 1; i >= 0; i--) {
        double s = wv[i];
        for (int k = i + 1; k < p; k++) s = fma(-L[TRI(k, i)], beta[k], s);
        beta[i] = s * inv_d[i];
    }
    double rss = yty;
    for (int i = 0; i < p; i++) rss = fma(-wv[i], wv[i], rss);
    double n = (double)n_rows;
    fit->converged = ok;
    fit->rss = rss;
    fit->sst = yty - sum_y * sum_y / n; /* workspace.rs:157-159 */
    fit->df_resid = n - (double)p;
    fit->sigma_sq = rss / fit->df_resid;
}

/* ‖L⁻¹c‖² for c = e_pos − e_neg (neg<0 → c = e_pos); posthoc.rs:124-148 */
static double linv_norm_sq(const double* L, const double* inv_d, int p, int pos, int neg, double* u) {
    int start = pos;
    if (neg >= 0 && neg < start) start = neg;
    double q = 0.0;
    for (int i = start; i < p; i++) {
        double s = (i == pos ? 1.0 : 0.0) - (i == neg ? 1.0 : 0.0);
        for (int k = start; k < i; k++) s = fma(-L[TRI(i, k)], u[k], s);
        u[i] = s * inv_d[i];
        q = fma(u[i], u[i], q);
    }
    return q;
}

static double target_t_sq(const double* L, const double* inv_d, const double* beta, double sigma_sq, int p, int pos, int neg, double* u) {
    double q = linv_norm_sq(L, inv_d, p, pos, neg, u);
    double est = neg >= 0 ? beta[pos] - beta[neg] : beta[pos];
    double var = sigma_sq * q;
    if (!(var > FLOAT_NEAR_ZERO) || !isfinite(var)) return NAN;
    return est * est / var;
}

int orc_fit_ols(const double* xc, const double* yv, int n, int p, double* betas, double* se, double* t_sq, double* sigma_sq,
                double* sst, double* rss, int* converged) {
    if (p > 64) return fail(ORC_ERR_INVALID_SPEC, "p > 64");
    /* introspect.rs:195-203: narrow supplied f64 to f32 first */
    float* x = (float*)malloc(sizeof(float) * (size_t)n * p);
    float* y = (float*)malloc(sizeof(float) * (size_t)n);
    for (size_t i = 0; i < (size_t)n * p; i++) x[i] = (float)xc[i];
    for (int i = 0; i < n; i++) y[i] = (float)yv[i];
    size_t nt = (size_t)p * (p + 1) / 2;
    double* xtx = (double*)calloc(nt * 2 + (size_t)p * 5, sizeof(double));
    double *L = xtx + nt, *xty = L + nt, *inv_d = xty + p, *wv = inv_d + p, *u = wv + p;
    SuffScalars sc = {0.0, 0.0, 0};
    suff_add_rows(x, y, n, p, 0, n, xtx, xty, &sc);
    OlsFit fit;
    ols_solve(xtx, xty, sc.yty, sc.sum_y, sc.n_rows, p, L, inv_d, wv, betas, &fit);
    for (int t = 0; t < p; t++) {
        double q = linv_norm_sq(L, inv_d, p, t, -1, u);
        double var = fit.sigma_sq * q;
        se[t] = sqrt(var);
        t_sq[t] = betas[t] * betas[t] / var;
    }
    *sigma_sq = fit.sigma_sq; *sst = fit.sst; *rss = fit.rss; *converged = fit.converged;
    free(xtx); free(x); free(y);
    return ORC_OK;
}

/* ======================================================================= */
/* GLM IRLS (glmm glm_irls_fit, call site batch.rs:1969-1981).  The crate's  */
/* constants are NOT visible in the reference tree ("parity      */
/* unpinned"); what IS visible is restated: truth start, no step-halving,    */
/* one LLT per iteration, W clamp, |beta| > 30 => not converged              */
/* (batch.rs:1942-1944,4550-4555, workspace.rs:185-192).  Stopping rule and  */
/* iteration cap follow R's glm.fit (|ddev|/(|dev|+0.1) < 1e-8, 25 its);     */
/* value-pinned to the 16 frozen R glm goldens at 1e-4 (tolerances.R:29-31). */
/* The CUDA kernel mirrors this function operation for operation.            */
/* ======================================================================= */

#define GLM_MAXIT 25
#define GLM_TOL 1e-8
#define GLM_WCLAMP 1e-10
#define GLM_BETA_CAP 30.0

typedef struct {
    int converged, iters;
    double deviance, deviance_null;
} GlmOut;

/* family 0 logit, 1 probit, 2 poisson(log) */
static inline void glm_row(int family, double eta, double y, double* w_o, double* z_o, double* dev_o) {
    double mu, w, z, dev;
    if (family == 0) {
        double e = orc_exp(-fabs(eta));
        double inv = 1.0 / (1.0 + e);
        mu = eta >= 0.0 ? inv : e * inv;
        w = mu * (1.0 - mu);
        if (w < GLM_WCLAMP) w = GLM_WCLAMP;
        z = eta + (y - mu) / w;
        dev = y > 0.5 ? -2.0 * orc_log(mu) : -2.0 * orc_log(1.0 - mu);
    } else if (family == 1) {
        mu = orc_phi(eta);
        double d = orc_exp(-0.5 * eta * eta) * 0.3989422804014327;
        if (d < 1e-300) d = 1e-300;
        double v = mu * (1.0 - mu);
        if (v < GLM_WCLAMP) v = GLM_WCLAMP;
        w = d * d / v;
        z = eta + (y - mu) / d;
        dev = y > 0.5 ? -2.0 * orc_log(mu) : -2.0 * orc_log(1.0 - mu);
    } else {
        mu = orc_exp(eta);
        w = mu;
        if (w < GLM_WCLAMP) w = GLM_WCLAMP;
        z = eta + (y - mu) / w;
        dev = y > 0.0 ? 2.0 * (y * orc_log(y / mu) - (y - mu)) : 2.0 * mu;
    }
    *w_o = w; *z_o = z; *dev_o = dev;
}

/* x: f32 column-major with leading dimension ld; kept = NULL (identity) or p column indices. */
static void glm_irls(const float* x, const float* y, int ld, int n, int p, const int* kept, int family,
                     const double* beta_start, double* A, double* b, double* L, double* inv_d, double* wv, double* beta,
                     GlmOut* out) {
    size_t nt = (size_t)p * (p + 1) / 2;
    double beta_tmp[64], xr[64];
    for (int j = 0; j < p; j++) beta[j] = beta_start ? beta_start[j] : 0.0;
    double dev_prev = INFINITY, dev = INFINITY, s0 = 0.0, s1 = 0.0;
    int conv = 0, it;
    OlsFit f;
    for (it = 0; it < GLM_MAXIT; it++) {
        memset(A, 0, sizeof(double) * nt);
        memset(b, 0, sizeof(double) * (size_t)p);
        dev = 0.0;
        for (int i = 0; i < n; i++) {
            double eta = 0.0;
            for (int j = 0; j < p; j++) {
                xr[j] = (double)x[(size_t)(kept ? kept[j] : j) * ld + i];
                eta = fma(xr[j], beta[j], eta);
            }
            double yi = (double)y[i], w, z, di;
            glm_row(family, eta, yi, &w, &z, &di);
            dev += di;
            if (it == 0) {
                s0 += yi;
                if (family == 2 && yi > 1.0) s1 += yi * orc_log(yi);
            }
            size_t t = 0;
            for (int a = 0; a < p; a++) {
                double wa = w * xr[a];
                for (int c = 0; c <= a; c++) { A[t] = fma(wa, xr[c], A[t]); t++; }
                b[a] = fma(wa, z, b[a]);
            }
        }
        if (!isfinite(dev)) break;
        if (it > 0 && fabs(dev - dev_prev) / (fabs(dev) + 0.1) < GLM_TOL) {
            /* converged at the current beta: factor X'WX(beta) for the Wald covariance, keep beta */
            ols_solve(A, b, 0.0, 0.0, (uint32_t)n, p, L, inv_d, wv, beta_tmp, &f);
            conv = f.converged;
            break;
        }
        dev_prev = dev;
        ols_solve(A, b, 0.0, 0.0, (uint32_t)n, p, L, inv_d, wv, beta_tmp, &f);
        if (!f.converged) break;
        int bad = 0;
        for (int j = 0; j < p; j++) {
            beta[j] = beta_tmp[j];
            if (!isfinite(beta[j]) || fabs(beta[j]) > GLM_BETA_CAP) bad = 1;
        }
        if (bad) break;
    }
    double dn, nn = (double)n, ybar = s0 / nn;
    if (family == 2) {
        dn = ybar > 0.0 ? 2.0 * (s1 - s0 * orc_log(ybar)) : 0.0;
    } else {
        double n1 = s0, n0 = nn - s0;
        dn = 0.0;
        if (n1 > 0.0) dn += n1 * orc_log(ybar);
        if (n0 > 0.0) dn += n0 * orc_log(1.0 - ybar);
        dn = -2.0 * dn;
    }
    out->converged = conv;
    out->iters = it;
    out->deviance = dev;
    out->deviance_null = dn;
}

int orc_fit_glm(const double* xc, const double* yv, int n, int p, int family, const double* beta_start, double* betas,
                double* se, double* deviance, double* deviance_null, int* converged, int* iters) {
    if (p > 64) return fail(ORC_ERR_INVALID_SPEC, "p > 64");
    float* x = (float*)malloc(sizeof(float) * (size_t)n * p);
    float* y = (float*)malloc(sizeof(float) * (size_t)n);
    for (size_t i = 0; i < (size_t)n * p; i++) x[i] = (float)xc[i]; /* introspect.rs:195-203 */
    for (int i = 0; i < n; i++) y[i] = (float)yv[i];
    size_t nt = (size_t)p * (p + 1) / 2;
    double* A = (double*)calloc(nt * 2 + (size_t)p * 5, sizeof(double));
    double *L = A + nt, *b = L + nt, *inv_d = b + p, *wv = inv_d + p, *u = wv + p, *beta = u + p;
    GlmOut o;
    glm_irls(x, y, n, n, p, NULL, family, beta_start, A, b, L, inv_d, wv, beta, &o);
    for (int t = 0; t < p; t++) {
        betas[t] = beta[t];
        se[t] = sqrt(linv_norm_sq(L, inv_d, p, t, -1, u));
    }
    *deviance = o.deviance; *deviance_null = o.deviance_null; *converged = o.converged; *iters = o.iters;
    free(A); free(x); free(y);
    return ORC_OK;
}

/* ======================================================================= */
/* Random-intercept LMM by profiled REML (MLE arm, batch.rs:1214-1561).      */
/* The reference fits through glmm::fit_on + BOBYQA over theta (not in the   */
/* tree).  For one random intercept the profiled REML deviance is a closed   */
/* form in theta = tau/sigma over cluster-level sufficient statistics        */
/* (scripts/lme_handcompute_theta1.py:13-36); this restatement minimises it  */
/* with Brent's fmin on [0, THETA_MAX].  Value-pinned to the 8 frozen lmer   */
/* goldens (1e-4 gate, tolerances.R:29-31); optimiser path UNPINNED.         */
/* ======================================================================= */

#define LMM_THETA_MAX 30.0
#define LMM_TOL 1e-10
#define LMM_MAXIT 200

typedef struct {
    int p, ng;
    uint32_t n_rows;
    const double *xtx, *xty;
    double yty;
    double grp_n[2], grp_m[2];      /* cluster size / number of clusters of that size */
    double M[2][64 * 65 / 2], v[2][64], w[2];
    /* state of the most recent evaluation */
    double A[64 * 65 / 2], b[64], L[64 * 65 / 2], inv_d[64], wv[64], beta[64];
    OlsFit fit;
} LmmCtx;

static double lmm_dev(LmmCtx* c, double theta) {
    const int p = c->p;
    const size_t nt = (size_t)p * (p + 1) / 2;
    double t2 = theta * theta, s[2] = {0.0, 0.0};
    for (int g = 0; g < c->ng; g++) s[g] = t2 / (1.0 + t2 * c->grp_n[g]);
    for (size_t e = 0; e < nt; e++) {
        double a = c->xtx[e];
        for (int g = 0; g < c->ng; g++) a = fma(-s[g], c->M[g][e], a);
        c->A[e] = a;
    }
    for (int j = 0; j < p; j++) {
        double a = c->xty[j];
        for (int g = 0; g < c->ng; g++) a = fma(-s[g], c->v[g][j], a);
        c->b[j] = a;
    }
    double cc = c->yty;
    for (int g = 0; g < c->ng; g++) cc = fma(-s[g], c->w[g], cc);
    ols_solve(c->A, c->b, cc, 0.0, c->n_rows, p, c->L, c->inv_d, c->wv, c->beta, &c->fit);
    if (!c->fit.converged || !(c->fit.rss > 0.0) || !isfinite(c->fit.rss)) return INFINITY;
    double ldv = 0.0;
    for (int g = 0; g < c->ng; g++) ldv += c->grp_m[g] * orc_log(1.0 + t2 * c->grp_n[g]);
    double lda = 0.0;
    for (int j = 0; j < p; j++) lda += orc_log(c->L[TRI(j, j)]);
    return ldv + 2.0 * lda + c->fit.df_resid * orc_log(c->fit.sigma_sq);
}

/* Brent's fmin (golden section + parabolic interpolation) on [a, b]; deterministic, restated in the CUDA kernel. */
static double lmm_fmin(LmmCtx* c, double a, double b, double tol, double* fmin_out, int* iters) {
    const double cgold = 0.3819660112501051, eps = 1.4901161193847656e-08;
    double v = a + cgold * (b - a), w = v, x = v, e = 0.0, d = 0.0;
    double fx = lmm_dev(c, x), fv = fx, fw = fx;
    int it;
    for (it = 0; it < LMM_MAXIT; it++) {
        double xm = 0.5 * (a + b), tol1 = eps * fabs(x) + tol / 3.0, tol2 = 2.0 * tol1;
        if (fabs(x - xm) <= tol2 - 0.5 * (b - a)) break;
        int golden = 1;
        if (fabs(e) > tol1) {
            double r = (x - w) * (fx - fv), q = (x - v) * (fx - fw), pp = (x - v) * q - (x - w) * r;
            q = 2.0 * (q - r);
            if (q > 0.0) pp = -pp; else q = -q;
            r = e;
            e = d;
            if (fabs(pp) < fabs(0.5 * q * r) && pp > q * (a - x) && pp < q * (b - x)) {
                d = pp / q;
                double u = x + d;
                if (u - a < tol2 || b - u < tol2) d = xm - x >= 0.0 ? tol1 : -tol1;
                golden = 0;
            }
        }
        if (golden) {
            e = x >= xm ? a - x : b - x;
            d = cgold * e;
        }
        double u = fabs(d) >= tol1 ? x + d : (d >= 0.0 ? x + tol1 : x - tol1);
        double fu = lmm_dev(c, u);
        if (fu <= fx) {
            if (u >= x) a = x; else b = x;
            v = w; fv = fw; w = x; fw = fx; x = u; fx = fu;
        } else {
            if (u < x) a = u; else b = u;
            if (fu <= fw || w == x) { v = w; fv = fw; w = u; fw = fu; }
            else if (fu <= fv || v == x || v == w) { v = u; fv = fu; }
        }
    }
    *fmin_out = fx;
    *iters = it;
    return x;
}

typedef struct {
    int converged, boundary_hit;
    double theta, sigma_sq, tau_sq;
} LmmOut;

/* csum: per-cluster [n_clusters][p+1] = (sum of x_j over the cluster's rows ..., sum of y); cn[c] = rows in cluster c */
static void lmm_fit(LmmCtx* c, const double* csum, const uint32_t* cn, int n_clusters, int p, LmmOut* out) {
    const size_t nt = (size_t)p * (p + 1) / 2;
    c->ng = 0;
    for (int g = 0; g < 2; g++) {
        c->grp_n[g] = 0.0; c->grp_m[g] = 0.0; c->w[g] = 0.0;
        memset(c->M[g], 0, sizeof(double) * nt);
        memset(c->v[g], 0, sizeof(double) * (size_t)p);
    }
    for (int k = 0; k < n_clusters; k++) {
        if (cn[k] == 0) continue;
        int g;
        if (c->ng == 0) { g = 0; c->ng = 1; c->grp_n[0] = (double)cn[k]; }
        else if ((double)cn[k] == c->grp_n[0]) g = 0;
        else { g = 1; if (c->ng == 1) { c->ng = 2; c->grp_n[1] = (double)cn[k]; } }
        c->grp_m[g] += 1.0;
        const double* sx = csum + (size_t)k * (p + 1);
        size_t t = 0;
        for (int a = 0; a < p; a++) {
            for (int bb = 0; bb <= a; bb++) { c->M[g][t] = fma(sx[a], sx[bb], c->M[g][t]); t++; }
            c->v[g][a] = fma(sx[a], sx[p], c->v[g][a]);
        }
        c->w[g] = fma(sx[p], sx[p], c->w[g]);
    }
    double fbest;
    int iters;
    double th = lmm_fmin(c, 0.0, LMM_THETA_MAX, LMM_TOL, &fbest, &iters);
    double f0 = lmm_dev(c, 0.0);
    out->boundary_hit = 0;
    out->converged = 1;
    if (!(fbest < f0)) { th = 0.0; out->boundary_hit = 1 | (1 << 2); } /* tau-hat = 0: OLS fallback (spec.rs:562-565); component 0 pinned */
    else if (LMM_THETA_MAX - th < 1e-6 * LMM_THETA_MAX) { out->boundary_hit = 2; out->converged = 0; }
    double f = lmm_dev(c, th);                                    /* leaves A, L, beta, sigma at theta-hat */
    if (!isfinite(f)) out->converged = 0;
    out->theta = th;
    out->sigma_sq = c->fit.sigma_sq;
    out->tau_sq = th * th * c->fit.sigma_sq;
}

/* ---- random intercept + ONE random slope (q = 2): profiled REML over theta = (l00, l10, l11), Lambda = [[l00, 0], [l10, l11]] ----
 * The reference fits these through glmm::fit_on + BOBYQA (batch.rs:1214-1561, mixed_workspace.rs:203-272; crates not in the
 * tree).  Restated in the lme4 form over per-cluster sufficient statistics: with Z_c = [1, x_s],
 *   M_c = Lambda' Z_c'Z_c Lambda + I,  L_c = chol(M_c),  R_c = L_c^-1 Lambda' Z_c'[X | y]  (2 x (p + 1)),
 *   A = X'X - sum_c R_c[:, :p]' R_c[:, :p],  b = X'y - sum_c R_c[:, :p]' R_c[:, p],  c = y'y - sum_c |R_c[:, p]|^2,
 *   dev(theta) = sum_c log det M_c + log det A + (n - p) log(rss / (n - p)).
 * Per cluster the statistics are the sums of [x | y] (cs1) and of x_s * [x | y] (cs2): Z'Z = [[n_c, cs1[s]], [cs1[s], cs2[s]]].
 * Minimised by a deterministic Nelder-Mead (one restart) with the diagonal of Lambda clamped at 0; restated op for op in the
 * CUDA kernel.  Pinned to lme4's REML fits of the reference's lmm_slope datasets (tolerances.R:29-31, 38-49). */
#define SLOPE_MAXQ 3
#define SLOPE_MAXD 8 /* q (q + 1) / 2 for the primary grouping + one scale per extra grouping */
#define SLOPE_NM_EVAL_PER_DIM 300
#define SLOPE_THETA_MAX 30.0
typedef struct {
    const double* cs[SLOPE_MAXQ]; /* cs[0]: per-cluster sums of [x | y]; cs[a]: of x_{s_a} * [x | y]; each [n_clusters][p + 1] */
    const uint32_t* cn;
    int n_clusters, q;
    int slope_col[SLOPE_MAXQ]; /* slope_col[a] = design column of slope a (a >= 1) */
    int n_eval;
} SlopeData;

/* theta = the lower triangle of Lambda row by row: (l00, l10, l11[, l20, l21, l22]); diagonal entries clamped at 0 */
static double slope_dev(LmmCtx* c, const SlopeData* d, const double* th) {
    const int p = c->p, q = d->q;
    const size_t nt = (size_t)p * (p + 1) / 2;
    double Lm[SLOPE_MAXQ][SLOPE_MAXQ];
    for (int i = 0, t = 0; i < q; i++)
        for (int j = 0; j <= i; j++, t++) Lm[i][j] = i == j ? fabs(th[t]) : th[t];
    for (size_t e = 0; e < nt; e++) c->A[e] = c->xtx[e];
    for (int j = 0; j < p; j++) c->b[j] = c->xty[j];
    double cc = c->yty, ldv = 0.0;
    double r[SLOPE_MAXQ][65];
    for (int k = 0; k < d->n_clusters; k++) {
        if (d->cn[k] == 0) continue;
        const double* sk[SLOPE_MAXQ];
        for (int a = 0; a < q; a++) sk[a] = d->cs[a] + (size_t)k * (p + 1);
        /* S = Z'Z (symmetric q x q): S[0][0] = n_k, S[a][0] = sum x_a, S[a][b] = sum x_a x_b */
        double S[SLOPE_MAXQ][SLOPE_MAXQ], T[SLOPE_MAXQ][SLOPE_MAXQ], M[SLOPE_MAXQ][SLOPE_MAXQ], C[SLOPE_MAXQ][SLOPE_MAXQ];
        S[0][0] = (double)d->cn[k];
        for (int a = 1; a < q; a++) {
            S[a][0] = S[0][a] = sk[0][d->slope_col[a]];
            for (int b = 1; b <= a; b++) S[a][b] = S[b][a] = sk[b][d->slope_col[a]];
        }
        /* T = S Lambda; M = Lambda' S Lambda + I (lower part: M[i][j] = sum_k T[k][i] Lambda[k][j]) */
        for (int i = 0; i < q; i++)
            for (int j = 0; j < q; j++) {
                double acc = 0.0;
                for (int m = j; m < q; m++) acc += Lm[m][j] * S[i][m];
                T[i][j] = acc;
            }
        for (int i = 0; i < q; i++)
            for (int j = 0; j <= i; j++) {
                double acc = 0.0;
                for (int m = j; m < q; m++) acc += T[m][i] * Lm[m][j];
                M[i][j] = i == j ? acc + 1.0 : acc;
            }
        for (int j = 0; j < q; j++) { /* C = chol(M) */
            double dg = M[j][j];
            for (int m = 0; m < j; m++) dg -= C[j][m] * C[j][m];
            C[j][j] = sqrt(dg);
            for (int i = j + 1; i < q; i++) {
                double acc = M[i][j];
                for (int m = 0; m < j; m++) acc -= C[i][m] * C[j][m];
                C[i][j] = acc / C[j][j];
            }
        }
        { /* log det, summed pairwise from the front like the q = 2 form: log c00 + log c11 (+ log c22) */
            double l = orc_log(C[0][0]);
            for (int j = 1; j < q; j++) l += orc_log(C[j][j]);
            ldv += l;
        }
        for (int j = 0; j <= p; j++) { /* R = C^-1 Lambda' Z'[X | y] */
            for (int i = 0; i < q; i++) {
                double t = 0.0;
                for (int m = i; m < q; m++) t += Lm[m][i] * sk[m][j];
                for (int m = 0; m < i; m++) t -= C[i][m] * r[m][j];
                r[i][j] = t / C[i][i];
            }
        }
        size_t t = 0;
        for (int a = 0; a < p; a++) {
            for (int bb = 0; bb <= a; bb++) {
                for (int i = 0; i < q; i++) c->A[t] = fma(-r[i][a], r[i][bb], c->A[t]);
                t++;
            }
            for (int i = 0; i < q; i++) c->b[a] = fma(-r[i][a], r[i][p], c->b[a]);
        }
        for (int i = 0; i < q; i++) cc = fma(-r[i][p], r[i][p], cc);
    }
    ((SlopeData*)d)->n_eval++;
    ols_solve(c->A, c->b, cc, 0.0, c->n_rows, p, c->L, c->inv_d, c->wv, c->beta, &c->fit);
    if (!c->fit.converged || !(c->fit.rss > 0.0) || !isfinite(c->fit.rss) || !isfinite(ldv)) return INFINITY;
    double lda = 0.0;
    for (int j = 0; j < p; j++) lda += orc_log(c->L[TRI(j, j)]);
    return 2.0 * ldv + 2.0 * lda + c->fit.df_resid * orc_log(c->fit.sigma_sq);
}

/* Nelder-Mead in D = q (q + 1) / 2 dimensions (reflection 1, expansion 2, contraction 1/2, shrink 1/2); x0 + step initial
 * simplex; stops when the simplex is flat (1e-11 relative in f) and small (1e-7 in every coordinate) or after max_eval
 * evaluations.  Ties and orderings are resolved by strict comparisons in a fixed order, so the CUDA twin takes the same path. */
typedef double (*NmEval)(LmmCtx*, void*, const double*);
static double nm_minimise(LmmCtx* c, void* d, NmEval ev, int* n_eval, int D, double* x, double step, int max_eval);
static double slope_eval(LmmCtx* c, void* d, const double* th) { return slope_dev(c, (const SlopeData*)d, th); }
static double slope_nm(LmmCtx* c, SlopeData* d, double* x, double step, int max_eval) {
    return nm_minimise(c, d, slope_eval, &d->n_eval, d->q * (d->q + 1) / 2, x, step, max_eval);
}
static double nm_minimise(LmmCtx* c, void* d, NmEval slope_dev, int* n_eval_p, int D, double* x, double step, int max_eval) {
    const int V = D + 1;
    double sx[SLOPE_MAXD + 1][SLOPE_MAXD], sf[SLOPE_MAXD + 1];
    for (int i = 0; i < V; i++) {
        for (int j = 0; j < D; j++) sx[i][j] = x[j] + (i == j + 1 ? step : 0.0);
        sf[i] = slope_dev(c, d, sx[i]);
    }
    for (;;) {
        /* order: best (lo), worst (hi), second worst (nh) */
        int lo = 0, hi = 0, nh;
        for (int i = 1; i < V; i++) {
            if (sf[i] < sf[lo]) lo = i;
            if (sf[i] > sf[hi] || !(sf[i] == sf[i])) hi = i;
        }
        nh = lo;
        for (int i = 0; i < V; i++)
            if (i != hi && (sf[i] > sf[nh])) nh = i;
        double size = 0.0;
        for (int i = 0; i < V; i++)
            for (int j = 0; j < D; j++) {
                double dd = fabs(sx[i][j] - sx[lo][j]);
                if (dd > size) size = dd;
            }
        if ((fabs(sf[hi] - sf[lo]) <= 1e-11 * (fabs(sf[lo]) + 1.0) && size <= 1e-7) || *n_eval_p >= max_eval) {
            for (int j = 0; j < D; j++) x[j] = sx[lo][j];
            return sf[lo];
        }
        double cen[SLOPE_MAXD], xr[SLOPE_MAXD], xe[SLOPE_MAXD], xc[SLOPE_MAXD];
        for (int j = 0; j < D; j++) {
            double sum = 0.0;
            for (int i = 0; i < V; i++)
                if (i != hi) sum += sx[i][j];
            cen[j] = sum / (double)D;
            xr[j] = cen[j] + (cen[j] - sx[hi][j]);
        }
        double fr = slope_dev(c, d, xr);
        if (fr < sf[lo]) {
            for (int j = 0; j < D; j++) xe[j] = cen[j] + 2.0 * (cen[j] - sx[hi][j]);
            double fe = slope_dev(c, d, xe);
            if (fe < fr) { for (int j = 0; j < D; j++) sx[hi][j] = xe[j]; sf[hi] = fe; }
            else { for (int j = 0; j < D; j++) sx[hi][j] = xr[j]; sf[hi] = fr; }
        } else if (fr < sf[nh]) {
            for (int j = 0; j < D; j++) sx[hi][j] = xr[j];
            sf[hi] = fr;
        } else {
            const int outside = fr < sf[hi];
            for (int j = 0; j < D; j++) xc[j] = outside ? cen[j] + 0.5 * (xr[j] - cen[j]) : cen[j] + 0.5 * (sx[hi][j] - cen[j]);
            double fc = slope_dev(c, d, xc);
            if (fc < (outside ? fr : sf[hi])) {
                for (int j = 0; j < D; j++) sx[hi][j] = xc[j];
                sf[hi] = fc;
            } else {
                for (int i = 0; i < V; i++) {
                    if (i == lo) continue;
                    for (int j = 0; j < D; j++) sx[i][j] = sx[lo][j] + 0.5 * (sx[i][j] - sx[lo][j]);
                    sf[i] = slope_dev(c, d, sx[i]);
                }
            }
        }
    }
}

typedef struct {
    int converged, boundary_hit, n_eval;
    double theta[SLOPE_MAXD], sigma_sq;
} SlopeOut;

static void lmm_slope_fit(LmmCtx* c, SlopeData* d, SlopeOut* out) {
    const int q = d->q, D = q * (q + 1) / 2, max_eval = SLOPE_NM_EVAL_PER_DIM * D;
    double x[SLOPE_MAXD];
    for (int i = 0, t = 0; i < q; i++)
        for (int j = 0; j <= i; j++, t++) x[t] = i == j ? 1.0 : 0.0;
    d->n_eval = 0;
    slope_nm(c, d, x, 0.5, max_eval / 2);
    double f = slope_nm(c, d, x, 0.05, max_eval); /* restart from the best vertex: guards against a collapsed simplex */
    out->boundary_hit = 0;
    for (int i = 0, t = 0; i < q; i++)
        for (int j = 0; j <= i; j++, t++)
            if (i == j) {
                x[t] = fabs(x[t]);
                if (x[t] < 1e-6) out->boundary_hit |= 1 | (4 << i); /* variance component i pinned at 0: code 1, pinned-mask bit i */
            }
    f = slope_dev(c, d, x); /* leaves A, L, beta, sigma at theta-hat */
    out->converged = isfinite(f) && d->n_eval < max_eval;
    for (int t = 0; t < D; t++)
        if (fabs(x[t]) > SLOPE_THETA_MAX) { out->boundary_hit = 2 | (out->boundary_hit & ~3); out->converged = 0; }
    out->n_eval = d->n_eval;
    for (int j = 0; j < D; j++) out->theta[j] = x[j];
    out->sigma_sq = c->fit.sigma_sq;
}

/* ---- general structure: primary grouping (intercept + up to two slopes) plus up to two intercept-only extra groupings (crossed /
 * nested).  The reference fits these through glmm::fit_on + BOBYQA (batch.rs:1214-1561, mixed_workspace.rs; crates not in the tree).
 * Restated in lme4's penalised-least-squares form on dense matrices: random-effect columns ordered [primary (level, component) |
 * extra 1 levels | extra 2 levels], Z'Z and Z'[X | y] accumulated once from the rows (one fma per term, rows in order), then per
 * evaluation  M = Lambda' Z'Z Lambda + I,  L = chol(M) (left-looking, one fma chain per entry),  R = L^-1 Lambda' Z'[X | y],
 *   A = X'X - R_x' R_x,  b = X'y - R_x' R_y,  c = y'y - |R_y|²,  dev = 2 sum log L_jj + log det A + (n - p) log(rss / (n - p)),
 * theta = (lower triangle of the primary's Lambda row by row, one scale per extra grouping), minimised by the same Nelder-Mead.
 * Structural zeros contribute exact zeros to every chain, so an implementation that skips them (the CUDA kernel keeps the primary
 * block diagonal and only the extras dense) produces the same bits. */
typedef struct {
    int Q, q0, K, n_extra, E[2], off[2]; /* off[g] = first column of extra grouping g */
    const double* ZtZ;  /* Q x Q, lower triangle used, row-major [i * Q + j] */
    const double* ZtXy; /* Q x (p + 1) */
    double *T, *M, *L, *R; /* work: Q x Q (T, M, L), Q x (p + 1) (R) */
    int n_eval;
} GenData;

static double gen_dev(LmmCtx* c, GenData* d, const double* th) {
    const int p = c->p, Q = d->Q, q0 = d->q0, nprim = d->K * q0;
    const size_t nt = (size_t)p * (p + 1) / 2;
    double Lm[SLOPE_MAXQ][SLOPE_MAXQ], te[2];
    int t = 0;
    for (int i = 0; i < q0; i++)
        for (int j = 0; j <= i; j++, t++) Lm[i][j] = i == j ? fabs(th[t]) : th[t];
    for (int g = 0; g < d->n_extra; g++) te[g] = fabs(th[t++]);
    double *T = d->T, *M = d->M, *L = d->L, *R = d->R;
    /* T = Z'Z Lambda (column j), M = Lambda' T + I (lower part) */
    for (int r = 0; r < Q; r++)
        for (int j = 0; j <= r; j++) {
            double acc = 0.0;
            if (j < nprim) {
                const int k = j / q0, aj = j % q0;
                for (int bb = aj; bb < q0; bb++) {
                    const int col = k * q0 + bb;
                    acc += Lm[bb][aj] * (r >= col ? d->ZtZ[(size_t)r * Q + col] : d->ZtZ[(size_t)col * Q + r]);
                }
            } else {
                acc += te[j >= d->off[1] && d->n_extra > 1 ? 1 : 0] * d->ZtZ[(size_t)r * Q + j];
            }
            T[(size_t)r * Q + j] = acc;
        }
    for (int i = 0; i < Q; i++)
        for (int j = 0; j <= i; j++) {
            double acc = 0.0;
            if (i < nprim) {
                const int k = i / q0, ai = i % q0;
                for (int aa = ai; aa < q0; aa++) acc += Lm[aa][ai] * T[(size_t)(k * q0 + aa) * Q + j];
            } else {
                acc += te[i >= d->off[1] && d->n_extra > 1 ? 1 : 0] * T[(size_t)i * Q + j];
            }
            M[(size_t)i * Q + j] = i == j ? acc + 1.0 : acc;
        }
    double ldv = 0.0;
    for (int j = 0; j < Q; j++) {
        double dg = M[(size_t)j * Q + j];
        for (int m = 0; m < j; m++) dg = fma(-L[(size_t)j * Q + m], L[(size_t)j * Q + m], dg);
        if (!(dg > 0.0)) { d->n_eval++; return INFINITY; }
        const double ljj = sqrt(dg);
        L[(size_t)j * Q + j] = ljj;
        ldv += orc_log(ljj);
        for (int i = j + 1; i < Q; i++) {
            double sacc = M[(size_t)i * Q + j];
            for (int m = 0; m < j; m++) sacc = fma(-L[(size_t)i * Q + m], L[(size_t)j * Q + m], sacc);
            L[(size_t)i * Q + j] = sacc / ljj;
        }
    }
    for (int cc = 0; cc <= p; cc++)
        for (int i = 0; i < Q; i++) {
            double u = 0.0;
            if (i < nprim) {
                const int k = i / q0, ai = i % q0;
                for (int aa = ai; aa < q0; aa++) u += Lm[aa][ai] * d->ZtXy[(size_t)(k * q0 + aa) * (p + 1) + cc];
            } else {
                u += te[i >= d->off[1] && d->n_extra > 1 ? 1 : 0] * d->ZtXy[(size_t)i * (p + 1) + cc];
            }
            for (int m = 0; m < i; m++) u = fma(-L[(size_t)i * Q + m], R[(size_t)m * (p + 1) + cc], u);
            R[(size_t)i * (p + 1) + cc] = u / L[(size_t)i * Q + i];
        }
    for (size_t e = 0; e < nt; e++) c->A[e] = c->xtx[e];
    for (int j = 0; j < p; j++) c->b[j] = c->xty[j];
    double cc = c->yty;
    for (int i = 0; i < Q; i++) {
        const double* ri = R + (size_t)i * (p + 1);
        size_t e = 0;
        for (int a = 0; a < p; a++) {
            for (int bb = 0; bb <= a; bb++, e++) c->A[e] = fma(-ri[a], ri[bb], c->A[e]);
            c->b[a] = fma(-ri[a], ri[p], c->b[a]);
        }
        cc = fma(-ri[p], ri[p], cc);
    }
    d->n_eval++;
    ols_solve(c->A, c->b, cc, 0.0, c->n_rows, p, c->L, c->inv_d, c->wv, c->beta, &c->fit);
    if (!c->fit.converged || !(c->fit.rss > 0.0) || !isfinite(c->fit.rss) || !isfinite(ldv)) return INFINITY;
    double lda = 0.0;
    for (int j = 0; j < p; j++) lda += orc_log(c->L[TRI(j, j)]);
    return 2.0 * ldv + 2.0 * lda + c->fit.df_resid * orc_log(c->fit.sigma_sq);
}

static double gen_eval(LmmCtx* c, void* d, const double* th) { return gen_dev(c, (GenData*)d, th); }

static void lmm_general_fit(LmmCtx* c, GenData* d, SlopeOut* out) {
    const int q0 = d->q0, D = q0 * (q0 + 1) / 2 + d->n_extra, max_eval = SLOPE_NM_EVAL_PER_DIM * D;
    double x[SLOPE_MAXD];
    int t = 0;
    for (int i = 0; i < q0; i++)
        for (int j = 0; j <= i; j++, t++) x[t] = i == j ? 1.0 : 0.0;
    for (int g = 0; g < d->n_extra; g++) x[t++] = 1.0;
    d->n_eval = 0;
    nm_minimise(c, d, gen_eval, &d->n_eval, D, x, 0.5, max_eval / 2);
    double f = nm_minimise(c, d, gen_eval, &d->n_eval, D, x, 0.05, max_eval);
    out->boundary_hit = 0;
    t = 0;
    for (int i = 0; i < q0; i++)
        for (int j = 0; j <= i; j++, t++)
            if (i == j) {
                x[t] = fabs(x[t]);
                if (x[t] < 1e-6) out->boundary_hit |= 1 | (4 << i);
            }
    for (int g = 0; g < d->n_extra; g++, t++) {
        x[t] = fabs(x[t]);
        if (x[t] < 1e-6) out->boundary_hit |= 1 | (4 << (q0 + g)); /* components of the fitted model: primary diagonal, extras */
    }
    f = gen_dev(c, d, x);
    out->converged = isfinite(f) && d->n_eval < max_eval;
    for (int j = 0; j < D; j++)
        if (fabs(x[j]) > SLOPE_THETA_MAX) { out->boundary_hit = 2 | (out->boundary_hit & ~3); out->converged = 0; }
    out->n_eval = d->n_eval;
    for (int j = 0; j < D; j++) out->theta[j] = x[j];
    out->sigma_sq = c->fit.sigma_sq;
}

/* one row into Z'Z (lower) and Z'[X | y]: the row's non-zero random-effect columns are (primary level, component a) with value
 * z_a = 1, x_s1, x_s2 and one column of value 1 per extra grouping, in column order */
static void gen_add_row(double* ZtZ, double* ZtXy, int Q, int p, const int* cols, const double* zv, int nz, const double* xy) {
    for (int u = 0; u < nz; u++) {
        for (int v = 0; v <= u; v++) {
            double* e = ZtZ + (size_t)cols[u] * Q + cols[v];
            *e = fma(zv[u], zv[v], *e);
        }
        double* r = ZtXy + (size_t)cols[u] * (p + 1);
        for (int cc = 0; cc <= p; cc++) r[cc] = fma(zv[u], xy[cc], r[cc]);
    }
}

/* Joint Wald chi-square of the targets: beta_T' [Cov_TT]^-1 beta_T with Cov = sigma^2 A^-1. */
static double lmm_joint_wald(const LmmCtx* c, const int* cols, int k) {
    if (k <= 0 || k > 32) return NAN;
    const int p = c->p;
    double U[32][64], G[32 * 33 / 2], Lg[32 * 33 / 2], invd[32], wv[32], u[64];
    for (int t = 0; t < k; t++) {
        linv_norm_sq(c->L, c->inv_d, p, cols[t], -1, u);
        for (int i = 0; i < p; i++) U[t][i] = i >= cols[t] ? u[i] : 0.0;
    }
    for (int a = 0; a < k; a++)
        for (int bb = 0; bb <= a; bb++) {
            double sacc = 0.0;
            for (int i = 0; i < p; i++) sacc = fma(U[a][i], U[bb][i], sacc);
            G[TRI(a, bb)] = sacc;
        }
    double bt[32], sol[32];
    for (int t = 0; t < k; t++) bt[t] = c->beta[cols[t]];
    OlsFit f;
    ols_solve(G, bt, 0.0, 0.0, 2u * (uint32_t)k + 2u, k, Lg, invd, wv, sol, &f);
    if (!f.converged) return NAN;
    double q = 0.0;
    for (int t = 0; t < k; t++) q = fma(wv[t], wv[t], q);   /* |Lg^-1 beta_T|^2 = beta_T' G^-1 beta_T */
    return q / c->fit.sigma_sq;
}

int orc_fit_lmm(const double* xc, const double* yv, const int32_t* ids, int n, int p, double* betas, double* se,
                double* theta, double* sigma_sq, int* converged, int* boundary) {
    if (p > 64) return fail(ORC_ERR_INVALID_SPEC, "p > 64");
    float* x = (float*)malloc(sizeof(float) * (size_t)n * p);
    float* y = (float*)malloc(sizeof(float) * (size_t)n);
    for (size_t i = 0; i < (size_t)n * p; i++) x[i] = (float)xc[i];
    for (int i = 0; i < n; i++) y[i] = (float)yv[i];
    int nc = 0;
    for (int i = 0; i < n; i++) if (ids[i] + 1 > nc) nc = ids[i] + 1;
    size_t nt = (size_t)p * (p + 1) / 2;
    double* xtx = (double*)calloc(nt + (size_t)p + (size_t)nc * (p + 1), sizeof(double));
    double *xty = xtx + nt, *csum = xty + p;
    uint32_t* cn = (uint32_t*)calloc((size_t)nc, sizeof(uint32_t));
    SuffScalars sc = {0.0, 0.0, 0};
    suff_add_rows(x, y, n, p, 0, n, xtx, xty, &sc);
    for (int i = 0; i < n; i++) {
        double* sx = csum + (size_t)ids[i] * (p + 1);
        for (int j = 0; j < p; j++) sx[j] += (double)x[(size_t)j * n + i];
        sx[p] += (double)y[i];
        cn[ids[i]]++;
    }
    LmmCtx* c = (LmmCtx*)calloc(1, sizeof(LmmCtx));
    c->p = p; c->n_rows = (uint32_t)n; c->xtx = xtx; c->xty = xty; c->yty = sc.yty;
    LmmOut o;
    lmm_fit(c, csum, cn, nc, p, &o);
    double u[64];
    for (int t = 0; t < p; t++) {
        betas[t] = c->beta[t];
        se[t] = sqrt(c->fit.sigma_sq * linv_norm_sq(c->L, c->inv_d, p, t, -1, u));
    }
    *theta = o.theta; *sigma_sq = o.sigma_sq; *converged = o.converged; *boundary = o.boundary_hit & 3;
    free(c); free(cn); free(xtx); free(x); free(y);
    return ORC_OK;
}

/* ======================================================================= */
/* Batch driver (batch.rs:199-776, OLS arm 871-1213)                        */
/* ======================================================================= */

/* batch.rs:2451-2496 */
/* variance components of Lambda-hat: Sigma = sigma² Lambda Lambda'; out = {Sigma_00 .. Sigma_{q-1,q-1}, corr(1,0)[, corr(2,0), corr(2,1)], sigma²} */
static void slope_components(const double* th, int q, double sigma_sq, double* out) {
    double Lm[SLOPE_MAXQ][SLOPE_MAXQ] = {{0}}, G[SLOPE_MAXQ][SLOPE_MAXQ];
    for (int i = 0, t = 0; i < q; i++)
        for (int j = 0; j <= i; j++, t++) Lm[i][j] = th[t];
    for (int i = 0; i < q; i++)
        for (int j = 0; j <= i; j++) {
            double acc = 0.0;
            for (int m = 0; m <= j; m++) acc += Lm[i][m] * Lm[j][m];
            G[i][j] = acc;
        }
    int o = 0;
    for (int i = 0; i < q; i++) out[o++] = sigma_sq * G[i][i];
    for (int i = 1; i < q; i++)
        for (int j = 0; j < i; j++) out[o++] = G[i][j] / sqrt(G[i][i] * G[j][j]);
    out[o] = sigma_sq;
}

int orc_fit_lmm_slopes(const double* xc, const double* yv, const int32_t* ids, int n, int p, int n_slopes, const int32_t* slope_cols,
                       double* betas, double* se, double* vc, int* converged, int* boundary, int* n_eval) {
    if (p > 64 || n_slopes < 1 || n_slopes > 2) return fail(ORC_ERR_INVALID_SPEC, "bad p / slope count");
    for (int a = 0; a < n_slopes; a++)
        if (slope_cols[a] < 0 || slope_cols[a] >= p) return fail(ORC_ERR_INVALID_SPEC, "bad slope column");
    const int q = 1 + n_slopes;
    float* x = (float*)malloc(sizeof(float) * (size_t)n * p);
    float* y = (float*)malloc(sizeof(float) * (size_t)n);
    for (size_t i = 0; i < (size_t)n * p; i++) x[i] = (float)xc[i];
    for (int i = 0; i < n; i++) y[i] = (float)yv[i];
    int nc = 0;
    for (int i = 0; i < n; i++) if (ids[i] + 1 > nc) nc = ids[i] + 1;
    size_t nt = (size_t)p * (p + 1) / 2, cl = (size_t)nc * (p + 1);
    double* xtx = (double*)calloc(nt + (size_t)p + cl * 3, sizeof(double));
    double *xty = xtx + nt, *cs0 = xty + p;
    uint32_t* cn = (uint32_t*)calloc((size_t)nc, sizeof(uint32_t));
    SuffScalars sc = {0.0, 0.0, 0};
    suff_add_rows(x, y, n, p, 0, n, xtx, xty, &sc);
    for (int i = 0; i < n; i++) {
        double* s1 = cs0 + (size_t)ids[i] * (p + 1);
        for (int j = 0; j < p; j++) s1[j] += (double)x[(size_t)j * n + i];
        s1[p] += (double)y[i];
        for (int a = 1; a < q; a++) {
            double* s2 = cs0 + cl * a + (size_t)ids[i] * (p + 1);
            const double xs = (double)x[(size_t)slope_cols[a - 1] * n + i];
            for (int j = 0; j < p; j++) s2[j] = fma(xs, (double)x[(size_t)j * n + i], s2[j]);
            s2[p] = fma(xs, (double)y[i], s2[p]);
        }
        cn[ids[i]]++;
    }
    LmmCtx* c = (LmmCtx*)calloc(1, sizeof(LmmCtx));
    c->p = p; c->n_rows = (uint32_t)n; c->xtx = xtx; c->xty = xty; c->yty = sc.yty;
    SlopeData d = {{cs0, cs0 + cl, cs0 + 2 * cl}, cn, nc, q, {0, slope_cols[0], n_slopes > 1 ? slope_cols[1] : 0}, 0};
    SlopeOut o;
    lmm_slope_fit(c, &d, &o);
    double u[64];
    for (int t = 0; t < p; t++) {
        betas[t] = c->beta[t];
        se[t] = sqrt(c->fit.sigma_sq * linv_norm_sq(c->L, c->inv_d, p, t, -1, u));
    }
    slope_components(o.theta, q, o.sigma_sq, vc);
    *converged = o.converged; *boundary = o.boundary_hit & 3; *n_eval = o.n_eval;
    free(c); free(cn); free(xtx); free(x); free(y);
    return ORC_OK;
}

int orc_fit_lmm_general(const double* xc, const double* yv, const int32_t* ids, int n, int p, int n_slopes, const int32_t* slope_cols,
                        int n_extra, double* betas, double* se, double* vc, int* converged, int* boundary, int* n_eval) {
    if (p > 64 || n_slopes < 0 || n_slopes > 2 || n_extra < 0 || n_extra > 2) return fail(ORC_ERR_INVALID_SPEC, "bad p / slope / grouping count");
    const int q0 = 1 + n_slopes;
    float* x = (float*)malloc(sizeof(float) * (size_t)n * p);
    float* y = (float*)malloc(sizeof(float) * (size_t)n);
    for (size_t i = 0; i < (size_t)n * p; i++) x[i] = (float)xc[i];
    for (int i = 0; i < n; i++) y[i] = (float)yv[i];
    int K = 0, E[2] = {0, 0};
    for (int i = 0; i < n; i++) {
        if (ids[i] + 1 > K) K = ids[i] + 1;
        for (int g = 0; g < n_extra; g++) if (ids[(size_t)(1 + g) * n + i] + 1 > E[g]) E[g] = ids[(size_t)(1 + g) * n + i] + 1;
    }
    const int Q = K * q0 + E[0] + E[1];
    size_t nt = (size_t)p * (p + 1) / 2;
    double* xtx = (double*)calloc(nt + (size_t)p, sizeof(double));
    double* xty = xtx + nt;
    double* ZtZ = (double*)calloc((size_t)Q * Q * 4 + (size_t)Q * (p + 1) * 2, sizeof(double));
    double* ZtXy = ZtZ + (size_t)Q * Q * 4;
    SuffScalars sc = {0.0, 0.0, 0};
    suff_add_rows(x, y, n, p, 0, n, xtx, xty, &sc);
    for (int i = 0; i < n; i++) {
        int cols[5];
        double zv[5], xy[65];
        int nz = 0;
        for (int a = 0; a < q0; a++) {
            cols[nz] = ids[i] * q0 + a;
            zv[nz++] = a == 0 ? 1.0 : (double)x[(size_t)slope_cols[a - 1] * n + i];
        }
        for (int g = 0; g < n_extra; g++) {
            cols[nz] = K * q0 + (g ? E[0] : 0) + ids[(size_t)(1 + g) * n + i];
            zv[nz++] = 1.0;
        }
        for (int j = 0; j < p; j++) xy[j] = (double)x[(size_t)j * n + i];
        xy[p] = (double)y[i];
        gen_add_row(ZtZ, ZtXy, Q, p, cols, zv, nz, xy);
    }
    LmmCtx* c = (LmmCtx*)calloc(1, sizeof(LmmCtx));
    c->p = p; c->n_rows = (uint32_t)n; c->xtx = xtx; c->xty = xty; c->yty = sc.yty;
    GenData d = {Q, q0, K, n_extra, {E[0], E[1]}, {K * q0, K * q0 + E[0]}, ZtZ, ZtXy, ZtZ + (size_t)Q * Q, ZtZ + (size_t)Q * Q * 2,
                 ZtZ + (size_t)Q * Q * 3, ZtXy + (size_t)Q * (p + 1), 0};
    SlopeOut o;
    lmm_general_fit(c, &d, &o);
    double u[64];
    for (int t = 0; t < p; t++) {
        betas[t] = c->beta[t];
        se[t] = sqrt(c->fit.sigma_sq * linv_norm_sq(c->L, c->inv_d, p, t, -1, u));
    }
    double pv[8];
    slope_components(o.theta, q0, o.sigma_sq, pv); /* {variances[q0], correlations, sigma²} */
    const int ncorr = q0 * (q0 - 1) / 2, D0 = q0 * (q0 + 1) / 2;
    int w = 0;
    for (int i = 0; i < q0 + ncorr; i++) vc[w++] = pv[i];
    for (int g = 0; g < n_extra; g++) vc[w++] = o.sigma_sq * (o.theta[D0 + g] * o.theta[D0 + g]);
    vc[w] = o.sigma_sq;
    *converged = o.converged; *boundary = o.boundary_hit & 3; *n_eval = o.n_eval;
    free(c); free(ZtZ); free(xtx); free(x); free(y);
    return ORC_OK;
}

int orc_fit_lmm_slope(const double* xc, const double* yv, const int32_t* ids, int n, int p, int slope_col, double* betas,
                      double* se, double* vc, int* converged, int* boundary, int* n_eval) {
    const int32_t cols[1] = {slope_col};
    return orc_fit_lmm_slopes(xc, yv, ids, n, p, 1, cols, betas, se, vc, converged, boundary, n_eval);
}

static void update_factor_exclusions(const OrcSpec* s, const float* x, int ld, int from, int to, uint32_t* counts, uint8_t* flags) {
    int off = 0, col = 1 + s->n_non_factor;
    for (int f = 0; f < s->n_factors; f++) {
        int l = s->factor_n_levels[f], width = l - 1;
        for (int i = from; i < to; i++) {
            int lvl = 0;
            for (int d = 0; d < width; d++)
                if (x[(size_t)(col + d) * ld + i] == 1.0f) { lvl = d + 1; break; }
            counts[off + lvl]++;
        }
        off += l;
        col += width;
    }
    off = 0;
    for (int f = 0; f < s->n_factors; f++) {
        int l = s->factor_n_levels[f];
        uint32_t mn = 0xffffffffu;
        for (int g = 0; g < l; g++) if (counts[off + g] < mn) mn = counts[off + g];
        flags[f] = s->factor_min_level_count > 0 && mn < (uint32_t)s->factor_min_level_count;
        off += l;
    }
}

/* batch.rs:2502-2562 */
static int build_exclusion_remap(const OrcSpec* s, const uint8_t* flags, int* kept, int* remap) {
    int k = s->n_non_factor, base = 1 + k, nfd = s->n_factor_dummies, P = base + nfd + s->n_interactions;
    uint8_t keep_dummy[256];
    for (int d = 0; d < nfd; d++) keep_dummy[d] = 1;
    int col = base;
    for (int f = 0; f < s->n_factors; f++) {
        int width = s->factor_n_levels[f] - 1;
        if (flags[f]) for (int d = 0; d < width; d++) keep_dummy[col + d - base] = 0;
        col += width;
    }
    for (int c = 0; c < P; c++) remap[c] = -1;
    int nk = 0;
    for (int c = 0; c < P; c++) {
        int keep = 1;
        if (c >= base && c < base + nfd) keep = keep_dummy[c - base];
        else if (c >= base + nfd) {
            int j = c - base - nfd;
            for (int q = s->inter_ptr[j]; q < s->inter_ptr[j + 1]; q++) {
                int cu = s->inter_cols[q];
                if (cu >= base && cu < base + nfd && !keep_dummy[cu - base]) keep = 0;
            }
        }
        if (keep && s->n_fit_columns > 0) {
            int found = 0;
            for (int q = 0; q < s->n_fit_columns; q++) if (s->fit_columns[q] == c) found = 1;
            keep = found;
        }
        if (keep) { remap[c] = nk; kept[nk++] = c; }
    }
    return nk;
}

static int run_one_sim(const OrcSpec* s, const uint32_t* ss, int n_ss, uint64_t sim_id, uint64_t base_seed, Ws* w,
                       const double* t_unc, const double* corr, const double* overall_crit, uint8_t* unc, uint8_t* cor,
                       uint8_t* conv, uint8_t* overall, uint8_t* fx, double* cap, const double* ph_unc_crit,
                       const double* ph_corr, uint8_t* ph_unc, uint8_t* ph_cor) {
    int rc = gen_sim(s, base_seed, sim_id, w);
    if (rc) return rc;
    const int P = w->P, T = s->n_marginal_targets + s->n_contrasts, ld = w->max_n;
    size_t nt = (size_t)P * (P + 1) / 2;
    int nprops = 0;
    for (int f = 0; f < s->n_factors; f++) nprops += s->factor_n_levels[f];
    memset(w->prefix_counts, 0, sizeof(uint32_t) * (size_t)(nprops > 0 ? nprops : 1));
    memset(w->excl_flags, 0, (size_t)(s->n_factors > 0 ? s->n_factors : 1));
    int test_reduces = s->n_fit_columns > 0 && s->n_fit_columns < P;

    memset(w->xtx, 0, sizeof(double) * nt);
    memset(w->xty, 0, sizeof(double) * (size_t)P);
    SuffScalars sc = {0.0, 0.0, 0};
    double L[64 * 65 / 2], Ared[64 * 65 / 2], bred[64], inv_d[64], wv[64], beta[64], u[64], all_t[256];
    int kept[64], remap[64];
    int last = 0;
    for (int g = 0; g < n_ss; g++) {
        int n = (int)ss[g];
        if (n > last) {
            suff_add_rows(w->x, w->y, ld, P, last, n, w->xtx, w->xty, &sc);
            if (s->n_factors > 0 && s->factor_min_level_count > 0)
                update_factor_exclusions(s, w->x, ld, last, n, w->prefix_counts, w->excl_flags);
            last = n;
        }
        int any_excl = 0;
        for (int f = 0; f < s->n_factors; f++) any_excl |= w->excl_flags[f] != 0;
        int reduced = any_excl || test_reduces;
        int p_red = P;
        OlsFit fit;
        const double *crit_unc_p = &t_unc[g], *corr_row = corr + (size_t)g * T, *ov_p = &overall_crit[g];
        double r_unc, r_corr[256], r_ov, r_ph, r_phc[64];
        const int NPH = ph_unc ? s->n_posthoc : 0;
        const double* ph_crit_p = NPH ? &ph_unc_crit[g] : NULL;
        const double* ph_corr_row = NPH ? ph_corr + (size_t)g * NPH : NULL;
        if (!reduced) {
            ols_solve(w->xtx, w->xty, sc.yty, sc.sum_y, sc.n_rows, P, L, inv_d, wv, beta, &fit);
        } else {
            p_red = build_exclusion_remap(s, w->excl_flags, kept, remap);
            for (int ri = 0; ri < p_red; ri++) {
                bred[ri] = w->xty[kept[ri]];
                for (int rj = 0; rj <= ri; rj++) Ared[TRI(ri, rj)] = w->xtx[TRI(kept[ri], kept[rj])];
            }
            ols_solve(Ared, bred, sc.yty, sc.sum_y, sc.n_rows, p_red, L, inv_d, wv, beta, &fit);
            /* reduced-df thresholds (batch.rs:2418-2437) */
            struct RedCrit* hit = NULL;
            for (int m = 0; m < w->n_red_memo; m++)
                if (w->red_memo[m].n == ss[g] && w->red_memo[m].p_red == p_red) hit = &w->red_memo[m];
            if (!hit) {
                w->red_memo = (struct RedCrit*)realloc(w->red_memo, sizeof(struct RedCrit) * (size_t)(w->n_red_memo + 1));
                hit = &w->red_memo[w->n_red_memo];
                hit->n = ss[g]; hit->p_red = p_red;
                rc = orc_build_crit_table(s, &ss[g], 1, p_red, &hit->unc, hit->corr, &hit->ph, &hit->ov);
                if (rc) return rc;
                if (s->n_posthoc > 0) {  /* crit_src carries the reduced-df post-hoc crits (batch.rs:1190-1203) */
                    rc = orc_posthoc_corr_rows(s, &ss[g], 1, p_red, hit->phc);
                    if (rc) return rc;
                }
                w->n_red_memo++;
            }
            (void)r_unc; (void)r_corr; (void)r_ov; (void)r_ph; (void)r_phc;
            crit_unc_p = &hit->unc; corr_row = hit->corr; ov_p = &hit->ov;
            if (NPH) { ph_crit_p = &hit->ph; ph_corr_row = hit->phc; }
        }
        conv[g] = (uint8_t)fit.converged;
        uint8_t* unc_row = unc + (size_t)g * T;
        uint8_t* cor_row = cor + (size_t)g * T;
        memset(unc_row, 0, (size_t)T);
        for (int t = 0; t < T; t++) all_t[t] = NAN;
        if (fit.converged) {
            for (int t = 0; t < s->n_marginal_targets; t++) {
                int c = reduced ? remap[s->target_indices[t]] : s->target_indices[t];
                if (c >= 0) all_t[t] = target_t_sq(L, inv_d, beta, fit.sigma_sq, p_red, c, -1, u);
            }
            for (int ci = 0; ci < s->n_contrasts; ci++) {
                int pc = s->contrast_pairs[2 * ci], nc = s->contrast_pairs[2 * ci + 1];
                if (reduced) { pc = remap[pc]; nc = remap[nc]; }
                all_t[s->n_marginal_targets + ci] = (pc >= 0 && nc >= 0) ? target_t_sq(L, inv_d, beta, fit.sigma_sq, p_red, pc, nc, u) : NAN;
            }
            for (int t = 0; t < T; t++) unc_row[t] = !isnan(all_t[t]) && all_t[t] > *crit_unc_p;
            orc_apply_correction(s->correction_method, all_t, T, *crit_unc_p, corr_row, cor_row);
        } else {
            memset(cor_row, 0, (size_t)T);
        }
        if (cap) memcpy(cap + (size_t)g * T, all_t, sizeof(double) * (size_t)T);
        if (NPH) {  /* evaluate_posthoc (posthoc.rs:40-148); under exclusion a contrast touching a dropped column is zeroed */
            uint8_t* pu = ph_unc + (size_t)g * NPH;
            uint8_t* pc_out = ph_cor + (size_t)g * NPH;
            double ph_t[64];
            memset(pu, 0, (size_t)NPH);
            memset(pc_out, 0, (size_t)NPH);
            if (fit.converged) {
                for (int c = 0; c < NPH; c++) {
                    int pos = s->posthoc_pairs[2 * c], neg = s->posthoc_pairs[2 * c + 1];
                    int pc = reduced ? remap[pos] : pos, nc = neg >= 0 ? (reduced ? remap[neg] : neg) : -1;
                    ph_t[c] = (pc >= 0 && (neg < 0 || nc >= 0)) ? target_t_sq(L, inv_d, beta, fit.sigma_sq, p_red, pc, nc, u) : NAN;
                    pu[c] = !isnan(ph_t[c]) && ph_t[c] > *ph_crit_p;
                }
                orc_apply_correction(s->correction_method, ph_t, NPH, *ph_crit_p, ph_corr_row, pc_out);
            }
        }
        if (overall) {
            uint8_t sig = 0;
            if (fit.converged) {
                double dfn = (double)(p_red - 1), dfd = fit.df_resid;
                if (dfn >= 1.0 && fit.sst > 1e-10 && fit.rss > 0.0 && isfinite(fit.rss) && isfinite(fit.sst)) {
                    double f = ((fit.sst - fit.rss) / dfn) / (fit.rss / dfd);
                    if (isfinite(f) && f > *ov_p) sig = 1;
                }
            }
            overall[g] = sig;
        }
        if (fx && s->n_factors > 0) memcpy(fx + (size_t)g * s->n_factors, w->excl_flags, (size_t)s->n_factors);
    }
    return ORC_OK;
}

/* batch.rs:2572-2597 */
static int included_factor_with_smallest_min_count(const OrcSpec* s, const uint32_t* counts, const uint8_t* flags) {
    int best = -1, off = 0;
    uint32_t best_min = 0;
    for (int f = 0; f < s->n_factors; f++) {
        int l = s->factor_n_levels[f];
        if (flags[f] == 0) {
            uint32_t mn = 0xffffffffu;
            for (int g = 0; g < l; g++) if (counts[off + g] < mn) mn = counts[off + g];
            if (best < 0 || mn < best_min) { best = f; best_min = mn; }
        }
        off += l;
    }
    return best;
}

/* GLM arm (batch.rs:1562-2135, unclustered): refit per (sim, N) on the row prefix from the truth start;
 * sparse-factor exclusion recounted from row 0; separation fallback = drop the sparsest included factor, cold start. */
static int run_one_sim_glm(const OrcSpec* s, const uint32_t* ss, int n_ss, uint64_t sim_id, uint64_t base_seed, Ws* w,
                           const double* t_unc, const double* corr, const double* overall_crit, uint8_t* unc, uint8_t* cor,
                           uint8_t* conv, uint8_t* overall, uint8_t* fx, double* cap) {
    int rc = gen_sim(s, base_seed, sim_id, w);
    if (rc) return rc;
    const int P = w->P, T = s->n_marginal_targets, ld = w->max_n;
    const int family = s->outcome_kind == ORC_OUT_COUNT ? 2 : (s->link_probit ? 1 : 0);
    int nprops = 0;
    for (int f = 0; f < s->n_factors; f++) nprops += s->factor_n_levels[f];
    int test_reduces = s->n_fit_columns > 0 && s->n_fit_columns < P;
    double A[64 * 65 / 2], L[64 * 65 / 2], b[64], inv_d[64], wv[64], beta[64], u[64], all_t[256], bstart[64];
    int kept[64], remap[64];
    for (int g = 0; g < n_ss; g++) {
        int n = (int)ss[g];
        memset(w->prefix_counts, 0, sizeof(uint32_t) * (size_t)(nprops > 0 ? nprops : 1));
        memset(w->excl_flags, 0, (size_t)(s->n_factors > 0 ? s->n_factors : 1));
        if (s->n_factors > 0 && s->factor_min_level_count > 0)
            update_factor_exclusions(s, w->x, ld, 0, n, w->prefix_counts, w->excl_flags);
        GlmOut go;
        int p_red = P, reduced = 0;
        for (int attempt = 0; attempt < 2; attempt++) {
            int any_excl = 0;
            for (int f = 0; f < s->n_factors; f++) any_excl |= w->excl_flags[f] != 0;
            reduced = any_excl || test_reduces;
            const double* start = NULL;
            if (!reduced) {
                p_red = P;
                if (attempt == 0) start = s->effect_sizes;
                glm_irls(w->x, w->y, ld, n, P, NULL, family, start, A, b, L, inv_d, wv, beta, &go);
            } else {
                p_red = build_exclusion_remap(s, w->excl_flags, kept, remap);
                for (int j = 0; j < p_red; j++) bstart[j] = s->effect_sizes[kept[j]];
                if (attempt == 0) start = bstart;
                glm_irls(w->x, w->y, ld, n, p_red, kept, family, start, A, b, L, inv_d, wv, beta, &go);
            }
            if (go.converged || attempt == 1) break;
            if (s->factor_min_level_count == 0) break;
            int f = included_factor_with_smallest_min_count(s, w->prefix_counts, w->excl_flags);
            if (f < 0) break;
            w->excl_flags[f] = 2;
        }
        conv[g] = (uint8_t)go.converged;
        uint8_t* unc_row = unc + (size_t)g * T;
        uint8_t* cor_row = cor + (size_t)g * T;
        memset(unc_row, 0, (size_t)T);
        memset(cor_row, 0, (size_t)T);
        for (int t = 0; t < T; t++) all_t[t] = NAN;
        if (go.converged) {
            for (int t = 0; t < T; t++) {
                int c = reduced ? remap[s->target_indices[t]] : s->target_indices[t];
                if (c >= 0) all_t[t] = target_t_sq(L, inv_d, beta, 1.0, p_red, c, -1, u);
            }
            for (int t = 0; t < T; t++) unc_row[t] = !isnan(all_t[t]) && all_t[t] > t_unc[g];
            orc_apply_correction(s->correction_method, all_t, T, t_unc[g], corr + (size_t)g * T, cor_row);
        }
        if (cap) memcpy(cap + (size_t)g * T, all_t, sizeof(double) * (size_t)T);
        if (overall) {
            double crit_ov = overall_crit[g];
            if (reduced) {
                double r_unc;
                rc = orc_build_crit_table(s, &ss[g], 1, p_red, &r_unc, NULL, NULL, &crit_ov);
                if (rc) return rc;
            }
            uint8_t sig = 0;
            if (go.converged && isfinite(go.deviance) && isfinite(go.deviance_null)) {
                double lrt = go.deviance_null - go.deviance;
                if (isfinite(lrt) && lrt > crit_ov) sig = 1;
            }
            overall[g] = sig;
        }
        if (fx && s->n_factors > 0) memcpy(fx + (size_t)g * s->n_factors, w->excl_flags, (size_t)s->n_factors);
    }
    return ORC_OK;
}

/* MLE arm (batch.rs:1214-1561), random intercept only: cluster-level sufficient statistics grow with the row
 * prefix (cluster ids are N-independent in both sizing regimes, generation.rs cluster_of_row), one profiled-REML
 * fit per grid point; z-based thresholds; joint Wald chi-square over the testable targets. */
static int run_one_sim_mle(const OrcSpec* s, const uint32_t* ss, int n_ss, uint64_t sim_id, uint64_t base_seed, Ws* w,
                           const double* t_unc, const double* corr, double joint_crit_full, uint8_t* unc, uint8_t* cor,
                           uint8_t* conv, uint8_t* overall, uint8_t* fx, double* cap, uint8_t* joint, uint8_t* bh) {
    int rc = gen_sim(s, base_seed, sim_id, w);
    if (rc) return rc;
    const int P = w->P, T = s->n_marginal_targets, ld = w->max_n;
    const int csz = s->sizing_value > 0 ? s->sizing_value : 1;
    const int nc_cap = s->sizing_fixed_size ? ld / csz + 1 : csz;
    size_t nt = (size_t)P * (P + 1) / 2;
    int nprops = 0;
    for (int f = 0; f < s->n_factors; f++) nprops += s->factor_n_levels[f];
    int test_reduces = s->n_fit_columns > 0 && s->n_fit_columns < P;
    double* csum = (double*)calloc((size_t)nc_cap * (P + 1) * 4, sizeof(double));
    double* csum_red = csum + (size_t)nc_cap * (P + 1);
    double* csum2 = csum_red + (size_t)nc_cap * (P + 1); /* random slopes: per-cluster sums of x_s * [x | y], one block per slope */
    const size_t cs_block = (size_t)nc_cap * (P + 1);
    uint32_t* cn = (uint32_t*)calloc((size_t)nc_cap, sizeof(uint32_t));
    LmmCtx* c = (LmmCtx*)calloc(1, sizeof(LmmCtx));
    /* extra groupings: dense Z'Z / Z'[X | y] over the capacity layout (levels that have no row yet stay all-zero columns: they
     * contribute exact zeros to every chain) */
    const int gq0 = 1 + s->n_slopes, gK = nc_cap;
    const int gE0 = s->n_extra > 0 ? orc_extra_levels_at(s, 0, ld) : 0, gE1 = s->n_extra > 1 ? orc_extra_levels_at(s, 1, ld) : 0;
    const int gQ = gK * gq0 + gE0 + gE1, gQr = gK + gE0 + gE1;
    double *gZtZ = NULL, *gZtXy = NULL, *gwork = NULL;
    if (s->n_extra > 0) {
        gZtZ = (double*)calloc((size_t)gQ * gQ + (size_t)gQ * (P + 1), sizeof(double));
        gZtXy = gZtZ + (size_t)gQ * gQ;
        gwork = (double*)calloc((size_t)gQ * gQ * 5 + (size_t)gQ * (P + 1) * 3, sizeof(double));
    }
    memset(w->xtx, 0, sizeof(double) * nt);
    memset(w->xty, 0, sizeof(double) * (size_t)P);
    SuffScalars sc = {0.0, 0.0, 0};
    double Ared[64 * 65 / 2], bred[64], all_t[256], u[64];
    int kept[64], remap[64], cols[256];
    int last = 0;
    for (int g = 0; g < n_ss; g++) {
        int n = (int)ss[g];
        if (n > last) {
            suff_add_rows(w->x, w->y, ld, P, last, n, w->xtx, w->xty, &sc);
            for (int i = last; i < n; i++) {
                int k = w->cluster_ids[i];
                double* sx = csum + (size_t)k * (P + 1);
                for (int j = 0; j < P; j++) sx[j] += (double)w->x[(size_t)j * ld + i];
                sx[P] += (double)w->y[i];
                for (int a = 0; a < s->n_slopes; a++) {
                    double* s2 = csum2 + cs_block * a + (size_t)k * (P + 1);
                    const double xs = (double)w->x[(size_t)s->slope_col[a] * ld + i];
                    for (int j = 0; j < P; j++) s2[j] = fma(xs, (double)w->x[(size_t)j * ld + i], s2[j]);
                    s2[P] = fma(xs, (double)w->y[i], s2[P]);
                }
                cn[k]++;
                if (s->n_extra > 0) {
                    int cols[5];
                    double zv[5], xy[65];
                    int nz = 0;
                    for (int a = 0; a < gq0; a++) {
                        cols[nz] = k * gq0 + a;
                        zv[nz++] = a == 0 ? 1.0 : (double)w->x[(size_t)s->slope_col[a - 1] * ld + i];
                    }
                    for (int e = 0; e < s->n_extra; e++) {
                        cols[nz] = gK * gq0 + (e ? gE0 : 0) + orc_extra_level_of_row(s, e, i);
                        zv[nz++] = 1.0;
                    }
                    for (int j = 0; j < P; j++) xy[j] = (double)w->x[(size_t)j * ld + i];
                    xy[P] = (double)w->y[i];
                    gen_add_row(gZtZ, gZtXy, gQ, P, cols, zv, nz, xy);
                }
            }
            last = n;
        }
        memset(w->prefix_counts, 0, sizeof(uint32_t) * (size_t)(nprops > 0 ? nprops : 1));
        memset(w->excl_flags, 0, (size_t)(s->n_factors > 0 ? s->n_factors : 1));
        if (s->n_factors > 0 && s->factor_min_level_count > 0)
            update_factor_exclusions(s, w->x, ld, 0, n, w->prefix_counts, w->excl_flags);
        int any_excl = 0;
        for (int f = 0; f < s->n_factors; f++) any_excl |= w->excl_flags[f] != 0;
        int reduced = any_excl || test_reduces, p_red = P;
        int n_clusters = s->sizing_fixed_size ? (n + csz - 1) / csz : csz;
        LmmOut lo;
        c->n_rows = sc.n_rows; c->yty = sc.yty;
        if (s->n_extra > 0) {
            /* extra groupings: the dense penalised-least-squares fit; a reduced fit keeps every grouping's intercept and strips the
             * primary slopes (batch.rs:1357-1375) */
            SlopeOut so;
            if (!reduced) {
                c->p = P; c->xtx = w->xtx; c->xty = w->xty;
                GenData gd = {gQ, gq0, gK, s->n_extra, {gE0, gE1}, {gK * gq0, gK * gq0 + gE0}, gZtZ, gZtXy, gwork, gwork + (size_t)gQ * gQ,
                              gwork + (size_t)gQ * gQ * 2, gwork + (size_t)gQ * gQ * 3, 0};
                lmm_general_fit(c, &gd, &so);
            } else {
                p_red = build_exclusion_remap(s, w->excl_flags, kept, remap);
                for (int ri = 0; ri < p_red; ri++) {
                    bred[ri] = w->xty[kept[ri]];
                    for (int rj = 0; rj <= ri; rj++) Ared[TRI(ri, rj)] = w->xtx[TRI(kept[ri], kept[rj])];
                }
                double* rZ = gwork + (size_t)gQ * gQ * 3 + (size_t)gQ * (P + 1); /* behind R: Qr x Qr, then Qr x (p_red + 1) */
                double* rX = rZ + (size_t)gQr * gQr;
                for (int i = 0; i < gQr; i++) {
                    const int fi = i < gK ? i * gq0 : gK * gq0 + (i - gK);
                    for (int j = 0; j <= i; j++) {
                        const int fj = j < gK ? j * gq0 : gK * gq0 + (j - gK);
                        rZ[(size_t)i * gQr + j] = gZtZ[(size_t)fi * gQ + fj];
                    }
                    for (int rj = 0; rj < p_red; rj++) rX[(size_t)i * (p_red + 1) + rj] = gZtXy[(size_t)fi * (P + 1) + kept[rj]];
                    rX[(size_t)i * (p_red + 1) + p_red] = gZtXy[(size_t)fi * (P + 1) + P];
                }
                c->p = p_red; c->xtx = Ared; c->xty = bred;
                GenData gd = {gQr, 1, gK, s->n_extra, {gE0, gE1}, {gK, gK + gE0}, rZ, rX, gwork, gwork + (size_t)gQ * gQ,
                              gwork + (size_t)gQ * gQ * 2, gwork + (size_t)gQ * gQ * 3, 0};
                lmm_general_fit(c, &gd, &so);
            }
            lo.converged = so.converged; lo.boundary_hit = so.boundary_hit; lo.theta = so.theta[0]; lo.sigma_sq = so.sigma_sq;
            lo.tau_sq = so.theta[0] * so.theta[0] * so.sigma_sq;
        } else if (!reduced && s->n_slopes > 0) {
            /* full model with the random slopes: 3- / 6-parameter REML; the reduced fit below strips the slopes like the reference
             * (batch.rs:1357-1375) */
            c->p = P; c->xtx = w->xtx; c->xty = w->xty;
            SlopeData sd = {{csum, csum2, csum2 + cs_block}, cn, n_clusters, 1 + s->n_slopes,
                            {0, s->slope_col[0], s->n_slopes > 1 ? s->slope_col[1] : 0}, 0};
            SlopeOut so;
            lmm_slope_fit(c, &sd, &so);
            lo.converged = so.converged; lo.boundary_hit = so.boundary_hit; lo.theta = so.theta[0]; lo.sigma_sq = so.sigma_sq;
            lo.tau_sq = so.theta[0] * so.theta[0] * so.sigma_sq;
        } else if (!reduced) {
            c->p = P; c->xtx = w->xtx; c->xty = w->xty;
            lmm_fit(c, csum, cn, n_clusters, P, &lo);
        } else {
            p_red = build_exclusion_remap(s, w->excl_flags, kept, remap);
            for (int ri = 0; ri < p_red; ri++) {
                bred[ri] = w->xty[kept[ri]];
                for (int rj = 0; rj <= ri; rj++) Ared[TRI(ri, rj)] = w->xtx[TRI(kept[ri], kept[rj])];
            }
            for (int k = 0; k < n_clusters; k++) {
                for (int rj = 0; rj < p_red; rj++) csum_red[(size_t)k * (p_red + 1) + rj] = csum[(size_t)k * (P + 1) + kept[rj]];
                csum_red[(size_t)k * (p_red + 1) + p_red] = csum[(size_t)k * (P + 1) + P];
            }
            c->p = p_red; c->xtx = Ared; c->xty = bred;
            lmm_fit(c, csum_red, cn, n_clusters, p_red, &lo);
        }
        conv[g] = (uint8_t)lo.converged;
        if (bh) bh[g] = (uint8_t)lo.boundary_hit;
        uint8_t* unc_row = unc + (size_t)g * T;
        uint8_t* cor_row = cor + (size_t)g * T;
        memset(unc_row, 0, (size_t)T);
        memset(cor_row, 0, (size_t)T);
        for (int t = 0; t < T; t++) all_t[t] = NAN;
        int k_red = 0;
        for (int t = 0; t < T; t++) {
            int col = reduced ? remap[s->target_indices[t]] : s->target_indices[t];
            if (col >= 0) cols[k_red++] = col;
        }
        uint8_t jsig = 0;
        if (lo.converged) {
            for (int t = 0; t < T; t++) {
                int col = reduced ? remap[s->target_indices[t]] : s->target_indices[t];
                if (col >= 0) all_t[t] = target_t_sq(c->L, c->inv_d, c->beta, c->fit.sigma_sq, p_red, col, -1, u);
            }
            for (int t = 0; t < T; t++) unc_row[t] = !isnan(all_t[t]) && all_t[t] > t_unc[g];
            orc_apply_correction(s->correction_method, all_t, T, t_unc[g], corr + (size_t)g * T, cor_row);
            double jc = !reduced ? joint_crit_full : (k_red == 0 ? INFINITY : orc_chi2_ppf(1.0 - s->alpha, (double)k_red));
            double jt = lmm_joint_wald(c, cols, k_red);
            jsig = isfinite(jt) && jt > jc;
        }
        if (joint) joint[g] = jsig;
        if (cap) memcpy(cap + (size_t)g * T, all_t, sizeof(double) * (size_t)T);
        if (overall) overall[g] = 0;
        if (fx && s->n_factors > 0) memcpy(fx + (size_t)g * s->n_factors, w->excl_flags, (size_t)s->n_factors);
    }
    free(c); free(cn); free(csum); free(gZtZ); free(gwork);
    return ORC_OK;
}

int orc_run_batch(const OrcSpec* s, const uint32_t* ss, int n_ss, uint32_t n_sims, uint64_t sim_begin, uint64_t base_seed,
                  int n_threads, uint8_t* uncorrected, uint8_t* corrected, uint8_t* converged, uint8_t* overall,
                  uint8_t* factor_excluded, double* t_sq_capture, uint8_t* joint, uint8_t* boundary_hit,
                  uint8_t* posthoc_unc, uint8_t* posthoc_cor) {
    int rc = validate_spec(s, ss, n_ss);
    if (rc) return rc;
    if (s->n_posthoc > 0 && s->estimator != ORC_EST_OLS)
        return fail(ORC_ERR_INVALID_SPEC, "posthoc tests are only supported for estimator=Ols");
    if (s->n_posthoc > 64) return fail(ORC_ERR_INVALID_SPEC, "posthoc contrasts > 64");
    if (s->estimator == ORC_EST_MLE && (s->n_contrasts > 0 || s->outcome_kind != ORC_OUT_CONTINUOUS))
        return fail(ORC_ERR_INVALID_SPEC, "oracle run_batch: only the Gaussian random-intercept MLE arm is restated");
    if (s->estimator == ORC_EST_GLM && (s->has_cluster || s->n_contrasts > 0))
        return fail(ORC_ERR_INVALID_SPEC, "oracle run_batch: clustered GLM (GLMM) / GLM contrasts are not restated");
    if (s->estimator == ORC_EST_GLM && s->outcome_kind == ORC_OUT_CONTINUOUS)
        return fail(ORC_ERR_INVALID_SPEC, "estimator Glm requires a binary or count outcome");
    const int P = 1 + s->n_non_factor + s->n_factor_dummies + s->n_interactions;
    if (P > 64) return fail(ORC_ERR_INVALID_SPEC, "P > 64");
    const int T = s->n_marginal_targets + s->n_contrasts;
    if (T > 256) return fail(ORC_ERR_INVALID_SPEC, "T > 256");
    const double joint_crit_full = (s->estimator == ORC_EST_MLE && T > 0) ? orc_chi2_ppf(1.0 - s->alpha, (double)T) : NAN;
    double* t_unc = (double*)malloc(sizeof(double) * (size_t)n_ss * (size_t)(T + 3));
    double *corr = t_unc + n_ss, *ovc = corr + (size_t)n_ss * T, *ph = ovc + n_ss;
    rc = orc_build_crit_table(s, ss, n_ss, P, t_unc, corr, ph, ovc);
    if (rc) { free(t_unc); return rc; }
    const int NPH = (posthoc_unc && posthoc_cor) ? s->n_posthoc : 0;
    double* ph_corr = NULL;
    if (NPH) {
        ph_corr = (double*)malloc(sizeof(double) * (size_t)n_ss * (size_t)NPH);
        rc = orc_posthoc_corr_rows(s, ss, n_ss, P, ph_corr);
        if (rc) { free(t_unc); free(ph_corr); return rc; }
    }
    if (!g_t3_ready) t3_build();
    const int max_n = (int)ss[n_ss - 1];
    int err = 0;
#ifdef _OPENMP
    if (n_threads < 1) n_threads = 1;
#pragma omp parallel num_threads(n_threads)
#endif
    {
        Ws* w = ws_new(s, max_n);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 64)
#endif
        for (int64_t i = 0; i < (int64_t)n_sims; i++) {
            if (err) continue;
            int r;
            if (s->estimator == ORC_EST_MLE)
                r = run_one_sim_mle(s, ss, n_ss, sim_begin + (uint64_t)i, base_seed, w, t_unc, corr, joint_crit_full,
                                    uncorrected + (size_t)i * n_ss * T, corrected + (size_t)i * n_ss * T,
                                    converged + (size_t)i * n_ss, overall ? overall + (size_t)i * n_ss : NULL,
                                    factor_excluded ? factor_excluded + (size_t)i * n_ss * s->n_factors : NULL,
                                    t_sq_capture ? t_sq_capture + (size_t)i * n_ss * T : NULL,
                                    joint ? joint + (size_t)i * n_ss : NULL, boundary_hit ? boundary_hit + (size_t)i * n_ss : NULL);
            else if (s->estimator == ORC_EST_GLM)
                r = run_one_sim_glm(s, ss, n_ss, sim_begin + (uint64_t)i, base_seed, w, t_unc, corr, ovc,
                                    uncorrected + (size_t)i * n_ss * T, corrected + (size_t)i * n_ss * T,
                                    converged + (size_t)i * n_ss, overall ? overall + (size_t)i * n_ss : NULL,
                                    factor_excluded ? factor_excluded + (size_t)i * n_ss * s->n_factors : NULL,
                                    t_sq_capture ? t_sq_capture + (size_t)i * n_ss * T : NULL);
            else
                r = run_one_sim(s, ss, n_ss, sim_begin + (uint64_t)i, base_seed, w, t_unc, corr, ovc,
                                uncorrected + (size_t)i * n_ss * T, corrected + (size_t)i * n_ss * T,
                                converged + (size_t)i * n_ss, overall ? overall + (size_t)i * n_ss : NULL,
                                factor_excluded ? factor_excluded + (size_t)i * n_ss * s->n_factors : NULL,
                                t_sq_capture ? t_sq_capture + (size_t)i * n_ss * T : NULL, ph, ph_corr,
                                NPH ? posthoc_unc + (size_t)i * n_ss * NPH : NULL, NPH ? posthoc_cor + (size_t)i * n_ss * NPH : NULL);
            if (r) {
#ifdef _OPENMP
#pragma omp critical(orc_err)
#endif
                err = r;
            }
        }
        ws_free(w);
    }
    free(t_unc);
    free(ph_corr);
    return err;
}
