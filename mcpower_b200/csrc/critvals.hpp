// critvals.hpp — host-side critical values (f64), restating
// crates/engine-core/src/critvals.rs so the thresholds the kernel compares against are the
// reference's, operation for operation: Acklam norm_ppf without refinement (:33-87), Lanczos
// ln_gamma (:112-135), NR betacf/betai (:91-186), Newton t_ppf (:211-284), gamma_inc_p (:294-345),
// chi2_ppf (:356-406), f_ppf (:416-485), studentized range by composite GL32 (:493-746).
// The kernel never computes a p-value (critvals.rs:1-9): it only compares t² with these.
#pragma once
#include <cmath>
#include <limits>
#include <stdexcept>
#include <vector>

namespace mcp {
namespace crit {

constexpr double kPi = 3.14159265358979323846;
constexpr double kLn2 = 0.693147180559945309417;
constexpr double kSqrt2 = 1.41421356237309504880;
constexpr double kInf = std::numeric_limits<double>::infinity();
constexpr double kNaN = std::numeric_limits<double>::quiet_NaN();

inline double horner_up(const double* c, int n, double x) {  // ((c0 x + c1) x + c2) ...
    double a = c[0];
    for (int i = 1; i < n; i++) a = a * x + c[i];
    return a;
}

inline double norm_ppf(double p) {
    static const double A[6] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
                                1.38357751867269e2,     -3.066479806614716e+01, 2.506628277459239e+00};
    static const double B[6] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
                                6.680131188771972e+01,  -1.328068155288572e+01, 1.0};
    static const double C[6] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
                                -2.549732539343734e+00, 4.374664141464968e+00,  2.938163982698783e+00};
    static const double D[5] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00,
                                3.754408661907416e+00, 1.0};
    const double p_low = 0.02425, p_high = 1.0 - p_low;
    if (p <= 0.0) return -kInf;
    if (p >= 1.0) return kInf;
    if (p < p_low) {
        const double q = std::sqrt(-2.0 * std::log(p));
        return horner_up(C, 6, q) / horner_up(D, 5, q);
    }
    if (p <= p_high) {
        const double q = p - 0.5, r = q * q;
        return horner_up(A, 6, r) * q / horner_up(B, 6, r);
    }
    const double q = std::sqrt(-2.0 * std::log(1.0 - p));
    return -horner_up(C, 6, q) / horner_up(D, 5, q);
}

inline double ln_gamma(double x) {
    static const double P[9] = {0.9999999999998099,   676.5203681218851,    -1259.1392167224028,
                                771.3234287776531,    -176.6150291621406,   12.507343278686905,
                                -0.13857109526572012, 9.984369578019572e-6, 1.5056327351493116e-7};
    if (x < 0.5) return std::log(kPi / std::sin(kPi * x)) - ln_gamma(1.0 - x);
    x -= 1.0;
    double a = P[0];
    const double t = x + 7.0 + 0.5;
    for (int i = 1; i < 9; i++) a += P[i] / (x + (double)i);
    return 0.5 * std::log(2.0 * kPi) + (x + 0.5) * std::log(t) - t + std::log(a);
}
inline double ln_beta(double a, double b) { return ln_gamma(a) + ln_gamma(b) - ln_gamma(a + b); }

inline double clampmin(double v, double fpmin) { return std::fabs(v) < fpmin ? fpmin : v; }

inline double betacf(double a, double b, double x) {
    const int maxit = 200;
    const double eps = 3e-15, fpmin = 1.0e-300;
    const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0;
    double d = 1.0 / clampmin(1.0 - qab * x / qap, fpmin);
    double h = d;
    for (int m = 1; m <= maxit; m++) {
        const double mf = (double)m, m2 = 2.0 * mf;
        double aa = mf * (b - mf) * x / ((qam + m2) * (a + m2));
        d = clampmin(1.0 + aa * d, fpmin);
        c = clampmin(1.0 + aa / c, fpmin);
        d = 1.0 / d;
        h *= d * c;
        aa = -(a + mf) * (qab + mf) * x / ((a + m2) * (qap + m2));
        d = clampmin(1.0 + aa * d, fpmin);
        c = clampmin(1.0 + aa / c, fpmin);
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (std::fabs(del - 1.0) < eps) break;
    }
    return h;
}

inline double betai(double a, double b, double x) {
    if (x <= 0.0) return 0.0;
    if (x >= 1.0) return 1.0;
    const double bt = std::exp(a * std::log(x) + b * std::log(1.0 - x) - ln_beta(a, b));
    if (x < (a + 1.0) / (a + b + 2.0)) return bt * betacf(a, b, x) / a;
    return 1.0 - bt * betacf(b, a, 1.0 - x) / b;
}

inline double t_cdf(double x, double df) {
    if (!std::isfinite(x)) return x > 0.0 ? 1.0 : 0.0;
    const double z = df / (df + x * x);
    const double half = 0.5 * betai(0.5 * df, 0.5, z);
    return x >= 0.0 ? 1.0 - half : half;
}

inline double t_ppf(double p, double df) {
    if (!std::isfinite(df) || df <= 0.0) return kNaN;
    if (p <= 0.0) return -kInf;
    if (p >= 1.0) return kInf;
    if (p == 0.5) return 0.0;
    const bool upper = p > 0.5;
    const double q = upper ? p : 1.0 - p;
    double x;
    if (df > 200.0) {  // Cornish–Fisher seed
        const double z = norm_ppf(q), z2 = z * z;
        x = z + (z * (z2 + 1.0)) / (4.0 * df) + (z * (5.0 * z2 * z2 + 16.0 * z2 + 3.0)) / (96.0 * df * df);
    } else if (std::fabs(df - 1.0) < 1e-9) {
        x = std::fabs(std::tan(kPi * (q - 0.5)));
    } else if (std::fabs(df - 2.0) < 1e-9) {
        const double alpha = 4.0 * q * (1.0 - q);
        x = std::sqrt(2.0 / alpha - 2.0) * (q > 0.5 ? 1.0 : -1.0);
    } else {
        x = norm_ppf(q);
    }
    if (df > 1.0e5) return upper ? x : -x;
    const double lnc = ln_gamma(0.5 * (df + 1.0)) - ln_gamma(0.5 * df) - 0.5 * std::log(df * kPi);
    for (int it = 0; it < 80; it++) {
        const double cdf = t_cdf(x, df);
        const double pdf = std::exp(lnc - 0.5 * (df + 1.0) * std::log(1.0 + x * x / df));
        if (pdf <= 0.0 || !std::isfinite(pdf)) break;
        const double nx = x - (cdf - q) / pdf;
        if (!std::isfinite(nx)) break;
        const double rel = std::fabs(nx - x) / (1.0 + std::fabs(x));
        x = nx;
        if (rel < 1e-14) break;
    }
    return upper ? x : -x;
}

inline double gamma_inc_p(double a, double x) {
    if (x <= 0.0 || a <= 0.0) return 0.0;
    const int maxit = 200;
    const double eps = 3e-15, fpmin = 1.0e-300;
    const double front = -x + a * std::log(x) - ln_gamma(a);
    if (x < a + 1.0) {
        double ap = a, sum = 1.0 / a, term = sum;
        for (int i = 0; i < maxit; i++) {
            ap += 1.0;
            term *= x / ap;
            sum += term;
            if (std::fabs(term) < std::fabs(sum) * eps) break;
        }
        return sum * std::exp(front);
    }
    double b = x + 1.0 - a, c = 1.0 / fpmin, d = 1.0 / b, h = d;
    for (int i = 1; i < maxit; i++) {
        const double an = -(double)i * ((double)i - a);
        b += 2.0;
        d = clampmin(an * d + b, fpmin);
        c = clampmin(b + an / c, fpmin);
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (std::fabs(del - 1.0) < eps) break;
    }
    return 1.0 - std::exp(front) * h;
}

inline double cube(double v) { return v * v * v; }

inline double chi2_ppf(double p, double k) {
    if (!std::isfinite(k) || k <= 0.0) return kNaN;
    if (p <= 0.0) return 0.0;
    if (p >= 1.0) return kInf;
    const double z = norm_ppf(p), h = 2.0 / (9.0 * k);
    double x = k * cube(1.0 - h + z * std::sqrt(h));  // Wilson–Hilferty seed
    if (!std::isfinite(x) || x <= 0.0) x = std::fmax(k, 1e-6);
    const double a = 0.5 * k;
    for (int it = 0; it < 80; it++) {
        const double cdf = gamma_inc_p(a, 0.5 * x);
        const double pdf = std::exp((a - 1.0) * std::log(0.5 * x) - 0.5 * x - kLn2 - ln_gamma(a));
        if (!std::isfinite(pdf) || pdf <= 0.0) break;
        double nx = x - (cdf - p) / pdf;
        if (!std::isfinite(nx) || nx <= 0.0) nx = 0.5 * x;
        const double rel = std::fabs(nx - x) / (1.0 + std::fabs(x));
        x = nx;
        if (rel < 1e-13) break;
    }
    return x;
}

inline double f_pdf(double x, double dfn, double dfd) {
    if (x <= 0.0) return 0.0;
    const double a = 0.5 * dfn, b = 0.5 * dfd;
    return std::exp(a * std::log(dfn / dfd) + (a - 1.0) * std::log(x) - (a + b) * std::log(1.0 + dfn / dfd * x) - ln_beta(a, b));
}
inline double f_cdf(double x, double dfn, double dfd) {
    if (x <= 0.0) return 0.0;
    if (!std::isfinite(x)) return 1.0;
    return 1.0 - betai(0.5 * dfd, 0.5 * dfn, dfd / (dfd + dfn * x));
}
inline double f_ppf(double p, double dfn, double dfd) {
    if (!std::isfinite(dfn) || !std::isfinite(dfd) || dfn <= 0.0 || dfd <= 0.0) return kNaN;
    if (p <= 0.0) return 0.0;
    if (p >= 1.0) return kInf;
    const double z = norm_ppf(p), h = 2.0 / (9.0 * dfn);
    double x = std::fmax(dfn * cube(1.0 - h + z * std::sqrt(h)) / dfn, 1e-6);
    for (int it = 0; it < 80; it++) {
        const double cdf = f_cdf(x, dfn, dfd), pdf = f_pdf(x, dfn, dfd);
        if (!std::isfinite(pdf) || pdf <= 0.0) break;
        double nx = x - (cdf - p) / pdf;
        if (!std::isfinite(nx) || nx <= 0.0) nx = 0.5 * x;
        const double rel = std::fabs(nx - x) / (1.0 + std::fabs(x));
        x = nx;
        if (rel < 1e-13) break;
    }
    return x;
}

// ---- studentized range (Tukey HSD).  phi here is the engine's A&S Φ composed with libm exp
// (the reference composes glmm's exp kernel, <=1 ULP of libm; the quantile is converged to 1e-9 anyway).
inline double as_phi(double z) {
    const double x = -z * 0.70710678118654752440;
    const double sign = x < 0.0 ? -1.0 : 1.0, ax = std::fabs(x);
    const double t = 1.0 / (1.0 + 0.3275911 * ax);
    const double y = 1.0 - (((((1.061405429 * t + -1.453152027) * t) + 1.421413741) * t + -0.284496736) * t + 0.254829592) * t *
                               std::exp(-ax * ax);
    return 0.5 * (1.0 - sign * y);
}
static const double GL32_X[16] = {0.04830766568773832, 0.1444719615827965, 0.23928736225213707, 0.33186860228212765,
                                  0.42135127613063535, 0.5068999089322294, 0.5877157572407623,  0.6630442669302152,
                                  0.7321821187402897,  0.7944837959679424, 0.84936761373257,    0.8963211557660521,
                                  0.9349060759377397,  0.9647622555875064, 0.9856115115452683,  0.9972638618494816};
static const double GL32_W[16] = {0.0965400885147278,  0.09563872007927486, 0.09384439908080457, 0.09117387869576389,
                                  0.08765209300440381, 0.08331192422694676, 0.07819389578707031, 0.0723457941088485,
                                  0.06582222277636185, 0.05868409347853555, 0.05099805926237618, 0.04283589802222669,
                                  0.03427386291302143, 0.02539206530926206, 0.01627439473090567, 0.007018610009470117};

inline double range_cdf(double w, double k) {
    if (w <= 0.0) return 0.0;
    const double lo = -8.0, hi = 8.0;
    const int kk = (int)k;
    const int panels = 4 + (kk > 0 ? kk - 1 : 0);
    const double step = (hi - lo) / (double)panels, half = 0.5 * step;
    auto f = [&](double z) {
        const double inner = as_phi(z) - as_phi(z - w);
        return inner > 0.0 ? std::exp(-0.5 * z * z) / std::sqrt(2.0 * kPi) * std::pow(inner, k - 1.0) : 0.0;
    };
    double acc = 0.0;
    for (int p = 0; p < panels; p++) {
        const double mid = lo + ((double)p + 0.5) * step;
        for (int i = 0; i < 16; i++) {
            const double dx = half * GL32_X[i];
            acc += GL32_W[i] * (f(mid + dx) + f(mid - dx));
        }
    }
    return std::fmin(k * (half * acc), 1.0);
}

inline double studentized_range_cdf(double q, double k, double df) {
    if (q <= 0.0) return 0.0;
    if (!std::isfinite(df) || df > 2.0e4) return range_cdf(q, k);
    const double half_df = 0.5 * df;
    const double ln_c = half_df * std::log(half_df) - ln_gamma(half_df);
    auto g = [&](double u) {
        if (u <= 0.0) return 0.0;
        return std::exp(ln_c + (half_df - 1.0) * std::log(u) - half_df * u) * range_cdf(q * std::sqrt(u), k);
    };
    const double u_max = 1.0 + 12.0 * std::sqrt(2.0 / df);
    const int panels = df < 6.0 ? 96 : 24;
    const double grade = df < 6.0 ? 3.0 : 2.0;
    double acc = 0.0, a = 0.0;
    for (int p = 1; p <= panels; p++) {
        const double b = u_max * std::pow((double)p / (double)panels, grade);
        const double half = 0.5 * (b - a), mid = 0.5 * (b + a);
        for (int i = 0; i < 16; i++) {
            const double dx = half * GL32_X[i];
            acc += half * GL32_W[i] * (g(mid + dx) + g(mid - dx));
        }
        a = b;
    }
    return std::fmin(acc, 1.0);
}

inline double q_tukey_ppf(double p, double k, double df) {
    if (!std::isfinite(k) || k <= 1.0 || !std::isfinite(df) || df < 2.0 || !std::isfinite(p)) return kNaN;
    if (p <= 0.0) return 0.0;
    if (p >= 1.0) return kInf;
    const double t_q = t_ppf(1.0 - (1.0 - p) / (k * (k - 1.0)), df);
    double q = (std::isfinite(t_q) && t_q > 0.0) ? kSqrt2 * t_q : 3.0;
    double lo = 0.0, hi = std::fmax(q, 1.0) * 2.0 + 4.0;
    while (studentized_range_cdf(hi, k, df) < p && hi < 1.0e3) hi *= 1.5;
    const double h = 1e-4;
    for (int it = 0; it < 60; it++) {
        const double f = studentized_range_cdf(q, k, df) - p;
        if (std::fabs(f) < 1e-9) return q;
        (f < 0.0 ? lo : hi) = q;
        const double deriv = (studentized_range_cdf(q + h, k, df) - studentized_range_cdf(q - h, k, df)) / (2.0 * h);
        double next = (std::isfinite(deriv) && deriv > 1e-12) ? q - f / deriv : kNaN;
        if (!std::isfinite(next) || next <= lo || next >= hi) next = 0.5 * (lo + hi);
        if (std::fabs(next - q) < 1e-10) return next;
        q = next;
    }
    return q;
}

// marginals/t3.rs:23-76 — censored t(3) PPF table standardised to unit variance.
inline std::vector<double> build_t3_table() {
    const int res = 2048;
    const double pmin = 0.00121, pmax = 0.99879, denom = (double)(res - 1);
    std::vector<double> v(res);
    for (int i = 0; i < res; i++) v[i] = t_ppf(pmin + (pmax - pmin) * (double)i / denom, 3.0);
    const double du = (pmax - pmin) / denom;
    double m1 = pmin * v[0] + (1.0 - pmax) * v[res - 1];
    double m2 = pmin * v[0] * v[0] + (1.0 - pmax) * v[res - 1] * v[res - 1];
    for (int i = 0; i + 1 < res; i++) {
        const double a = v[i], b = v[i + 1];
        m1 += du * (a + b) / 2.0;
        m2 += du * (a * a + a * b + b * b) / 3.0;
    }
    const double sd = std::sqrt(m2 - m1 * m1);
    for (double& x : v) x /= sd;
    return v;
}

}  // namespace crit
}  // namespace mcp
