// orchestrator.hpp — the callers on either side of the hot path, host-side C++ above the
// kernels: the per-N PowerResult fold of aggregate_batch (engine-orchestrator/src/aggregation.rs:18-271), done from the
// per-cell COUNTERS the kernels already reduce; the cluster-aware sample-size grid (grid.rs:37-147); the first-N
// lookups (grid.rs:162-206, 368-389); and the model-based crossing fit of find_sample_size — monotonicity gate, PAVA,
// Wilson band inversion, probit-in-sqrt(N) hint (fit.rs:45-246).  Pure functions of integers and doubles: no device code.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "critvals.hpp"

namespace mcp {
namespace orch {

struct Ci {
    double lo, hi;
};

// Wilson score 95 % interval, z = 1.96 hard-coded (aggregation.rs:18-33); n == 0 → (0, 1).
inline Ci wilson_ci(uint64_t k, uint64_t n) {
    const double Z = 1.96;
    if (n == 0) return {0.0, 1.0};
    const double nf = (double)n, p = (double)k / nf, z2 = Z * Z;
    const double denom = 1.0 + z2 / nf;
    const double center = (p + z2 / (2.0 * nf)) / denom;
    const double margin = Z * std::sqrt(p * (1.0 - p) / nf + z2 / (4.0 * nf * nf)) / denom;
    return {std::max(center - margin, 0.0), std::min(center + margin, 1.0)};
}

inline double rate(uint64_t k, uint64_t n) { return n == 0 ? 0.0 : (double)k / (double)n; }  // aggregation.rs:38-46

// ------------------------------------------------------------------------------------------ grid (grid.rs:16-147)
inline uint64_t ceil_to(uint64_t x, uint64_t a) {
    a = std::max<uint64_t>(a, 1);
    return (x + a - 1) / a * a;
}
inline uint64_t floor_to(uint64_t x, uint64_t a) {
    a = std::max<uint64_t>(a, 1);
    return x / a * a;
}
inline uint64_t round_to(uint64_t x, uint64_t a) {
    a = std::max<uint64_t>(a, 1);
    return (x + a / 2) / a * a;
}

enum GridError { GRID_OK = 0, GRID_INVALID_ATOM, GRID_INVALID_BOUNDS, GRID_EMPTY, GRID_SINGLE_POINT };

// by_auto: `by` is a point COUNT (ByValue::Auto{count}); otherwise a fixed step (linear) / point count (log).
inline GridError build_grid(uint64_t from_req, uint64_t to_req, uint64_t by, bool by_auto, bool log_mode, uint64_t atom,
                            uint64_t hard_min, std::vector<uint64_t>& grid, std::vector<std::string>& warnings) {
    grid.clear();
    if (atom == 0) return GRID_INVALID_ATOM;
    if (from_req < 1 || to_req < from_req) return GRID_INVALID_BOUNDS;
    const uint64_t snapped_from = std::max(ceil_to(from_req, atom), ceil_to(hard_min, atom));
    const uint64_t snapped_to = floor_to(to_req, atom);
    char buf[256];
    if (snapped_from > from_req) {
        std::snprintf(buf, sizeof buf, "raised `from` from %llu to %llu so each cluster keeps enough rows (atom=%llu)",
                      (unsigned long long)from_req, (unsigned long long)snapped_from, (unsigned long long)atom);
        warnings.push_back(buf);
    }
    if (snapped_to < to_req) {
        std::snprintf(buf, sizeof buf, "lowered `to` from %llu to %llu (nearest multiple of the cluster atom %llu)",
                      (unsigned long long)to_req, (unsigned long long)snapped_to, (unsigned long long)atom);
        warnings.push_back(buf);
    }
    if (snapped_to < snapped_from) return GRID_EMPTY;
    const uint64_t max_feasible = (snapped_to - snapped_from) / atom + 1;
    if (max_feasible < 2) return GRID_SINGLE_POINT;
    if (max_feasible < 4) {
        std::snprintf(buf, sizeof buf, "only %llu grid points after cluster snapping; required-N (and model-based crossing) will be coarse",
                      (unsigned long long)max_feasible);
        warnings.push_back(buf);
    }
    if (!log_mode) {
        uint64_t step;
        if (!by_auto) {
            step = std::max(ceil_to(by, atom), atom);
        } else {
            const uint64_t c = std::min(std::max<uint64_t>(by, 2), max_feasible);
            const uint64_t span = snapped_to - snapped_from, denom = atom * (c - 1);
            step = atom * std::max<uint64_t>((span + denom / 2) / denom, 1);
        }
        for (uint64_t n = snapped_from; n <= snapped_to; n += step) grid.push_back(n);
        if (grid.back() != snapped_to) grid.push_back(snapped_to);  // the `to` endpoint is pinned
    } else {
        const uint64_t c = std::min(std::max<uint64_t>(by, 2), max_feasible);
        const double lo = std::log10((double)snapped_from), hi = std::log10((double)snapped_to);
        for (uint64_t i = 0; i < c; i++) {
            const double t = (double)i / (double)(c - 1);
            const uint64_t v = (uint64_t)std::llround(std::pow(10.0, lo + t * (hi - lo)));
            grid.push_back(std::min(std::max(round_to(v, atom), snapped_from), snapped_to));
        }
        grid.front() = snapped_from;
        grid.back() = snapped_to;
        std::sort(grid.begin(), grid.end());
        grid.erase(std::unique(grid.begin(), grid.end()), grid.end());
    }
    return GRID_OK;
}

inline double as_proportion(double p) { return p > 1.0 ? p / 100.0 : p; }  // grid.rs:151-158

// ------------------------------------------------------------------------------------------ crossing fit (fit.rs)
enum CrossingStatus { FIT_FITTED = 0, FIT_AT_OR_BELOW_MIN = 1, FIT_NOT_REACHED = 2, FIT_NON_MONOTONE = 3 };

struct CrossingFit {
    int status = FIT_NOT_REACHED;
    double n_star = NAN;          // FITTED: interpolated crossing
    uint64_t n_achievable = 0;    // FITTED: ceil(n_star) to the cluster atom
    bool has_ci_lo = false, has_ci_hi = false;
    double ci_lo = NAN, ci_hi = NAN;
    uint64_t n_min = 0;           // AT_OR_BELOW_MIN
    bool has_n_approx = false;    // NOT_REACHED: probit extrapolation hint
    uint64_t n_approx = 0;
    double max_violation = NAN;   // NON_MONOTONE
};

// Largest decrease max_{i<j}(p_i − p_j) against 2·sqrt((p_i(1−p_i) + p_j(1−p_j))/n_sims) for the argmax pair.
inline bool monotonicity_violation(const std::vector<double>& p, double n_sims, double* max_drop_out) {
    double max_drop = 0.0, pi = 0.0, pj = 0.0;
    bool found = false;
    for (size_t i = 0; i < p.size(); i++)
        for (size_t j = i + 1; j < p.size(); j++) {
            const double d = p[i] - p[j];
            if (d > max_drop) {
                max_drop = d;
                pi = p[i];
                pj = p[j];
                found = true;
            }
        }
    if (!found) return false;
    const double se = std::sqrt((pi * (1.0 - pi) + pj * (1.0 - pj)) / n_sims);
    *max_drop_out = max_drop;
    return max_drop > 2.0 * se;
}

// Pool-adjacent-violators, unweighted; only strict decreases pool.
inline std::vector<double> pava(const std::vector<double>& y) {
    std::vector<double> means;
    std::vector<size_t> lens;
    for (double v : y) {
        means.push_back(v);
        lens.push_back(1);
        while (means.size() >= 2 && means[means.size() - 2] > means[means.size() - 1]) {
            const double m1 = means.back();
            const size_t l1 = lens.back();
            means.pop_back();
            lens.pop_back();
            const size_t k = means.size() - 1, l0 = lens[k];
            means[k] = (means[k] * (double)l0 + m1 * (double)l1) / (double)(l0 + l1);
            lens[k] = l0 + l1;
        }
    }
    std::vector<double> out;
    for (size_t b = 0; b < means.size(); b++) out.insert(out.end(), lens[b], means[b]);
    return out;
}

inline bool crossing(const std::vector<uint64_t>& grid, const std::vector<double>& fitted, double target, double* out) {
    if (fitted[0] >= target) {
        *out = (double)grid[0];
        return true;
    }
    for (size_t i = 1; i < fitted.size(); i++)
        if (fitted[i] >= target) {
            const double x0 = (double)grid[i - 1], x1 = (double)grid[i], y0 = fitted[i - 1], y1 = fitted[i];
            *out = x0 + (target - y0) * (x1 - x0) / (y1 - y0);
            return true;
        }
    return false;
}

inline uint64_t ceil_to_atom(double n, uint64_t atom) {  // 1e-9 backoff absorbs interpolation noise (fit.rs:180-188)
    const uint64_t a = std::max<uint64_t>(atom, 1);
    const uint64_t n_int = (uint64_t)std::max(std::ceil(n - 1e-9), 1.0);
    return (n_int + a - 1) / a * a;
}

inline bool probit_hint(const std::vector<uint64_t>& grid, const std::vector<double>& p_hat, const std::vector<double>& fitted,
                        double n_sims, double target, uint64_t atom, uint64_t* out) {
    if (fitted.empty() || fitted.back() < 0.5) return false;  // MIN_HINT_POWER
    const double clamp_lo = 0.5 / n_sims, clamp_hi = 1.0 - 0.5 / n_sims;
    double sw = 0, swx = 0, swy = 0, swxx = 0, swxy = 0;
    for (size_t i = 0; i < grid.size(); i++) {
        const double p = std::min(std::max(p_hat[i], clamp_lo), clamp_hi);
        const double z = crit::norm_ppf(p);
        const double pdf = std::exp(-0.5 * z * z) / std::sqrt(2.0 * M_PI);
        const double w = n_sims * pdf * pdf / (p * (1.0 - p));
        const double x = std::sqrt((double)grid[i]);
        sw += w;
        swx += w * x;
        swy += w * z;
        swxx += w * x * x;
        swxy += w * x * z;
    }
    const double denom = sw * swxx - swx * swx;
    if (!std::isfinite(denom) || denom <= 0.0) return false;
    const double slope = (sw * swxy - swx * swy) / denom;
    if (!std::isfinite(slope) || slope <= 0.0) return false;
    const double intercept = (swy - slope * swx) / sw;
    const double sqrt_n = (crit::norm_ppf(target) - intercept) / slope;
    if (!std::isfinite(sqrt_n) || sqrt_n <= 0.0) return false;
    const double n_cross = sqrt_n * sqrt_n;
    if (n_cross > 2.0 * (double)grid.back()) return false;  // EXTRAPOLATION_CAP
    *out = ceil_to_atom(n_cross, atom);
    return true;
}

inline CrossingFit fit_crossing(const std::vector<uint64_t>& grid, const std::vector<uint64_t>& counts, uint64_t n_sims,
                                double target_power, uint64_t atom) {
    CrossingFit f;
    if (n_sims == 0 || grid.empty()) return f;  // NotReached, no hint
    const double target = as_proportion(target_power), nf = (double)n_sims;
    std::vector<double> p_hat(counts.size());
    for (size_t i = 0; i < counts.size(); i++) p_hat[i] = (double)counts[i] / nf;
    double drop = 0.0;
    if (monotonicity_violation(p_hat, nf, &drop)) {
        f.status = FIT_NON_MONOTONE;
        f.max_violation = drop;
        return f;
    }
    const std::vector<double> fitted = pava(p_hat);
    if (fitted[0] >= target) {
        f.status = FIT_AT_OR_BELOW_MIN;
        f.n_min = grid[0];
        return f;
    }
    if (fitted.back() < target) {
        f.status = FIT_NOT_REACHED;
        f.has_n_approx = probit_hint(grid, p_hat, fitted, nf, target, atom, &f.n_approx);
        return f;
    }
    f.status = FIT_FITTED;
    crossing(grid, fitted, target, &f.n_star);
    f.n_achievable = ceil_to_atom(f.n_star, atom);
    std::vector<double> lo_s(counts.size()), hi_s(counts.size());
    for (size_t i = 0; i < counts.size(); i++) {
        const Ci c = wilson_ci(counts[i], n_sims);
        lo_s[i] = c.lo;
        hi_s[i] = c.hi;
    }
    const std::vector<double> hi_band = pava(hi_s), lo_band = pava(lo_s);
    f.has_ci_lo = hi_band[0] >= target ? false : crossing(grid, hi_band, target, &f.ci_lo);
    f.has_ci_hi = crossing(grid, lo_band, target, &f.ci_hi);
    return f;
}

// ------------------------------------------------------------------------------------------ PowerResult (result.rs:54-144)
struct PosthocPower {
    uint64_t n_levels = 0;
    std::vector<uint64_t> k_unc, k_cor;
};

struct PowerResult {
    uint64_t n = 0, n_sims = 0;
    std::vector<int> target_indices;
    std::vector<std::pair<int, int>> contrast_pairs;
    std::vector<uint64_t> k_unc, k_cor;           // success counts per target (marginals then contrasts)
    uint64_t conv_count = 0, overall_count = 0;
    bool has_overall = false;
    std::vector<uint64_t> hist_unc, hist_cor;     // #significant tests per sim (targets + posthoc contrasts)
    std::vector<PosthocPower> posthoc;
    std::vector<uint64_t> excl_sparse, excl_sep;  // per factor
    int estimator = 0;                            // 0 ols, 1 glm, 2 mle
    uint64_t boundary_hits = 0, joint_count = 0, singular_count = 0;
    int n_components = 0;          // variance components of the model (1 + slopes + extra groupings; result.rs:362)
    uint64_t pin_counts[5] = {0, 0, 0, 0, 0};  // converged fits with component k pinned at 0
    std::vector<uint8_t> boundary_hit;            // per sim (only when the byte tensors were requested)
    std::vector<std::string> grid_warnings;

    double convergence_rate() const { return n_sims == 0 ? 1.0 : (double)conv_count / (double)n_sims; }
};

// ---- JSON (serde field names of PowerResult / SampleSizeResult; NaN → null like serde_json)
inline void jnum(std::string& o, double v) {
    if (!std::isfinite(v)) {
        o += "null";
        return;
    }
    char b[40];
    std::snprintf(b, sizeof b, "%.17g", v);
    o += b;
}
inline void jint(std::string& o, uint64_t v) { o += std::to_string(v); }
inline void jstr(std::string& o, const std::string& s) {
    o += '"';
    for (char c : s) {
        if (c == '"' || c == '\\') o += '\\';
        o += c;
    }
    o += '"';
}
template <typename T, typename F>
inline void jarr(std::string& o, const std::vector<T>& v, F each) {
    o += '[';
    for (size_t i = 0; i < v.size(); i++) {
        if (i) o += ',';
        each(o, v[i]);
    }
    o += ']';
}
inline void jci(std::string& o, const Ci& c) {
    o += "{\"lo\":";
    jnum(o, c.lo);
    o += ",\"hi\":";
    jnum(o, c.hi);
    o += '}';
}
inline void jcounts_power(std::string& o, const char* key, const std::vector<uint64_t>& k, uint64_t n) {
    o += '"';
    o += key;
    o += "\":";
    jarr(o, k, [n](std::string& s, uint64_t v) { jnum(s, rate(v, n)); });
}
inline void jcounts_ci(std::string& o, const char* key, const std::vector<uint64_t>& k, uint64_t n) {
    o += '"';
    o += key;
    o += "\":";
    jarr(o, k, [n](std::string& s, uint64_t v) { jci(s, wilson_ci(v, n)); });
}
inline void jcounts(std::string& o, const char* key, const std::vector<uint64_t>& k) {
    o += '"';
    o += key;
    o += "\":";
    jarr(o, k, [](std::string& s, uint64_t v) { jint(s, v); });
}

inline void power_result_json(std::string& o, const PowerResult& r) {
    o += "{\"n\":";
    jint(o, r.n);
    o += ",\"n_sims\":";
    jint(o, r.n_sims);
    o += ",\"target_indices\":";
    jarr(o, r.target_indices, [](std::string& s, int v) { s += std::to_string(v); });
    o += ",\"contrast_pairs\":";
    jarr(o, r.contrast_pairs, [](std::string& s, const std::pair<int, int>& p) {
        s += '[' + std::to_string(p.first) + ',' + std::to_string(p.second) + ']';
    });
    o += ',';
    jcounts_power(o, "power_uncorrected", r.k_unc, r.n_sims);
    o += ',';
    jcounts_power(o, "power_corrected", r.k_cor, r.n_sims);
    o += ',';
    jcounts_ci(o, "ci_uncorrected", r.k_unc, r.n_sims);
    o += ',';
    jcounts_ci(o, "ci_corrected", r.k_cor, r.n_sims);
    o += ",\"convergence_rate\":";
    jnum(o, r.convergence_rate());
    o += ",\"boundary_hit\":";
    jarr(o, r.boundary_hit, [](std::string& s, uint8_t v) { s += std::to_string((int)v); });
    o += ",\"estimator_extras\":";
    if (r.estimator == 0) {
        o += "{\"estimator\":\"ols\"}";
    } else if (r.estimator == 1) {  // EstimatorExtras::Glm (result.rs:297-333): no GLMM on this backend → nothing singular
        o += "{\"estimator\":\"glm\",\"baseline_prob_realized\":null,\"baseline_prob_sum\":0.0,\"baseline_prob_n\":0,"
             "\"singular_fit_rate\":0.0,\"singular_count\":0,\"singular_n\":";
        jint(o, r.conv_count);
        o += ",\"tau_squared_hat_mean\":null,\"tau_squared_hat_sum\":0.0,\"tau_squared_hat_n\":0}";
    } else {  // EstimatorExtras::Mle (result.rs:334-403), one variance component (random intercept)
        o += "{\"estimator\":\"mle\",\"tau_estimate\":null,\"boundary_hits\":";
        jint(o, r.boundary_hits);
        o += ",\"joint_uncorrected_rate\":";
        jnum(o, rate(r.joint_count, r.n_sims));
        o += ",\"joint_corrected_rate\":";
        jnum(o, rate(r.joint_count, r.n_sims));
        o += ",\"tau_sum\":0.0,\"tau_n\":0,\"joint_uncorrected_count\":";
        jint(o, r.joint_count);
        o += ",\"joint_corrected_count\":";
        jint(o, r.joint_count);
        o += ",\"singular_fit_rate\":";
        jnum(o, r.conv_count ? (double)r.singular_count / (double)r.conv_count : 0.0);
        o += ",\"singular_count\":";
        jint(o, r.singular_count);
        o += ",\"singular_n\":";
        jint(o, r.conv_count);
        o += ",\"boundary_rate_per_component\":[";
        for (int k = 0; k < r.n_components; k++) {
            if (k) o += ',';
            jnum(o, r.conv_count ? (double)r.pin_counts[k] / (double)r.conv_count : 0.0);
        }
        o += "],\"boundary_component_counts\":[";
        for (int k = 0; k < r.n_components; k++) {
            if (k) o += ',';
            jint(o, r.pin_counts[k]);
        }
        o += "]}";
    }
    o += ",\"overall_significant_rate\":";
    if (r.has_overall) jnum(o, rate(r.overall_count, r.n_sims));
    else o += "null";
    o += ',';
    jcounts(o, "success_counts_uncorrected", r.k_unc);
    o += ',';
    jcounts(o, "success_counts_corrected", r.k_cor);
    o += ",\"convergence_count\":";
    jint(o, r.conv_count);
    o += ",\"overall_significant_count\":";
    jint(o, r.overall_count);
    o += ",\"overall_significant_ci\":";
    if (r.has_overall) jci(o, wilson_ci(r.overall_count, r.n_sims));
    else o += "null";
    o += ',';
    jcounts(o, "success_count_histogram_uncorrected", r.hist_unc);
    o += ',';
    jcounts(o, "success_count_histogram_corrected", r.hist_cor);
    o += ",\"grid_warnings\":";
    jarr(o, r.grid_warnings, [](std::string& s, const std::string& w) { jstr(s, w); });
    o += ",\"posthoc\":";
    const uint64_t ns = r.n_sims;
    jarr(o, r.posthoc, [ns](std::string& s, const PosthocPower& b) {
        s += "{\"n_levels\":";
        jint(s, b.n_levels);
        s += ',';
        jcounts_power(s, "power_uncorrected", b.k_unc, ns);
        s += ',';
        jcounts_power(s, "power_corrected", b.k_cor, ns);
        s += ',';
        jcounts_ci(s, "ci_uncorrected", b.k_unc, ns);
        s += ',';
        jcounts_ci(s, "ci_corrected", b.k_cor, ns);
        s += ',';
        jcounts(s, "success_counts_uncorrected", b.k_unc);
        s += ',';
        jcounts(s, "success_counts_corrected", b.k_cor);
        s += '}';
    });
    o += ',';
    jcounts(o, "factor_exclusion_counts", r.excl_sparse);
    o += ',';
    jcounts(o, "factor_separation_counts", r.excl_sep);
    o += '}';
}

inline void crossing_fit_json(std::string& o, const CrossingFit& f) {  // #[serde(tag = "status", rename_all = "snake_case")]
    switch (f.status) {
        case FIT_FITTED:
            o += "{\"status\":\"fitted\",\"n_star\":";
            jnum(o, f.n_star);
            o += ",\"n_achievable\":";
            jint(o, f.n_achievable);
            o += ",\"ci_lo\":";
            if (f.has_ci_lo) jnum(o, f.ci_lo);
            else o += "null";
            o += ",\"ci_hi\":";
            if (f.has_ci_hi) jnum(o, f.ci_hi);
            else o += "null";
            o += '}';
            break;
        case FIT_AT_OR_BELOW_MIN:
            o += "{\"status\":\"at_or_below_min\",\"n_min\":";
            jint(o, f.n_min);
            o += '}';
            break;
        case FIT_NOT_REACHED:
            o += "{\"status\":\"not_reached\",\"n_approx\":";
            if (f.has_n_approx) jint(o, f.n_approx);
            else o += "null";
            o += '}';
            break;
        default:
            o += "{\"status\":\"non_monotone\",\"max_violation\":";
            jnum(o, f.max_violation);
            o += '}';
    }
}

// ------------------------------------------------------------------------------------------ sample-size derivations
struct SampleSizeDerivations {
    std::vector<int64_t> first_achieved, first_joint_achieved;  // −1 = never reached on the grid
    std::vector<CrossingFit> fitted, fitted_joint;
    bool has_overall = false;
    int64_t first_overall_achieved = -1;
    CrossingFit fitted_overall;
};

// derive_sample_size_outputs (grid.rs:226-330): corrected power per target, corrected joint histograms (P(>=k) tail
// sums; the joint n_sims is the histogram row sum), and the overall test when the grid carries it.
inline SampleSizeDerivations derive_sample_size_outputs(const std::vector<PowerResult>& per_n, const std::vector<uint64_t>& grid,
                                                        double target_power, uint64_t atom) {
    SampleSizeDerivations d;
    if (per_n.empty()) return d;
    const double p_target = as_proportion(target_power);
    const uint64_t n_sims = per_n[0].n_sims;
    const size_t T = per_n[0].k_cor.size();
    for (size_t t = 0; t < T; t++) {
        int64_t first = -1;
        std::vector<uint64_t> counts(per_n.size());
        for (size_t i = 0; i < per_n.size(); i++) {
            counts[i] = per_n[i].k_cor[t];
            if (first < 0 && rate(per_n[i].k_cor[t], per_n[i].n_sims) >= p_target) first = (int64_t)grid[i];
        }
        d.first_achieved.push_back(first);
        d.fitted.push_back(fit_crossing(grid, counts, n_sims, target_power, atom));
    }
    uint64_t joint_n = 0;
    for (uint64_t v : per_n[0].hist_cor) joint_n += v;
    const size_t n_joint = per_n[0].hist_cor.empty() ? 0 : per_n[0].hist_cor.size() - 1;
    for (size_t k = 1; k <= n_joint; k++) {
        int64_t first = -1;
        std::vector<uint64_t> counts(per_n.size());
        for (size_t i = 0; i < per_n.size(); i++) {
            uint64_t ge = 0;
            for (size_t j = k; j < per_n[i].hist_cor.size(); j++) ge += per_n[i].hist_cor[j];
            counts[i] = ge;
            if (first < 0 && joint_n > 0 && (double)ge / (double)joint_n >= p_target) first = (int64_t)grid[i];
        }
        d.first_joint_achieved.push_back(first);
        d.fitted_joint.push_back(fit_crossing(grid, counts, joint_n, target_power, atom));
    }
    d.has_overall = per_n[0].has_overall;
    if (d.has_overall) {
        std::vector<uint64_t> counts(per_n.size());
        for (size_t i = 0; i < per_n.size(); i++) {
            counts[i] = per_n[i].overall_count;
            if (d.first_overall_achieved < 0 && rate(per_n[i].overall_count, per_n[i].n_sims) >= p_target)
                d.first_overall_achieved = (int64_t)grid[i];
        }
        d.fitted_overall = fit_crossing(grid, counts, n_sims, target_power, atom);
    }
    return d;
}

inline void jopt(std::string& o, int64_t v) {
    if (v < 0) o += "null";
    else o += std::to_string(v);
}

}  // namespace orch
}  // namespace mcp
