// spec.hpp — `SimulationSpec` (crates/engine-core/src/spec.rs:250-417) parsed from its serde JSON
// shape, plus the spec-level validation run_batch performs (batch.rs:208-330).
#pragma once
#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>

#include "json.hpp"
#include "plan.h"

namespace mcp {

// Error classes mirror EngineError (spec.rs:19-29); codes are the C-ABI return values.
enum McpStatus { MCP_OK = 0, MCP_INVALID_SPEC = 1, MCP_CORRELATION_NOT_PSD = 2, MCP_RANK_DEFICIENT = 3, MCP_CANCELLED = 4,
                 MCP_CUDA_ERROR = 5, MCP_UNSUPPORTED = 6 };

struct EngineError : std::runtime_error {
    int code;
    EngineError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
[[noreturn]] inline void invalid(const std::string& m) { throw EngineError(MCP_INVALID_SPEC, "invalid spec: " + m); }
[[noreturn]] inline void unsupported(const std::string& m) { throw EngineError(MCP_UNSUPPORTED, "unsupported by the CUDA backend: " + m); }

struct Spec {
    int k = 0, n_fd = 0;
    std::vector<double> correlation;
    std::vector<int> var_types;
    std::vector<uint8_t> var_pinned;
    std::vector<double> var_params;
    std::vector<int> factor_n_levels;
    std::vector<double> factor_proportions;
    std::vector<int> factor_sampled;  // -1 inherit
    std::vector<double> effect_sizes;
    std::vector<int> target_indices;
    std::vector<int> fit_columns;
    std::vector<std::pair<int, int>> contrast_pairs;
    std::vector<std::vector<int>> interactions;
    int correction = C_NONE;
    double alpha = 0.05;
    double posthoc_alpha = NAN;
    int hsk_driver = -1;
    int residual_dist = R_NORMAL;
    bool residual_pinned = false;
    int outcome = O_CONTINUOUS;
    bool link_probit = false;
    int estimator = E_OLS;
    bool report_overall = false;
    int min_level_count = 0;
    bool has_posthoc = false;
    // posthoc blocks (spec.rs:90-103): per factor the target (dummy) columns; the pairwise contrasts of all blocks,
    // flat in block order (build_posthoc_contrasts, batch.rs:2183-2227): c = e_pos − e_neg, neg < 0 → reference level
    std::vector<std::vector<int>> posthoc_blocks;
    std::vector<std::pair<int, int>> posthoc_contrasts;
    std::vector<int> posthoc_block_levels, posthoc_block_ncontrasts;
    // uploaded-data arms (spec.rs upload_* fields; contract_adapter.rs:359-455)
    std::vector<double> upload_normal, upload_data;
    int un_rows = 0, un_cols = 0, ud_rows = 0, ud_cols = 0;
    std::vector<int> frame_map;  // bootstrap_frame_map, −1 = None; empty = not bootstrap
    bool has_slopes_or_groupings = false;  // what the CUDA path does not run: > 1 slope, slope-slope correlations, extra groupings
    int n_slopes = 0, slope_col[2] = {-1, -1};  // random slopes on the primary grouping (two at most)
    double slope_variance[2] = {0.0, 0.0}, slope_corr[2] = {0.0, 0.0}, slope_corr12 = 0.0;
    int n_extra = 0, extra_nested[2] = {0, 0}, extra_count[2] = {0, 0};  // intercept-only extra groupings (two at most)
    double extra_tau_squared[2] = {0.0, 0.0};
    std::vector<int> between_cols;  // between_var_indices (spec.rs:279): design columns broadcast within a cluster
    // cluster
    bool has_cluster = false;
    bool sizing_fixed_size = false;
    int sizing_value = 0;
    double tau_squared = 0.0;
    // scenario
    double heterogeneity = 0.0, hsk_ratio = 1.0, corr_noise_sd = 0.0, dist_change_prob = 0.0, resid_change_prob = 0.0;
    std::vector<int> new_dists, resid_dists;
    double residual_df = 0.0;
    bool sampled_factor_proportions = false;
    bool has_lme = false;
    int re_dist = R_NORMAL;
    double re_df = 0.0, icc_noise_sd = 0.0;

    int P() const { return 1 + k + n_fd + (int)interactions.size(); }
    int T() const { return (int)target_indices.size() + (int)contrast_pairs.size(); }
    bool is_optimistic() const {  // spec.rs:229-236
        return heterogeneity == 0.0 && hsk_ratio == 1.0 && corr_noise_sd == 0.0 && dist_change_prob == 0.0 &&
               resid_change_prob == 0.0 && !has_lme;
    }
};

namespace detail {
inline const JsonValue& req(const JsonValue& o, const char* key) {
    const JsonValue* v = o.get(key);
    if (!v) invalid(std::string("missing field `") + key + "`");
    return *v;
}
inline double num(const JsonValue& v, const char* what) {
    if (v.kind != JsonValue::Number) invalid(std::string("field `") + what + "` must be a number");
    return v.num;
}
inline int to_int(const JsonValue& v, const char* what) {
    const double d = num(v, what);
    if (d != std::floor(d) || std::fabs(d) > 2147483647.0) invalid(std::string("field `") + what + "` must be an integer");
    return (int)d;
}
inline bool to_bool(const JsonValue& v, const char* what) {
    if (v.kind != JsonValue::Bool) invalid(std::string("field `") + what + "` must be a bool");
    return v.b;
}
inline const std::vector<JsonValue>& arr(const JsonValue& v, const char* what) {
    if (v.kind != JsonValue::Array) invalid(std::string("field `") + what + "` must be an array");
    return v.arr;
}
inline int enum_of(const JsonValue& v, const char* what, std::initializer_list<std::pair<const char*, int>> table) {
    if (v.kind != JsonValue::String) invalid(std::string("field `") + what + "` must be a string");
    for (auto& e : table)
        if (v.str == e.first) return e.second;
    invalid(std::string("unknown variant `") + v.str + "` for `" + what + "`");
}
inline int dist_of(const JsonValue& v, const char* what) {
    // engine-contract/src/distribution.rs:15-26 (snake_case)
    const int d = enum_of(v, what, {{"normal", D_NORMAL}, {"binary", D_BINARY}, {"right_skewed", D_RIGHT_SKEWED},
                                    {"left_skewed", D_LEFT_SKEWED}, {"high_kurtosis", D_HIGH_KURTOSIS}, {"uniform", D_UNIFORM},
                                    {"uploaded_factor", D_UPLOADED_FACTOR}, {"uploaded_binary", D_UPLOADED_BINARY},
                                    {"uploaded_data", D_UPLOADED_DATA}});
    return d;
}
inline int resid_of(const JsonValue& v, const char* what) {
    // engine-contract/src/outcome.rs:67-75
    return enum_of(v, what, {{"normal", R_NORMAL}, {"right_skewed", R_RIGHT_SKEWED}, {"left_skewed", R_LEFT_SKEWED},
                             {"high_kurtosis", R_HIGH_KURTOSIS}, {"uniform", R_UNIFORM}});
}
}  // namespace detail

inline Spec parse_spec(const char* json, size_t len) {
    using namespace detail;
    JsonValue root;
    try {
        root = JsonParser(json, len).parse();
    } catch (const std::runtime_error& e) {
        invalid(e.what());
    }
    if (root.kind != JsonValue::Object) invalid("spec must be a JSON object");
    Spec s;
    s.k = to_int(req(root, "n_non_factor"), "n_non_factor");
    s.n_fd = to_int(req(root, "n_factor_dummies"), "n_factor_dummies");
    if (s.k < 0 || s.n_fd < 0) invalid("negative column count");
    for (auto& v : arr(req(root, "correlation"), "correlation")) s.correlation.push_back(num(v, "correlation"));
    for (auto& v : arr(req(root, "var_types"), "var_types")) s.var_types.push_back(dist_of(v, "var_types"));
    if (auto* p = root.get("var_pinned"))
        if (!p->is_null())
            for (auto& v : arr(*p, "var_pinned")) s.var_pinned.push_back(to_bool(v, "var_pinned"));
    if (auto* p = root.get("var_params"))
        if (!p->is_null())
            for (auto& v : arr(*p, "var_params")) s.var_params.push_back(num(v, "var_params"));
    auto nonempty = [&](const char* key) {
        const JsonValue* p = root.get(key);
        return p && p->kind == JsonValue::Array && !p->arr.empty();
    };
    auto table = [&](const char* key, const char* shape_key, std::vector<double>& out, int& rows, int& cols) {
        if (nonempty(key))
            for (auto& v : arr(*root.get(key), key)) out.push_back(num(v, key));
        if (const JsonValue* p = root.get(shape_key))
            if (!p->is_null()) {
                auto& sh = arr(*p, shape_key);
                if (sh.size() != 2) invalid(std::string("`") + shape_key + "` must be a (rows, cols) pair");
                rows = to_int(sh[0], shape_key);
                cols = to_int(sh[1], shape_key);
            }
        if (rows < 0 || cols < 0 || (size_t)rows * (size_t)cols != out.size())
            invalid(std::string("`") + key + "` length does not match `" + shape_key + "`");
    };
    table("upload_normal", "upload_normal_shape", s.upload_normal, s.un_rows, s.un_cols);
    table("upload_data", "upload_data_shape", s.upload_data, s.ud_rows, s.ud_cols);
    if (nonempty("bootstrap_frame_map"))
        for (auto& v : arr(*root.get("bootstrap_frame_map"), "bootstrap_frame_map"))
            s.frame_map.push_back(v.is_null() ? -1 : to_int(v, "bootstrap_frame_map"));
    if (nonempty("between_var_indices"))
        for (auto& v : arr(*root.get("between_var_indices"), "between_var_indices")) s.between_cols.push_back(to_int(v, "between_var_indices"));
    for (auto& v : arr(req(root, "factor_n_levels"), "factor_n_levels")) s.factor_n_levels.push_back(to_int(v, "factor_n_levels"));
    for (auto& v : arr(req(root, "factor_proportions"), "factor_proportions")) s.factor_proportions.push_back(num(v, "factor_proportions"));
    if (auto* p = root.get("factor_sampled"))
        if (!p->is_null())
            for (auto& v : arr(*p, "factor_sampled")) s.factor_sampled.push_back(v.is_null() ? -1 : (to_bool(v, "factor_sampled") ? 1 : 0));
    for (auto& v : arr(req(root, "effect_sizes"), "effect_sizes")) s.effect_sizes.push_back(num(v, "effect_sizes"));
    for (auto& v : arr(req(root, "target_indices"), "target_indices")) s.target_indices.push_back(to_int(v, "target_indices"));
    if (auto* p = root.get("fit_columns"))
        if (!p->is_null())
            for (auto& v : arr(*p, "fit_columns")) s.fit_columns.push_back(to_int(v, "fit_columns"));
    if (auto* p = root.get("contrast_pairs"))
        if (!p->is_null())
            for (auto& v : arr(*p, "contrast_pairs")) {
                auto& pr = arr(v, "contrast_pairs");
                if (pr.size() != 2) invalid("contrast_pairs entries must be (pos, neg) pairs");
                s.contrast_pairs.push_back({to_int(pr[0], "contrast_pairs"), to_int(pr[1], "contrast_pairs")});
            }
    if (auto* p = root.get("interactions"))
        if (!p->is_null())
            for (auto& v : arr(*p, "interactions")) {
                std::vector<int> comp;
                for (auto& c : arr(v, "interactions")) comp.push_back(to_int(c, "interactions"));
                s.interactions.push_back(comp);
            }
    // engine-contract/src/test_spec.rs:38-47
    s.correction = enum_of(req(root, "correction_method"), "correction_method",
                           {{"none", C_NONE}, {"bonferroni", C_BONFERRONI}, {"holm", C_HOLM}, {"benjamini_hochberg", C_BH}, {"tukey_hsd", C_TUKEY}});
    {
        const JsonValue& cv = req(root, "crit_values");
        s.alpha = num(req(cv, "alpha"), "crit_values.alpha");
        if (auto* pa = cv.get("posthoc_alpha"))
            if (!pa->is_null()) s.posthoc_alpha = num(*pa, "crit_values.posthoc_alpha");
    }
    if (auto* p = root.get("heteroskedasticity_driver"))
        if (!p->is_null()) s.hsk_driver = to_int(*p, "heteroskedasticity_driver");
    s.residual_dist = resid_of(req(root, "residual_dist"), "residual_dist");
    if (auto* p = root.get("residual_pinned")) s.residual_pinned = !p->is_null() && to_bool(*p, "residual_pinned");
    s.outcome = enum_of(req(root, "outcome_kind"), "outcome_kind", {{"continuous", O_CONTINUOUS}, {"binary", O_BINARY}, {"count", O_COUNT}});
    if (auto* p = root.get("link"))
        if (!p->is_null()) s.link_probit = enum_of(*p, "link", {{"probit", 1}}) == 1;
    s.estimator = enum_of(req(root, "estimator"), "estimator", {{"ols", E_OLS}, {"glm", E_GLM}, {"mle", E_MLE}});
    if (auto* p = root.get("posthoc")) {
        s.has_posthoc = p->kind == JsonValue::Array && !p->arr.empty();
        if (s.has_posthoc)
            for (auto& blk : arr(*p, "posthoc")) {
                std::vector<int> tis;
                for (auto& v : arr(req(blk, "target_indices"), "posthoc.target_indices")) tis.push_back(to_int(v, "posthoc.target_indices"));
                s.posthoc_blocks.push_back(tis);
                // batch.rs:2196-2225: canonical order a < b over the levels, level 0 = the reference level
                const int kl = (int)tis.size() + 1;
                for (int a = 0; a < kl; a++)
                    for (int b = a + 1; b < kl; b++) s.posthoc_contrasts.push_back({tis[b - 1], a > 0 ? tis[a - 1] : -1});
                s.posthoc_block_levels.push_back(kl);
                s.posthoc_block_ncontrasts.push_back(kl * (kl - 1) / 2);
            }
    }
    if (auto* p = root.get("report_overall")) s.report_overall = !p->is_null() && to_bool(*p, "report_overall");
    if (auto* p = root.get("factor_min_level_count"))
        if (!p->is_null()) s.min_level_count = to_int(*p, "factor_min_level_count");
    std::vector<int> slope_cols;
    if (auto* p = root.get("cluster_slope_design_cols"))
        if (p->kind == JsonValue::Array)
            for (auto& v : p->arr) slope_cols.push_back(to_int(v, "cluster_slope_design_cols"));
    if (auto* p = root.get("extra_slope_cols"))
        if (p->kind == JsonValue::Array)
            for (auto& v : p->arr) s.has_slopes_or_groupings |= v.kind == JsonValue::Array && !v.arr.empty();
    if (auto* p = root.get("cluster")) {
        if (!p->is_null()) {
            s.has_cluster = true;
            const JsonValue& sz = req(*p, "sizing");
            if (auto* fc = sz.get("FixedClusters")) {
                s.sizing_fixed_size = false;
                s.sizing_value = to_int(req(*fc, "n_clusters"), "n_clusters");
            } else if (auto* fs = sz.get("FixedSize")) {
                s.sizing_fixed_size = true;
                s.sizing_value = to_int(req(*fs, "cluster_size"), "cluster_size");
            } else {
                invalid("cluster.sizing must be FixedClusters or FixedSize");
            }
            s.tau_squared = num(req(*p, "tau_squared"), "tau_squared");
            if (auto* q = p->get("extra_groupings"))
                if (q->kind == JsonValue::Array && !q->arr.empty()) {
                    if (q->arr.size() > 2) s.has_slopes_or_groupings = true;
                    s.n_extra = (int)std::min<size_t>(q->arr.size(), 2);
                    for (int g = 0; g < s.n_extra; g++) {
                        const JsonValue& eg = q->arr[g];
                        const JsonValue& rel = req(eg, "relation");
                        if (auto* cr = rel.get("Crossed")) {
                            s.extra_nested[g] = 0;
                            s.extra_count[g] = to_int(req(*cr, "n_clusters"), "Crossed.n_clusters");
                        } else if (auto* ne = rel.get("NestedWithin")) {
                            s.extra_nested[g] = 1;
                            s.extra_count[g] = to_int(req(*ne, "n_per_parent"), "NestedWithin.n_per_parent");
                        } else {
                            invalid("extra_groupings.relation must be Crossed or NestedWithin");
                        }
                        s.extra_tau_squared[g] = num(req(eg, "tau_squared"), "extra_groupings.tau_squared");
                        if (auto* sl = eg.get("slopes")) s.has_slopes_or_groupings |= sl->kind == JsonValue::Array && !sl->arr.empty();
                    }
                }
            if (auto* q = p->get("slopes"))
                if (q->kind == JsonValue::Array && !q->arr.empty()) {
                    if (q->arr.size() > 2) s.has_slopes_or_groupings = true;
                    if (slope_cols.size() != q->arr.size()) invalid("cluster_slope_design_cols must name one design column per random slope");
                    s.n_slopes = (int)std::min<size_t>(q->arr.size(), 2);
                    for (int a = 0; a < s.n_slopes; a++) {
                        const JsonValue& sl = q->arr[a];
                        s.slope_variance[a] = num(req(sl, "variance"), "slopes.variance");
                        s.slope_corr[a] = num(req(sl, "corr_with_intercept"), "slopes.corr_with_intercept");
                        s.slope_col[a] = slope_cols[a];
                        if (a == 1)
                            if (auto* cw = sl.get("corr_with"))
                                if (cw->kind == JsonValue::Array && !cw->arr.empty()) s.slope_corr12 = num(cw->arr[0], "slopes.corr_with");
                    }
                }
        }
    }
    {
        const JsonValue& sc = req(root, "scenario");
        s.heterogeneity = num(req(sc, "heterogeneity"), "heterogeneity");
        s.hsk_ratio = num(req(sc, "heteroskedasticity_ratio"), "heteroskedasticity_ratio");
        s.corr_noise_sd = num(req(sc, "correlation_noise_sd"), "correlation_noise_sd");
        s.dist_change_prob = num(req(sc, "distribution_change_prob"), "distribution_change_prob");
        for (auto& v : arr(req(sc, "new_distributions"), "new_distributions")) s.new_dists.push_back(dist_of(v, "new_distributions"));
        s.resid_change_prob = num(req(sc, "residual_change_prob"), "residual_change_prob");
        for (auto& v : arr(req(sc, "residual_dists"), "residual_dists")) s.resid_dists.push_back(resid_of(v, "residual_dists"));
        s.residual_df = num(req(sc, "residual_df"), "residual_df");
        if (auto* p = sc.get("sampled_factor_proportions")) s.sampled_factor_proportions = !p->is_null() && to_bool(*p, "sampled_factor_proportions");
        if (auto* p = sc.get("lme"))
            if (!p->is_null()) {
                s.has_lme = true;
                s.re_dist = resid_of(req(*p, "random_effect_dist"), "random_effect_dist");
                s.re_df = num(req(*p, "random_effect_df"), "random_effect_df");
                s.icc_noise_sd = num(req(*p, "icc_noise_sd"), "icc_noise_sd");
            }
    }
    return s;
}

// Lower Cholesky of the k×k correlation (column-major in, row-major L out); left-looking column order
// of data_gen.rs:55-75 / correlation.rs:48-74.  Returns false on a non-positive pivot.
inline bool correlation_llt(const std::vector<double>& corr_colmajor, int k, std::vector<double>& L_rowmajor) {
    std::vector<double> a(corr_colmajor);
    for (int j = 0; j < k; j++) {
        double diag = a[j * k + j];
        for (int m = 0; m < j; m++) diag -= a[m * k + j] * a[m * k + j];
        if (!(diag > 0.0) || !std::isfinite(diag)) return false;
        const double ljj = std::sqrt(diag);
        a[j * k + j] = ljj;
        for (int i = j + 1; i < k; i++) {
            double s = a[j * k + i];
            for (int m = 0; m < j; m++) s -= a[m * k + i] * a[m * k + j];
            a[j * k + i] = s / ljj;
        }
    }
    L_rowmajor.assign((size_t)k * k, 0.0);
    for (int j = 0; j < k; j++)
        for (int i = j; i < k; i++) L_rowmajor[(size_t)i * k + j] = a[j * k + i];
    return true;
}

// batch.rs:208-330 — the spec-level checks run_batch owns (contract-level invariants ran upstream).
inline void validate_spec(const Spec& s, const uint32_t* sample_sizes, uint32_t n_ss) {
    if (s.estimator == E_MLE && !s.has_cluster) invalid("estimator Mle requires ClusterSpec");
    if (s.has_posthoc && s.estimator != E_OLS) invalid("posthoc tests are only supported for estimator=Ols");
    if (sample_sizes) {
        if (n_ss == 0) invalid("sample_sizes must be non-empty");
        for (uint32_t i = 1; i < n_ss; i++)
            if (sample_sizes[i] <= sample_sizes[i - 1])
                invalid("sample_sizes must be strictly increasing; got " + std::to_string(sample_sizes[i - 1]) + " then " +
                        std::to_string(sample_sizes[i]));
        if (sample_sizes[0] == 0) invalid("sample_sizes[0] must be > 0");
    }
    const int P = s.P();
    if ((int)s.correlation.size() != s.k * s.k)
        invalid("correlation length " + std::to_string(s.correlation.size()) + " does not match n_non_factor² = " + std::to_string(s.k * s.k));
    if ((int)s.var_types.size() != s.k)
        invalid("var_types length " + std::to_string(s.var_types.size()) + " does not match n_non_factor " + std::to_string(s.k));
    if ((int)s.effect_sizes.size() != P)
        invalid("effect_sizes length " + std::to_string(s.effect_sizes.size()) + " does not match n_predictors " + std::to_string(P));
    int expected = 0, dummies = 0;
    for (int lv : s.factor_n_levels) {
        if (lv < 2) invalid("factor_n_levels entries must be >= 2; got " + std::to_string(lv));
        expected += lv;
        dummies += lv - 1;
    }
    if ((int)s.factor_proportions.size() != expected)
        invalid("factor_proportions length " + std::to_string(s.factor_proportions.size()) + " does not match Σfactor_n_levels " +
                std::to_string(expected));
    if (dummies != s.n_fd) invalid("n_factor_dummies does not match Σ(factor_n_levels − 1)");
    for (int ti : s.target_indices)
        if (ti < 0 || ti >= P) invalid("target_indices entry " + std::to_string(ti) + " out of range (n_predictors = " + std::to_string(P) + ")");
    for (auto& pr : s.contrast_pairs) {
        if (pr.first < 0 || pr.first >= P) invalid("contrast_pairs positive " + std::to_string(pr.first) + " out of range");
        if (pr.second < 0 || pr.second >= P) invalid("contrast_pairs negative " + std::to_string(pr.second) + " out of range");
    }
    for (auto& comp : s.interactions)
        for (int c : comp)
            if (c < 1 || c >= 1 + s.k + s.n_fd) invalid("interaction component out of range");
    for (auto& dummies : s.posthoc_blocks)
        for (int ti : dummies)
            if (ti < 0 || ti >= P) invalid("posthoc target index " + std::to_string(ti) + " out of range (n_predictors = " + std::to_string(P) + ")");
    if (s.n_slopes > 0) {
        for (int a = 0; a < s.n_slopes; a++) {
            if (s.slope_col[a] < 1 || s.slope_col[a] >= P) invalid("cluster_slope_design_cols entry " + std::to_string(s.slope_col[a]) + " out of range");
            if (!(s.slope_variance[a] >= 0.0) || !(std::fabs(s.slope_corr[a]) <= 1.0)) invalid("random slope variance / correlation out of range");
        }
        if (!(std::fabs(s.slope_corr12) <= 1.0)) invalid("random slope variance / correlation out of range");
        if (!(s.tau_squared > 0.0)) invalid("random slopes need tau_squared > 0");  // engine-contract validate(): z0 = u0 / tau0
    }
    if (s.n_extra > 0) {  // engine-contract validate.rs:720-745
        int n_nested = 0;
        for (int g = 0; g < s.n_extra; g++) {
            if (s.extra_count[g] < 1) invalid("extra grouping needs at least one level");
            if (!(s.extra_tau_squared[g] >= 0.0)) invalid("extra grouping tau_squared must be >= 0");
            if (!s.extra_nested[g] && s.sizing_fixed_size) invalid("GroupingRelation::Crossed requires ClusterSizing::FixedClusters");
            if (s.extra_nested[g]) {
                n_nested++;
                if (s.sizing_fixed_size && s.sizing_value % s.extra_count[g] != 0) invalid("NestedWithin requires cluster_size % n_per_parent == 0");
            }
        }
        if (n_nested > 1) invalid("at most one NestedWithin grouping is supported by this engine version");
    }
    for (int c : s.between_cols)  // the reference indexes x_full unchecked (data_gen.rs:631-633); reject instead
        if (c < 1 || c >= 1 + s.k + s.n_fd) invalid("between_var_indices entry " + std::to_string(c) + " out of range");
    // uploaded-data tables: the reference indexes them unchecked (a malformed spec panics → RuntimeError); reject instead
    if (!s.frame_map.empty()) {
        if ((int)s.frame_map.size() != s.k + (int)s.factor_n_levels.size())
            invalid("bootstrap_frame_map length must equal n_non_factor + n_factors");  // data_gen.rs:117-121
        if (s.ud_rows <= 0 || s.ud_cols <= 0) invalid("bootstrap mode requires a non-empty frame");  // data_gen.rs:563
        for (int fc : s.frame_map)
            if (fc >= s.ud_cols) invalid("bootstrap_frame_map entry " + std::to_string(fc) + " out of range");
    }
    for (int j = 0; j < s.k; j++)
        if (s.var_types[j] == D_UPLOADED_DATA && s.un_rows > 0 && j >= s.un_cols)
            invalid("upload_normal has no column for uploaded_data variable " + std::to_string(j));
    if (s.k > 0) {
        std::vector<double> L;
        if (!correlation_llt(s.correlation, s.k, L)) throw EngineError(MCP_CORRELATION_NOT_PSD, "correlation matrix is not positive semi-definite");
    }
    if (!std::isfinite(s.alpha) || s.alpha <= 0.0 || s.alpha >= 1.0) invalid("alpha must lie in (0, 1); got " + std::to_string(s.alpha));
    if (!std::isnan(s.posthoc_alpha) && (!std::isfinite(s.posthoc_alpha) || s.posthoc_alpha <= 0.0 || s.posthoc_alpha >= 1.0))
        invalid("posthoc_alpha must lie in (0, 1)");
}

}  // namespace mcp
