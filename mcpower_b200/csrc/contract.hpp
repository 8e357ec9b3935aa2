// contract.hpp — `SimulationContract` (crates/engine-contract/src/contract.rs:17-52) in, the kernel-side `SimulationSpec`
// JSON (crates/engine-core/src/spec.rs:250-417) out: the lowering `contract_to_simulation_spec`
// (crates/engine-core/src/contract_adapter.rs:74-230) performs in front of `run_batch`, so that the hosts' own payload —
// the msgpack `Vec<SimulationContract>` blob engine-py / engine-r forward verbatim (engine-py/src/orchestrator_bridge.rs:23-30,
// `rmp_serde::to_vec_named`: structs are string-keyed maps) or the same tree as serde JSON (engine-wasm) — can be handed to
// this backend unchanged.  SURVEY.md §8(f) rank 3.  Host side only; nothing here touches the device.
//
// Wire conventions (serde defaults): externally tagged enums — a unit variant is the string "Const", a struct variant the
// one-entry map {"Direct": {"column": 0}}; `TestTarget` is internally tagged ({"kind": "marginal", "term": 1},
// test_spec.rs:22-37); `ColumnId` is a transparent u32 (ids.rs); `Option` is nil / the value; fields marked
// `#[serde(default)]` may be absent.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "json.hpp"
#include "spec.hpp"

namespace mcp {
namespace contract {

// ------------------------------------------------------------------------------------------------ msgpack → JsonValue
class MsgpackReader {
  public:
    MsgpackReader(const uint8_t* p, size_t n) : p_(p), end_(p + n) {}
    JsonValue parse() {
        JsonValue v = value(0);
        if (p_ != end_) fail("trailing bytes");
        return v;
    }

  private:
    const uint8_t* p_;
    const uint8_t* end_;
    [[noreturn]] static void fail(const char* what) { invalid(std::string("malformed contracts bytes: ") + what); }
    void need(size_t n) {
        if ((size_t)(end_ - p_) < n) fail("unexpected end of input");
    }
    uint64_t be(int n) {
        need((size_t)n);
        uint64_t v = 0;
        for (int i = 0; i < n; i++) v = (v << 8) | *p_++;
        return v;
    }
    static JsonValue number(double d) {
        JsonValue v;
        v.kind = JsonValue::Number;
        v.num = d;
        return v;
    }
    JsonValue str(size_t n) {
        need(n);
        JsonValue v;
        v.kind = JsonValue::String;
        v.str.assign(reinterpret_cast<const char*>(p_), n);
        p_ += n;
        return v;
    }
    JsonValue array(size_t n, int depth) {
        JsonValue v;
        v.kind = JsonValue::Array;
        if (n > (size_t)(end_ - p_)) fail("array length exceeds input");
        v.arr.reserve(n);
        for (size_t i = 0; i < n; i++) v.arr.push_back(value(depth + 1));
        return v;
    }
    JsonValue map(size_t n, int depth) {
        JsonValue v;
        v.kind = JsonValue::Object;
        for (size_t i = 0; i < n; i++) {
            JsonValue k = value(depth + 1);
            if (k.kind != JsonValue::String) fail("map key is not a string");
            v.obj[k.str] = value(depth + 1);
        }
        return v;
    }
    JsonValue value(int depth) {
        if (depth > 64) fail("nesting too deep");
        need(1);
        const uint8_t t = *p_++;
        if (t <= 0x7f) return number((double)t);
        if (t >= 0xe0) return number((double)(int8_t)t);
        if (t >= 0xa0 && t <= 0xbf) return str(t & 0x1f);
        if (t >= 0x90 && t <= 0x9f) return array(t & 0x0f, depth);
        if (t >= 0x80 && t <= 0x8f) return map(t & 0x0f, depth);
        JsonValue v;
        switch (t) {
            case 0xc0: return v;
            case 0xc2: v.kind = JsonValue::Bool; v.b = false; return v;
            case 0xc3: v.kind = JsonValue::Bool; v.b = true; return v;
            case 0xc4: return str((size_t)be(1));
            case 0xc5: return str((size_t)be(2));
            case 0xc6: return str((size_t)be(4));
            case 0xca: { const uint32_t b = (uint32_t)be(4); float f; std::memcpy(&f, &b, 4); return number((double)f); }
            case 0xcb: { const uint64_t b = be(8); double d; std::memcpy(&d, &b, 8); return number(d); }
            case 0xcc: return number((double)be(1));
            case 0xcd: return number((double)be(2));
            case 0xce: return number((double)be(4));
            case 0xcf: return number((double)be(8));
            case 0xd0: return number((double)(int8_t)be(1));
            case 0xd1: return number((double)(int16_t)be(2));
            case 0xd2: return number((double)(int32_t)be(4));
            case 0xd3: return number((double)(int64_t)be(8));
            case 0xd9: return str((size_t)be(1));
            case 0xda: return str((size_t)be(2));
            case 0xdb: return str((size_t)be(4));
            case 0xdc: return array((size_t)be(2), depth);
            case 0xdd: return array((size_t)be(4), depth);
            case 0xde: return map((size_t)be(2), depth);
            case 0xdf: return map((size_t)be(4), depth);
            default: fail("unsupported msgpack type");
        }
    }
};

// ------------------------------------------------------------------------------------------------ JsonValue → text
inline void write_number(std::string& out, double d) {
    if (!std::isfinite(d)) invalid("non-finite number in the contract (JSON cannot carry NaN / infinity)");
    char buf[40];
    if (d == std::floor(d) && std::fabs(d) < 9007199254740992.0) std::snprintf(buf, sizeof buf, "%.0f", d);
    else std::snprintf(buf, sizeof buf, "%.17g", d);  // round-trips every finite f64 through strtod
    out += buf;
}
inline void write_string(std::string& out, const std::string& s) {
    out += '"';
    for (char c : s) {
        if (c == '"' || c == '\\') { out += '\\'; out += c; }
        else if ((unsigned char)c < 0x20) { char b[8]; std::snprintf(b, sizeof b, "\\u%04x", c); out += b; }
        else out += c;
    }
    out += '"';
}
inline void write_json(std::string& out, const JsonValue& v) {
    switch (v.kind) {
        case JsonValue::Null: out += "null"; break;
        case JsonValue::Bool: out += v.b ? "true" : "false"; break;
        case JsonValue::Number: write_number(out, v.num); break;
        case JsonValue::String: write_string(out, v.str); break;
        case JsonValue::Array: {
            out += '[';
            for (size_t i = 0; i < v.arr.size(); i++) {
                if (i) out += ',';
                write_json(out, v.arr[i]);
            }
            out += ']';
            break;
        }
        case JsonValue::Object: {
            out += '{';
            bool first = true;
            for (auto& kv : v.obj) {
                if (!first) out += ',';
                first = false;
                write_string(out, kv.first);
                out += ':';
                write_json(out, kv.second);
            }
            out += '}';
            break;
        }
    }
}

// ------------------------------------------------------------------------------------------------ typed contract
struct Column {  // generation.rs ColumnSpec
    enum Kind { Synthetic, Resampled, ResampledBinary, FactorSynthetic, FactorFromFrame } kind = Synthetic;
    std::string dist = "normal";  // SyntheticKind lowered to the kernel's snake_case Distribution name
    double param = 0.0;           // Binary p / ResampledBinary proportion
    bool pinned = false;
    uint32_t frame_column = 0, n_levels = 0;
    std::vector<double> proportions;
    int sampled = -1;             // Option<bool>: −1 None
    bool continuous() const { return kind == Synthetic || kind == Resampled || kind == ResampledBinary; }
};
struct Term {  // design.rs DesignTerm
    enum Kind { Const, Direct, DummyOf, Interaction } kind = Const;
    uint32_t column = 0, level_index = 0;
    std::vector<Term> components;
    bool operator==(const Term& o) const {
        return kind == o.kind && column == o.column && level_index == o.level_index && components == o.components;
    }
};
struct Target {  // test_spec.rs TestTarget
    enum Kind { Marginal, Joint, Contrast } kind = Marginal;
    uint32_t term = 0, positive = 0, negative = 0;
    std::vector<uint32_t> terms;
};
struct Posthoc {
    uint32_t factor_column = 0;
    std::vector<uint32_t> target_term_indices;
    bool has_alpha = false;
    double alpha = 0.0;
};
struct Contract {
    std::vector<Column> columns;
    bool corr_identity = true;
    std::vector<uint32_t> corr_columns;
    std::vector<double> corr_values;
    JsonValue cluster;  // Option<ClusterSpec>, re-exported verbatim into the spec (contract_adapter.rs:118-121)
    std::vector<uint32_t> slope_columns;                     // cluster.slopes[].column
    std::vector<std::vector<uint32_t>> extra_slope_columns;  // cluster.extra_groupings[].slopes[].column
    bool has_frame = false, frame_bootstrap = false;
    std::vector<double> frame_data;
    uint32_t frame_rows = 0, frame_cols = 0;
    std::vector<uint32_t> cluster_level_columns;
    std::vector<Term> design_generation, design_test;
    bool has_design_test = false;
    std::string outcome_kind = "continuous", residual_dist = "normal", estimator = "ols", wald_se = "rx", correction = "none";
    double intercept = 0.0;
    std::vector<double> coefficients;
    bool residual_pinned = false, has_hsk_driver = false, link_probit = false;
    uint32_t hsk_driver = 0;
    int nagq = 1;
    std::vector<Target> targets;
    double alpha = 0.05;
    std::vector<Posthoc> posthoc;
    JsonValue scenario;  // ScenarioPerturbations with new_distributions already lowered to Distribution names
    double max_failed_fraction = 0.0;
    const std::vector<Term>& test_design() const { return has_design_test ? design_test : design_generation; }
};

namespace detail {
using mcp::detail::arr;
using mcp::detail::num;
using mcp::detail::req;
using mcp::detail::to_bool;
inline uint32_t u32(const JsonValue& v, const char* what) {
    const double d = num(v, what);
    if (d != std::floor(d) || d < 0.0 || d > 4294967295.0) invalid(std::string("field `") + what + "` must be a u32");
    return (uint32_t)d;
}
inline const std::string& text(const JsonValue& v, const char* what) {
    if (v.kind != JsonValue::String) invalid(std::string("field `") + what + "` must be a string");
    return v.str;
}
// externally tagged enum: "Unit" or {"Variant": body}
inline const std::string& variant(const JsonValue& v, const char* what, const JsonValue** body) {
    *body = nullptr;
    if (v.kind == JsonValue::String) return v.str;
    if (v.kind == JsonValue::Object && v.obj.size() == 1) {
        *body = &v.obj.begin()->second;
        return v.obj.begin()->first;
    }
    invalid(std::string("field `") + what + "` is not an enum variant");
}
inline const JsonValue* opt(const JsonValue& o, const char* key) {
    const JsonValue* p = o.get(key);
    return p && !p->is_null() ? p : nullptr;
}
inline std::string synthetic_kind(const JsonValue& v, double* param) {  // contract_adapter.rs:232-245
    const JsonValue* body;
    const std::string& name = variant(v, "SyntheticKind", &body);
    *param = 0.0;
    if (name == "Normal") return "normal";
    if (name == "Binary") {
        if (!body) invalid("SyntheticKind::Binary needs `p`");
        *param = num(req(*body, "p"), "p");
        return "binary";
    }
    if (name == "RightSkewed") return "right_skewed";
    if (name == "LeftSkewed") return "left_skewed";
    if (name == "HighKurtosis") return "high_kurtosis";
    if (name == "Uniform") return "uniform";
    invalid("unknown SyntheticKind `" + name + "`");
}
inline Term term_of(const JsonValue& v, int depth = 0) {
    const JsonValue* body;
    const std::string& name = variant(v, "DesignTerm", &body);
    Term t;
    if (name == "Const") return t;
    if (!body) invalid("DesignTerm::" + name + " needs a body");
    if (name == "Direct") {
        t.kind = Term::Direct;
        t.column = u32(req(*body, "column"), "column");
    } else if (name == "DummyOf") {
        t.kind = Term::DummyOf;
        t.column = u32(req(*body, "column"), "column");
        t.level_index = u32(req(*body, "level_index"), "level_index");
    } else if (name == "Interaction") {
        if (depth > 4) invalid("DesignTerm nesting too deep");
        t.kind = Term::Interaction;
        for (auto& c : arr(req(*body, "components"), "components")) t.components.push_back(term_of(c, depth + 1));
    } else {
        invalid("unknown DesignTerm `" + name + "`");
    }
    return t;
}
inline std::vector<Term> design_of(const JsonValue& v) {
    std::vector<Term> out;
    for (auto& t : arr(req(v, "terms"), "terms")) out.push_back(term_of(t));
    return out;
}
inline std::vector<uint32_t> slope_columns(const JsonValue& owner) {
    std::vector<uint32_t> out;
    if (const JsonValue* s = opt(owner, "slopes"))
        for (auto& sl : arr(*s, "slopes")) out.push_back(u32(req(sl, "column"), "slopes.column"));
    return out;
}
}  // namespace detail

inline Contract parse_contract(const JsonValue& root) {
    using namespace detail;
    if (root.kind != JsonValue::Object) invalid("a SimulationContract must be a map");
    Contract c;
    const JsonValue& gen = req(root, "generation");
    for (auto& cv : arr(req(gen, "columns"), "generation.columns")) {
        const JsonValue* b;
        const std::string& name = variant(cv, "ColumnSpec", &b);
        if (!b) invalid("ColumnSpec::" + name + " needs a body");
        Column col;
        if (name == "Synthetic") {
            col.kind = Column::Synthetic;
            col.dist = synthetic_kind(req(*b, "kind"), &col.param);
            if (const JsonValue* p = opt(*b, "pinned")) col.pinned = to_bool(*p, "pinned");
        } else if (name == "Resampled") {
            col.kind = Column::Resampled;
            col.frame_column = u32(req(*b, "frame_column"), "frame_column");
        } else if (name == "ResampledBinary") {
            col.kind = Column::ResampledBinary;
            col.frame_column = u32(req(*b, "frame_column"), "frame_column");
            col.param = num(req(*b, "proportion"), "proportion");
        } else if (name == "FactorSynthetic" || name == "FactorFromFrame") {
            col.kind = name == "FactorSynthetic" ? Column::FactorSynthetic : Column::FactorFromFrame;
            if (col.kind == Column::FactorFromFrame) col.frame_column = u32(req(*b, "frame_column"), "frame_column");
            col.n_levels = u32(req(*b, "n_levels"), "n_levels");
            for (auto& p : arr(req(*b, "proportions"), "proportions")) col.proportions.push_back(num(p, "proportions"));
            if (const JsonValue* p = opt(*b, "sampled_proportions")) col.sampled = to_bool(*p, "sampled_proportions") ? 1 : 0;
        } else {
            invalid("unknown ColumnSpec `" + name + "`");
        }
        c.columns.push_back(col);
    }
    {
        const JsonValue* b;
        const std::string& name = variant(req(gen, "correlations"), "correlations", &b);
        if (name == "Matrix") {
            if (!b) invalid("Correlations::Matrix needs a body");
            c.corr_identity = false;
            for (auto& v : arr(req(*b, "continuous_columns"), "continuous_columns")) c.corr_columns.push_back(u32(v, "continuous_columns"));
            for (auto& v : arr(req(*b, "values"), "values")) c.corr_values.push_back(num(v, "values"));
        } else if (name != "Identity") {
            invalid("unknown Correlations `" + name + "`");
        }
    }
    if (const JsonValue* cl = opt(gen, "cluster")) {
        c.cluster = *cl;
        req(*cl, "sizing");
        num(req(*cl, "tau_squared"), "tau_squared");
        c.slope_columns = slope_columns(*cl);
        if (const JsonValue* eg = opt(*cl, "extra_groupings"))
            for (auto& g : arr(*eg, "extra_groupings")) c.extra_slope_columns.push_back(slope_columns(g));
        // serde(default) fields: make them explicit so that the spec reader sees the same shape either way
        if (!c.cluster.get("slopes")) c.cluster.obj["slopes"].kind = JsonValue::Array;
        if (!c.cluster.get("extra_groupings")) c.cluster.obj["extra_groupings"].kind = JsonValue::Array;
    }
    if (const JsonValue* fr = opt(gen, "uploaded_frame")) {
        c.has_frame = true;
        for (auto& v : arr(req(*fr, "data"), "uploaded_frame.data")) c.frame_data.push_back(num(v, "uploaded_frame.data"));
        c.frame_rows = u32(req(*fr, "n_rows"), "n_rows");
        c.frame_cols = u32(req(*fr, "n_cols"), "n_cols");
        if (const JsonValue* p = opt(*fr, "bootstrap")) c.frame_bootstrap = to_bool(*p, "bootstrap");
        if ((size_t)c.frame_rows * c.frame_cols != c.frame_data.size()) invalid("uploaded_frame.data length does not match n_rows × n_cols");
    }
    if (const JsonValue* p = opt(gen, "cluster_level_columns"))
        for (auto& v : arr(*p, "cluster_level_columns")) c.cluster_level_columns.push_back(u32(v, "cluster_level_columns"));
    c.design_generation = design_of(req(root, "design_generation"));
    if (const JsonValue* dt = opt(root, "design_test")) {
        c.has_design_test = true;
        c.design_test = design_of(*dt);
    }
    {
        const JsonValue& o = req(root, "outcome");
        c.outcome_kind = text(req(o, "kind"), "outcome.kind");
        if (c.outcome_kind != "continuous" && c.outcome_kind != "binary" && c.outcome_kind != "count") invalid("unknown OutcomeKind `" + c.outcome_kind + "`");
        c.intercept = num(req(o, "intercept"), "outcome.intercept");
        for (auto& v : arr(req(o, "coefficients"), "outcome.coefficients")) c.coefficients.push_back(num(v, "outcome.coefficients"));
        const JsonValue& r = req(o, "residual");
        c.residual_dist = text(req(r, "distribution"), "residual.distribution");
        mcp::detail::resid_of(req(r, "distribution"), "residual.distribution");
        if (const JsonValue* p = opt(r, "pinned")) c.residual_pinned = to_bool(*p, "residual.pinned");
        if (const JsonValue* p = opt(o, "heteroskedasticity_driver")) {
            c.has_hsk_driver = true;
            c.hsk_driver = u32(*p, "heteroskedasticity_driver");
        }
        if (const JsonValue* p = opt(o, "link")) {
            if (text(*p, "link") != "probit") invalid("unknown LinkKind `" + p->str + "`");
            c.link_probit = true;
        }
    }
    c.estimator = text(req(root, "estimator"), "estimator");
    if (c.estimator != "ols" && c.estimator != "glm" && c.estimator != "mle") invalid("unknown EstimatorSpec `" + c.estimator + "`");
    if (const JsonValue* p = opt(root, "wald_se")) c.wald_se = text(*p, "wald_se");
    if (c.wald_se != "rx" && c.wald_se != "hessian") invalid("unknown WaldSe `" + c.wald_se + "`");
    if (const JsonValue* p = opt(root, "nagq")) c.nagq = (int)u32(*p, "nagq");
    {
        const JsonValue& t = req(root, "test");
        for (auto& tv : arr(req(t, "targets"), "test.targets")) {
            const std::string& kind = text(req(tv, "kind"), "target.kind");
            Target tg;
            if (kind == "marginal") {
                tg.kind = Target::Marginal;
                tg.term = u32(req(tv, "term"), "term");
            } else if (kind == "joint") {
                tg.kind = Target::Joint;
                for (auto& v : arr(req(tv, "terms"), "terms")) tg.terms.push_back(u32(v, "terms"));
            } else if (kind == "contrast") {
                tg.kind = Target::Contrast;
                tg.positive = u32(req(tv, "positive"), "positive");
                tg.negative = u32(req(tv, "negative"), "negative");
            } else {
                invalid("unknown TestTarget kind `" + kind + "`");
            }
            c.targets.push_back(tg);
        }
        c.correction = text(req(t, "correction"), "test.correction");
        mcp::detail::enum_of(req(t, "correction"), "test.correction",
                             {{"none", 0}, {"bonferroni", 1}, {"holm", 2}, {"benjamini_hochberg", 3}, {"tukey_hsd", 4}});
        c.alpha = num(req(t, "alpha"), "test.alpha");
    }
    if (const JsonValue* p = opt(root, "posthoc"))
        for (auto& pv : arr(*p, "posthoc")) {
            Posthoc ph;
            ph.factor_column = u32(req(pv, "factor_column"), "factor_column");
            for (auto& v : arr(req(pv, "target_term_indices"), "target_term_indices")) ph.target_term_indices.push_back(u32(v, "target_term_indices"));
            if (const JsonValue* a = opt(pv, "posthoc_alpha")) {
                ph.has_alpha = true;
                ph.alpha = num(*a, "posthoc_alpha");
            }
            c.posthoc.push_back(ph);
        }
    {
        // translate_scenario (contract_adapter.rs:868-896): everything passes through except new_distributions
        const JsonValue& s = req(root, "scenario");
        if (s.kind != JsonValue::Object) invalid("scenario must be a map");
        c.scenario = s;
        for (const char* key : {"heterogeneity", "heteroskedasticity_ratio", "correlation_noise_sd", "distribution_change_prob",
                                "residual_change_prob", "residual_df"})
            num(req(s, key), key);
        text(req(s, "name"), "scenario.name");
        JsonValue nd;
        nd.kind = JsonValue::Array;
        for (auto& v : arr(req(s, "new_distributions"), "new_distributions")) {
            double unused;
            JsonValue name;
            name.kind = JsonValue::String;
            name.str = synthetic_kind(v, &unused);
            nd.arr.push_back(name);
        }
        c.scenario.obj["new_distributions"] = nd;
        for (auto& v : arr(req(s, "residual_dists"), "residual_dists")) mcp::detail::resid_of(v, "residual_dists");
        for (const char* key : {"sampled_factor_proportions", "truth_start"})
            if (!opt(s, key)) {
                c.scenario.obj[key].kind = JsonValue::Bool;
                c.scenario.obj[key].b = false;
            }
        if (!s.get("lme")) c.scenario.obj["lme"] = JsonValue();
    }
    c.max_failed_fraction = num(req(root, "max_failed_fraction"), "max_failed_fraction");
    return c;
}

// ------------------------------------------------------------------------------------------------ validate()
// engine-contract/src/validate.rs:22-50, invariants in the reference's order.  19-21 (random-slope / extra-grouping
// structure) are not restated: the CUDA backend rejects those specs as unsupported before any of their fields is read.
[[noreturn]] inline void contract_invalid(const std::string& m) { invalid("contract failed validate(): " + m); }

inline void validate_contract(const Contract& c) {
    const size_t ncol = c.columns.size();
    const std::vector<Term>& dt = c.test_design();
    std::vector<const std::vector<Term>*> designs{&c.design_generation};
    if (c.has_design_test) designs.push_back(&c.design_test);
    // 01
    if (c.coefficients.size() != c.design_generation.size())
        contract_invalid("outcome.coefficients length " + std::to_string(c.coefficients.size()) + " does not match design_generation terms " +
                         std::to_string(c.design_generation.size()));
    // 02
    auto check_term = [&](uint32_t t) {
        if (t >= dt.size()) contract_invalid("test target term " + std::to_string(t) + " out of range (n_terms = " + std::to_string(dt.size()) + ")");
    };
    for (auto& t : c.targets) {
        if (t.kind == Target::Marginal) check_term(t.term);
        else if (t.kind == Target::Joint) for (uint32_t x : t.terms) check_term(x);
        else { check_term(t.positive); check_term(t.negative); }
    }
    // 03
    if (c.targets.empty() && c.posthoc.empty()) contract_invalid("test.targets must be non-empty");
    {
        std::set<uint32_t> seen_m;
        std::set<std::pair<uint32_t, uint32_t>> seen_c;
        for (auto& t : c.targets) {
            if (t.kind == Target::Marginal) {
                if (!seen_m.insert(t.term).second) contract_invalid("duplicate Marginal target");
            } else if (t.kind == Target::Joint) {
                if (t.terms.size() < 2) contract_invalid("Joint target must have >= 2 terms");
                std::set<uint32_t> local;
                for (uint32_t x : t.terms)
                    if (!local.insert(x).second) contract_invalid("duplicate term inside Joint target");
            } else {
                if (t.positive == t.negative) contract_invalid("Contrast positive and negative must differ");
                if (!seen_c.insert({std::min(t.positive, t.negative), std::max(t.positive, t.negative)}).second) contract_invalid("duplicate Contrast pair");
            }
        }
    }
    // 04
    auto check_col = [&](uint32_t id) {
        if (id >= ncol) contract_invalid("column id " + std::to_string(id) + " out of range (n_columns = " + std::to_string(ncol) + ")");
    };
    for (auto* d : designs)
        for (auto& t : *d) {
            if (t.kind == Term::Direct || t.kind == Term::DummyOf) check_col(t.column);
            for (auto& comp : t.components)
                if (comp.kind == Term::Direct || comp.kind == Term::DummyOf) check_col(comp.column);
        }
    if (c.has_hsk_driver) check_col(c.hsk_driver);
    for (uint32_t id : c.corr_columns) check_col(id);
    for (auto& ph : c.posthoc) check_col(ph.factor_column);
    for (uint32_t id : c.cluster_level_columns) check_col(id);
    // 05, 06
    if (!c.corr_identity) {
        const size_t n = c.corr_columns.size();
        if (c.corr_values.size() != n * n)
            contract_invalid("correlation values length " + std::to_string(c.corr_values.size()) + " does not match " + std::to_string(n) + "²");
        for (uint32_t id : c.corr_columns)
            if (!c.columns[id].continuous()) contract_invalid("correlation references non-continuous column " + std::to_string(id));
    }
    // 07
    for (auto& col : c.columns) {
        if (col.continuous()) continue;
        if (col.n_levels < 2) contract_invalid("invalid factor proportions: n_levels must be >= 2");
        if (col.proportions.size() != col.n_levels) contract_invalid("invalid factor proportions: proportions.len must equal n_levels");
        double sum = 0.0;
        for (double p : col.proportions) sum += p;
        if (!(std::fabs(sum - 1.0) <= 1e-6)) contract_invalid("invalid factor proportions: proportions must sum to 1.0");
    }
    // 08, 09
    auto check_dummy = [&](const Term& t) {
        const Column& col = c.columns[t.column];
        if (col.continuous()) contract_invalid("DummyOf references non-factor column " + std::to_string(t.column));
        if (t.level_index == 0 || t.level_index >= col.n_levels)
            contract_invalid("dummy level " + std::to_string(t.level_index) + " out of range for column " + std::to_string(t.column));
    };
    auto check_direct = [&](const Term& t) {
        if (!c.columns[t.column].continuous()) contract_invalid("Direct term on factor column " + std::to_string(t.column));
    };
    for (auto* d : designs)
        for (auto& t : *d) {
            if (t.kind == Term::DummyOf) check_dummy(t);
            for (auto& comp : t.components)
                if (comp.kind == Term::DummyOf) check_dummy(comp);
        }
    for (auto* d : designs)
        for (auto& t : *d) {
            if (t.kind == Term::Direct) check_direct(t);
            for (auto& comp : t.components)
                if (comp.kind == Term::Direct) check_direct(comp);
        }
    // 10
    if (c.has_hsk_driver && !c.columns[c.hsk_driver].continuous()) contract_invalid("heteroskedasticity driver on factor column " + std::to_string(c.hsk_driver));
    // 11
    {
        long max_fc = -1;
        for (auto& col : c.columns)
            if (col.kind == Column::Resampled || col.kind == Column::ResampledBinary || col.kind == Column::FactorFromFrame)
                max_fc = std::max(max_fc, (long)col.frame_column);
        if (max_fc >= 0 && (!c.has_frame || max_fc >= (long)c.frame_cols))
            contract_invalid("uploaded frame missing for frame_column " + std::to_string(max_fc));
    }
    // 12, 13
    if (c.estimator == "glm" && c.outcome_kind == "continuous") contract_invalid("estimator Glm requires a binary or count outcome");
    if (c.estimator == "mle" && c.cluster.is_null()) contract_invalid("estimator Mle requires generation.cluster");
    if (const JsonValue* lme = c.scenario.get("lme"))
        if (!lme->is_null() && !(c.estimator == "mle" || (c.estimator == "glm" && !c.cluster.is_null())))
            contract_invalid("scenario.lme requires estimator Mle");
    // 14, 15
    if (!(c.max_failed_fraction >= 0.0 && c.max_failed_fraction <= 1.0)) contract_invalid("max_failed_fraction out of [0, 1]");
    if (!(c.alpha > 0.0 && c.alpha < 1.0)) contract_invalid("alpha out of (0, 1)");
    // 16 — tolerant Cholesky (EPS = −1e-8; zero pivots allowed when their column is zero too)
    if (!c.corr_identity && !c.corr_columns.empty()) {
        const size_t n = c.corr_columns.size();
        const std::vector<double>& a = c.corr_values;
        std::vector<double> l(n * n, 0.0);
        for (size_t j = 0; j < n; j++) {
            double s = a[j * n + j];
            for (size_t k = 0; k < j; k++) s -= l[k * n + j] * l[k * n + j];
            if (s < -1e-8) contract_invalid("correlation matrix is not positive semi-definite");
            s = std::sqrt(std::max(s, 0.0));
            l[j * n + j] = s;
            for (size_t i = j + 1; i < n; i++) {
                double sum = a[j * n + i];
                for (size_t k = 0; k < j; k++) sum -= l[k * n + i] * l[k * n + j];
                if (s == 0.0) {
                    if (std::fabs(sum) > 1e-8) contract_invalid("correlation matrix is not positive semi-definite");
                } else {
                    l[j * n + i] = sum / s;
                }
            }
        }
    }
    // 17
    if (!c.posthoc.empty()) {
        if (c.estimator != "ols") contract_invalid("posthoc requires estimator Ols");
        for (auto& ph : c.posthoc) {
            if (c.columns[ph.factor_column].continuous()) contract_invalid("invalid posthoc: factor_column is not a factor");
            for (uint32_t t : ph.target_term_indices) {
                if (t >= dt.size()) contract_invalid("invalid posthoc: posthoc target_term_index out of range");
                if (!(dt[t].kind == Term::DummyOf && dt[t].column == ph.factor_column))
                    contract_invalid("invalid posthoc: posthoc target must be a DummyOf of factor_column");
            }
        }
    }
    // 18
    for (auto* d : designs)
        for (auto& t : *d)
            if (t.kind == Term::Interaction) {
                if (t.components.size() < 2) contract_invalid("interaction with fewer than 2 components");
                for (auto& comp : t.components)
                    if (comp.kind != Term::Direct && comp.kind != Term::DummyOf) contract_invalid("interaction component must be Direct or DummyOf");
            }
    // 22
    {
        auto f = [&](const char* key) { return c.scenario.get(key)->num; };
        auto fin_nonneg = [&](const char* key) {
            if (!(std::isfinite(f(key)) && f(key) >= 0.0)) contract_invalid(std::string("scenario perturbation `") + key + "` out of range");
        };
        auto prob = [&](const char* key) {
            if (!(std::isfinite(f(key)) && f(key) >= 0.0 && f(key) <= 1.0)) contract_invalid(std::string("scenario perturbation `") + key + "` out of range");
        };
        fin_nonneg("heterogeneity");
        fin_nonneg("correlation_noise_sd");
        fin_nonneg("heteroskedasticity_ratio");
        prob("distribution_change_prob");
        prob("residual_change_prob");
        if (!std::isfinite(f("residual_df"))) contract_invalid("scenario perturbation `residual_df` out of range");
        if (const JsonValue* lme = c.scenario.get("lme"))
            if (!lme->is_null())
                for (const char* key : {"random_effect_df", "icc_noise_sd"}) {
                    const double v = mcp::detail::num(mcp::detail::req(*lme, key), key);
                    if (!(std::isfinite(v) && v >= 0.0)) contract_invalid(std::string("lme perturbation `") + key + "` out of range");
                }
    }
    // 23
    for (auto& col : c.columns) {
        const bool bin = (col.kind == Column::Synthetic && col.dist == "binary") || col.kind == Column::ResampledBinary;
        if (bin && !(std::isfinite(col.param) && col.param >= 0.0 && col.param <= 1.0))
            contract_invalid(col.kind == Column::Synthetic ? "binary probability out of [0, 1]" : "resampled binary proportion out of [0, 1]");
    }
    // 24, 25
    if (c.link_probit && c.outcome_kind != "binary") contract_invalid("probit link requires a binary outcome");
    if (c.nagq == 0 || c.nagq > 25 || c.nagq % 2 == 0) contract_invalid("nagq out of range");
}

// ------------------------------------------------------------------------------------------------ lowering
struct Lowering {
    std::string spec_json;
    std::vector<uint32_t> report_targets;                        // design_test term positions, in the spec's target order
    std::vector<std::pair<uint32_t, uint32_t>> report_contrasts; // (contract_adapter.rs:629-661)
};

namespace detail {
struct Positions {
    const Contract& c;
    uint32_t n_non_factor = 0, n_factor_dummies = 0;
    explicit Positions(const Contract& cc) : c(cc) {
        for (auto& col : c.columns) {
            if (col.continuous()) n_non_factor++;
            else n_factor_dummies += col.n_levels > 0 ? col.n_levels - 1 : 0;
        }
    }
    uint32_t continuous(uint32_t gen_idx) const {  // contract_adapter.rs:722-737 (1-based x_full index)
        uint32_t cursor = 1;
        for (size_t i = 0; i < c.columns.size(); i++)
            if (c.columns[i].continuous()) {
                if (i == gen_idx) return cursor;
                cursor++;
            }
        contract_invalid("Direct references a non-continuous column");
    }
    uint32_t dummy(uint32_t gen_idx, uint32_t level_index) const {  // contract_adapter.rs:739-757
        uint32_t cursor = 1 + n_non_factor;
        for (size_t i = 0; i < c.columns.size(); i++)
            if (!c.columns[i].continuous()) {
                if (i == gen_idx) return cursor + (level_index - 1);
                cursor += c.columns[i].n_levels > 0 ? c.columns[i].n_levels - 1 : 0;
            }
        contract_invalid("DummyOf references a non-factor column");
    }
    uint32_t interaction(const std::vector<Term>& components) const {  // contract_adapter.rs:703-720
        uint32_t j = 0;
        for (auto& t : c.design_generation)
            if (t.kind == Term::Interaction) {
                if (t.components == components) return 1 + n_non_factor + n_factor_dummies + j;
                j++;
            }
        contract_invalid("test interaction has no matching generation interaction");
    }
    uint32_t term(const Term& t) const {  // term_to_kernel_col, contract_adapter.rs:527-545
        switch (t.kind) {
            case Term::Const: return 0;
            case Term::Direct: return continuous(t.column);
            case Term::DummyOf: return dummy(t.column, t.level_index);
            default: return interaction(t.components);
        }
    }
    uint32_t factor_position(uint32_t col) const {  // contract_adapter.rs:794-808
        uint32_t k = 0;
        for (size_t i = 0; i < c.columns.size(); i++)
            if (!c.columns[i].continuous()) {
                if (i == col) return k;
                k++;
            }
        contract_invalid("posthoc.factor_column is not a factor");
    }
};
template <class T, class F>
inline void write_list(std::string& out, const char* key, const std::vector<T>& v, F each) {
    out += '"';
    out += key;
    out += "\":[";
    for (size_t i = 0; i < v.size(); i++) {
        if (i) out += ',';
        each(out, v[i]);
    }
    out += "],";
}
inline void write_f64s(std::string& out, const char* key, const std::vector<double>& v) {
    write_list(out, key, v, [](std::string& o, double d) { write_number(o, d); });
}
inline void write_u32s(std::string& out, const char* key, const std::vector<uint32_t>& v) {
    write_list(out, key, v, [](std::string& o, uint32_t d) { o += std::to_string(d); });
}
}  // namespace detail

// contract_to_simulation_spec (contract_adapter.rs:74-230)
// factor_min_level_count: the kernel default is 0; the orchestrator forces its configured limit on every lowered scenario
// (find_power.rs:271, configs/config.json:31).
inline Lowering lower_contract(const Contract& c, int factor_min_level_count = 0) {
    using namespace detail;
    validate_contract(c);
    const std::vector<Term>& dt = c.test_design();
    // Joint targets: the omnibus shape routes to report_overall, anything else is rejected (contract_adapter.rs:50-97)
    bool report_overall = false;
    for (auto& t : c.targets)
        if (t.kind == Target::Joint) {
            std::vector<uint32_t> expected, got(t.terms);
            for (size_t i = 0; i < dt.size(); i++)
                if (dt[i].kind != Term::Const) expected.push_back((uint32_t)i);
            std::sort(got.begin(), got.end());
            got.erase(std::unique(got.begin(), got.end()), got.end());
            if (got != expected) {
                std::string s;
                for (uint32_t x : got) s += (s.empty() ? "" : ", ") + std::to_string(x);
                invalid("TestTarget::Joint only supports the v1-parity omnibus shape (all non-intercept terms); got [" + s + "]");
            }
            report_overall = true;
        }
    const Positions pos(c);
    const uint32_t k = pos.n_non_factor, nfd = pos.n_factor_dummies;
    Lowering out;
    std::string& o = out.spec_json;
    o += '{';
    o += "\"n_non_factor\":" + std::to_string(k) + ",\"n_factor_dummies\":" + std::to_string(nfd) + ",";
    // translate_generation (contract_adapter.rs:247-317)
    std::vector<std::string> var_types;
    std::vector<double> var_params, f_props;
    std::vector<int> var_pinned, f_sampled;
    std::vector<uint32_t> f_levels;
    for (auto& col : c.columns) {
        if (col.continuous()) {
            var_types.push_back(col.kind == Column::Synthetic ? col.dist : col.kind == Column::Resampled ? "uploaded_data" : "uploaded_binary");
            var_pinned.push_back(col.kind == Column::Synthetic && col.pinned);
            var_params.push_back(col.kind == Column::Resampled ? 0.0 : col.param);
        } else {
            f_levels.push_back(col.n_levels);
            f_props.insert(f_props.end(), col.proportions.begin(), col.proportions.end());
            f_sampled.push_back(col.sampled);
        }
    }
    // translate_correlation (contract_adapter.rs:319-334)
    std::vector<double> corr;
    if (c.corr_identity) {
        corr.assign((size_t)k * k, 0.0);
        for (uint32_t i = 0; i < k; i++) corr[(size_t)i * k + i] = 1.0;
    } else {
        corr = c.corr_values;
    }
    write_f64s(o, "correlation", corr);
    write_list(o, "var_types", var_types, [](std::string& s, const std::string& v) { write_string(s, v); });
    write_list(o, "var_pinned", var_pinned, [](std::string& s, int v) { s += v ? "true" : "false"; });
    write_f64s(o, "var_params", var_params);
    // translate_uploaded_frame (contract_adapter.rs:359-455)
    {
        std::vector<double> upload_normal;
        uint32_t un_rows = 0, un_cols = 0;
        std::string fmap = "[";
        if (c.has_frame && c.frame_bootstrap) {
            bool first = true;
            auto push = [&](long v) {
                if (!first) fmap += ',';
                first = false;
                fmap += v < 0 ? std::string("null") : std::to_string(v);
            };
            for (auto& col : c.columns)
                if (col.continuous()) push(col.kind == Column::Synthetic ? -1 : (long)col.frame_column);
            for (auto& col : c.columns)
                if (!col.continuous()) push(col.kind == Column::FactorFromFrame ? (long)col.frame_column : -1);
        } else if (c.has_frame) {
            un_rows = c.frame_rows;
            un_cols = k;
            upload_normal.assign((size_t)un_rows * k, 0.0);
            uint32_t j = 0;
            for (auto& col : c.columns) {
                if (!col.continuous()) continue;
                if (col.kind != Column::Synthetic) {
                    std::vector<double> vals(un_rows);
                    for (uint32_t r = 0; r < un_rows; r++) vals[r] = c.frame_data[(size_t)r * c.frame_cols + col.frame_column];
                    std::sort(vals.begin(), vals.end());  // validate() requires finite data: total_cmp == operator<
                    for (uint32_t r = 0; r < un_rows; r++) upload_normal[(size_t)r * k + j] = vals[r];
                }
                j++;
            }
        }
        fmap += ']';
        write_f64s(o, "upload_normal", upload_normal);
        o += "\"upload_normal_shape\":[" + std::to_string(un_rows) + "," + std::to_string(un_cols) + "],";
        write_f64s(o, "upload_data", c.has_frame ? c.frame_data : std::vector<double>());
        o += "\"upload_data_shape\":[" + std::to_string(c.has_frame ? c.frame_rows : 0) + "," + std::to_string(c.has_frame ? c.frame_cols : 0) + "],";
        o += "\"bootstrap_frame_map\":" + fmap + ",";
    }
    // expand_cluster_level_columns (contract_adapter.rs:759-792)
    {
        std::vector<uint32_t> between;
        for (uint32_t id : c.cluster_level_columns) {
            if (c.columns[id].continuous()) between.push_back(pos.continuous(id));
            else
                for (uint32_t lv = 1; lv < c.columns[id].n_levels; lv++) between.push_back(pos.dummy(id, lv));
        }
        write_u32s(o, "between_var_indices", between);
    }
    write_u32s(o, "factor_n_levels", f_levels);
    write_f64s(o, "factor_proportions", f_props);
    write_list(o, "factor_sampled", f_sampled, [](std::string& s, int v) { s += v < 0 ? "null" : (v ? "true" : "false"); });
    // translate_interactions (contract_adapter.rs:672-698) + translate_effect_sizes (461-525)
    std::vector<std::vector<uint32_t>> interactions;
    for (auto& t : c.design_generation)
        if (t.kind == Term::Interaction) {
            std::vector<uint32_t> comp;
            for (auto& x : t.components) comp.push_back(x.kind == Term::Direct ? pos.continuous(x.column) : pos.dummy(x.column, x.level_index));
            interactions.push_back(comp);
        }
    const uint32_t n_total = 1 + k + nfd + (uint32_t)interactions.size();
    std::vector<double> effects(n_total, 0.0);
    {
        uint32_t inter_idx = 0;
        for (size_t i = 0; i < c.design_generation.size(); i++) {
            const Term& t = c.design_generation[i];
            const uint32_t col = t.kind == Term::Interaction ? 1 + k + nfd + inter_idx++ : pos.term(t);
            effects[col] = c.coefficients[i];
        }
        effects[0] = c.intercept;  // the kernel reads the intercept from effect_sizes[0] (contract_adapter.rs:518-523)
    }
    write_f64s(o, "effect_sizes", effects);
    // translate_targets_and_contrasts (contract_adapter.rs:571-661)
    std::map<uint32_t, uint32_t> marginal;
    std::vector<std::pair<uint32_t, uint32_t>> contrasts;
    for (auto& t : c.targets) {
        if (t.kind == Target::Marginal) {
            marginal[pos.term(dt[t.term])] = t.term;
        } else if (t.kind == Target::Contrast) {
            if (dt[t.negative].kind == Term::Const) {
                marginal[pos.term(dt[t.positive])] = t.positive;  // reference-level collapse: β_ref = 0 in dummy coding
            } else {
                contrasts.push_back({pos.term(dt[t.positive]), pos.term(dt[t.negative])});
                out.report_contrasts.push_back({t.positive, t.negative});
            }
        }
    }
    std::vector<uint32_t> target_indices;
    for (auto& kv : marginal) {
        target_indices.push_back(kv.first);
        out.report_targets.push_back(kv.second);
    }
    write_u32s(o, "target_indices", target_indices);
    // fit_columns (contract_adapter.rs:160-177)
    {
        std::vector<uint32_t> fit_columns;
        if (c.has_design_test) {
            std::set<uint32_t> kept;
            for (auto& t : dt) kept.insert(pos.term(t));
            if (kept.size() != n_total) fit_columns.assign(kept.begin(), kept.end());
        }
        write_u32s(o, "fit_columns", fit_columns);
    }
    write_list(o, "contrast_pairs", contrasts,
               [](std::string& s, const std::pair<uint32_t, uint32_t>& p) { s += "[" + std::to_string(p.first) + "," + std::to_string(p.second) + "]"; });
    write_list(o, "interactions", interactions, [](std::string& s, const std::vector<uint32_t>& comp) {
        s += '[';
        for (size_t i = 0; i < comp.size(); i++) s += (i ? "," : "") + std::to_string(comp[i]);
        s += ']';
    });
    o += "\"correction_method\":\"" + c.correction + "\",";
    // translate_posthoc (contract_adapter.rs:810-850): posthoc_alpha is shared, the first block's value
    o += "\"crit_values\":{\"alpha\":";
    write_number(o, c.alpha);
    o += ",\"posthoc_alpha\":";
    if (!c.posthoc.empty() && c.posthoc[0].has_alpha) write_number(o, c.posthoc[0].alpha);
    else o += "null";
    o += "},";
    o += "\"heteroskedasticity_driver\":" + (c.has_hsk_driver ? std::to_string(pos.continuous(c.hsk_driver)) : std::string("null")) + ",";
    {
        std::vector<uint32_t> slope_cols;
        for (uint32_t id : c.slope_columns) slope_cols.push_back(pos.continuous(id));
        write_u32s(o, "cluster_slope_design_cols", slope_cols);
        write_list(o, "extra_slope_cols", c.extra_slope_columns, [&](std::string& s, const std::vector<uint32_t>& ids) {
            s += '[';
            for (size_t i = 0; i < ids.size(); i++) s += (i ? "," : "") + std::to_string(pos.continuous(ids[i]));
            s += ']';
        });
    }
    o += "\"residual_dist\":\"" + c.residual_dist + "\",\"residual_pinned\":" + (c.residual_pinned ? "true" : "false") + ",";
    o += "\"outcome_kind\":\"" + c.outcome_kind + "\",\"link\":" + (c.link_probit ? "\"probit\"" : "null") + ",";
    o += "\"estimator\":\"" + c.estimator + "\",\"wald_se\":\"" + c.wald_se + "\",\"nagq\":" + std::to_string(c.nagq) + ",";
    o += "\"intercept\":";
    write_number(o, c.intercept);
    o += ',';
    write_list(o, "posthoc", c.posthoc, [&](std::string& s, const Posthoc& ph) {
        s += "{\"factor_index\":" + std::to_string(pos.factor_position(ph.factor_column)) + ",\"target_indices\":[";
        for (size_t i = 0; i < ph.target_term_indices.size(); i++) {
            const Term& t = dt[ph.target_term_indices[i]];
            s += (i ? "," : "") + std::to_string(pos.dummy(t.column, t.level_index));
        }
        s += "]}";
    });
    o += "\"max_failed_fraction\":";
    write_number(o, c.max_failed_fraction);
    o += ",\"cluster\":";
    write_json(o, c.cluster);
    o += ",\"scenario\":";
    write_json(o, c.scenario);
    o += std::string(",\"report_overall\":") + (report_overall ? "true" : "false") + ",\"factor_min_level_count\":" +
         std::to_string(factor_min_level_count) + "}";
    return out;
}

// The contracts of one payload: msgpack `Vec<SimulationContract>` (format 0; a bare map is accepted as a one-element list)
// or the same tree as JSON text (format 1).
inline std::vector<JsonValue> decode_contracts(const uint8_t* bytes, size_t len, int format) {
    JsonValue root;
    if (format == 0) {
        root = MsgpackReader(bytes, len).parse();
    } else if (format == 1) {
        try {
            root = JsonParser(reinterpret_cast<const char*>(bytes), len).parse();
        } catch (const std::runtime_error& e) {
            invalid(std::string("malformed contracts JSON: ") + e.what());
        }
    } else {
        invalid("unknown contract format (0 = msgpack, 1 = JSON)");
    }
    if (root.kind == JsonValue::Object) return {root};
    if (root.kind != JsonValue::Array) invalid("malformed contracts bytes: expected a list of contracts");
    if (root.arr.empty()) invalid("contracts list cannot be empty");  // orchestrator_bridge.rs:26-28
    return root.arr;
}

}  // namespace contract
}  // namespace mcp
