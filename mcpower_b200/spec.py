"""`SimulationSpec` builders — plain dicts in the serde JSON shape the reference uses.

Field names, enum spellings (snake_case) and defaults follow
crates/engine-core/src/spec.rs:250-417 and the frozen bench builders in
crates/engine-core/src/bin/throughput.rs:88-219.  The dict is what
`mcpower_b200.run_batch` serialises and hands to the C ABI (`mcp_run_batch`).
"""
from __future__ import annotations

import copy
import math

_MENUS = (["right_skewed", "left_skewed", "uniform"], ["high_kurtosis", "right_skewed"])


def scenario(name: str = "optimistic", mixed: bool = False) -> dict:
    """Scenario presets (configs/scenarios.json; frozen copies in throughput.rs:77-133).  `mixed=True` is a model with random
    effects: the hosts then carry the preset's random-effect knobs as the `lme` block (heavy-tailed = the t kernel,
    project_contract.rs:922-980); for every other model they zero them, so `lme` stays None (spec_builder.py:246-253)."""
    base = {
        "name": name,
        "heterogeneity": 0.0,
        "heteroskedasticity_ratio": 1.0,
        "correlation_noise_sd": 0.0,
        "distribution_change_prob": 0.0,
        "new_distributions": list(_MENUS[0]),
        "residual_change_prob": 0.0,
        "residual_dists": list(_MENUS[1]),
        "residual_df": 10.0,
        "sampled_factor_proportions": False,
        "truth_start": True,
        "lme": None,
    }
    if name == "optimistic":
        return base
    if name == "realistic":
        base.update(heterogeneity=0.2, heteroskedasticity_ratio=2.0, correlation_noise_sd=0.15,
                    distribution_change_prob=0.5, residual_change_prob=0.5, residual_df=8.0,
                    sampled_factor_proportions=True, truth_start=False)
        if mixed:
            base["lme"] = {"random_effect_dist": "high_kurtosis", "random_effect_df": 10.0, "icc_noise_sd": 0.15}
        return base
    if name == "doomer":
        base.update(heterogeneity=0.4, heteroskedasticity_ratio=4.0, correlation_noise_sd=0.3,
                    distribution_change_prob=0.8, residual_change_prob=0.8, residual_df=5.0,
                    sampled_factor_proportions=True, truth_start=False)
        if mixed:
            base["lme"] = {"random_effect_dist": "high_kurtosis", "random_effect_df": 5.0, "icc_noise_sd": 0.3}
        return base
    raise ValueError(f"unknown scenario preset {name!r}")


def base_spec(k: int) -> dict:
    """All-defaults OLS spec: k independent N(0,1) predictors, zero effects (throughput.rs:167-219)."""
    corr = [0.0] * (k * k)
    for i in range(k):
        corr[i * k + i] = 1.0
    return {
        "n_non_factor": k,
        "n_factor_dummies": 0,
        "correlation": corr,
        "var_types": ["normal"] * k,
        "var_pinned": [],
        "var_params": [0.0] * k,
        "upload_normal": [],
        "upload_normal_shape": [0, 0],
        "upload_data": [],
        "upload_data_shape": [0, 0],
        "bootstrap_frame_map": [],
        "between_var_indices": [],
        "factor_n_levels": [],
        "factor_proportions": [],
        "factor_sampled": [],
        "effect_sizes": [0.0] * (1 + k),
        "target_indices": [1, 2] if k >= 2 else list(range(1, 1 + k)),
        "fit_columns": [],
        "contrast_pairs": [],
        "interactions": [],
        "correction_method": "none",
        "crit_values": {"alpha": 0.05, "posthoc_alpha": None},
        "heteroskedasticity_driver": None,
        "cluster_slope_design_cols": [],
        "extra_slope_cols": [],
        "residual_dist": "normal",
        "residual_pinned": False,
        "outcome_kind": "continuous",
        "link": None,
        "estimator": "ols",
        "wald_se": "hessian",
        "nagq": 1,
        "intercept": 0.0,
        "posthoc": [],
        "max_failed_fraction": 0.1,
        "cluster": None,
        "scenario": scenario("optimistic"),
        "report_overall": False,
        "factor_min_level_count": 0,
    }


def headline_spec() -> dict:
    """BASELINE.json metric shape M: OLS n=500, k=8 (P=9), effects x1=x2=0.18, targets {x1,x2}, alpha 0.05."""
    s = base_spec(8)
    s["effect_sizes"][1] = 0.18
    s["effect_sizes"][2] = 0.18
    return s


def add_factor(spec: dict, proportions, effects) -> dict:
    """Append one synthetic factor (reference-coded dummies after the continuous block)."""
    s = copy.deepcopy(spec)
    if s["interactions"]:
        raise ValueError("add factors before interactions (kernel column layout, data_gen.rs:79-83)")
    nl = len(proportions)
    if len(effects) != nl - 1:
        raise ValueError("one effect per non-reference level")
    p_before = 1 + s["n_non_factor"] + s["n_factor_dummies"]
    s["factor_n_levels"].append(nl)
    s["factor_proportions"] += [float(p) for p in proportions]
    s["n_factor_dummies"] += nl - 1
    s["effect_sizes"] = s["effect_sizes"][:p_before] + [float(e) for e in effects] + s["effect_sizes"][p_before:]
    return s


def add_interaction(spec: dict, cols, effect: float) -> dict:
    s = copy.deepcopy(spec)
    s["interactions"].append([int(c) for c in cols])
    s["effect_sizes"].append(float(effect))
    return s


def n_predictors(spec: dict) -> int:
    return 1 + spec["n_non_factor"] + spec["n_factor_dummies"] + len(spec.get("interactions", []))


def sweep_spec(rho: float = 0.3) -> dict:
    """BASELINE config 2: 8 correlated predictors (AR(1)-style rho^|i-j|) + a 3-level factor
    (0.5, 0.3, 0.2) ⇒ P = 11; every non-intercept coefficient is a target; Bonferroni by default."""
    k = 8
    s = base_spec(k)
    corr = [rho ** abs(i - j) for j in range(k) for i in range(k)]
    s["correlation"] = corr
    s["effect_sizes"] = [0.0, 0.25, 0.2, 0.15, 0.1, 0.1, 0.05, 0.05, 0.0]
    s = add_factor(s, [0.5, 0.3, 0.2], [0.3, 0.5])
    s["target_indices"] = list(range(1, 11))
    s["correction_method"] = "bonferroni"
    s["factor_min_level_count"] = 5  # forced by the orchestrator (find_power.rs:271)
    s["report_overall"] = True
    return s


def logit_intercept(p: float) -> float:
    return math.log(p / (1.0 - p))


def logit_spec(baseline_p: float = 0.3) -> dict:
    """BASELINE config 4: logistic GLM y ~ treatment*x1 + x2 (treatment = Bernoulli(0.5) marginal),
    P = 5 columns [1, treatment, x1, x2, treatment:x1], binary outcome, deviance-LRT overall test."""
    s = base_spec(3)
    s["var_types"] = ["binary", "normal", "normal"]
    s["var_params"] = [0.5, 0.0, 0.0]
    s = add_interaction(s, [1, 2], 0.3)
    s["intercept"] = logit_intercept(baseline_p)
    s["effect_sizes"] = [s["intercept"], 0.5, 0.3, 0.2, 0.3]
    s["target_indices"] = [1, 2, 3, 4]
    s["outcome_kind"] = "binary"
    s["estimator"] = "glm"
    s["report_overall"] = True
    s["max_failed_fraction"] = 0.25
    return s


def lmm_spec(n_clusters: int = 30, icc: float = 0.2, preset: str = "optimistic") -> dict:
    """BASELINE config 5: linear mixed model with a random intercept, `n_clusters` clusters (round-robin rows),
    REML fit, y ~ x1 + x2 + (1 | cluster); tau² = ICC / (1 − ICC) with unit residual variance.  `preset` = the scenario
    preset with the mixed-model knobs (heavy-tailed random effects, ICC jitter) attached."""
    s = base_spec(2)
    s["scenario"] = scenario(preset, mixed=True)
    s["effect_sizes"] = [0.0, 0.3, 0.2]
    s["target_indices"] = [1, 2]
    s["estimator"] = "mle"
    s["cluster"] = {"sizing": {"FixedClusters": {"n_clusters": n_clusters}}, "tau_squared": icc / (1.0 - icc),
                    "slopes": [], "extra_groupings": []}
    s["max_failed_fraction"] = 0.25
    return s


def with_uploaded_frame(spec: dict, frame, columns: dict, factors: dict | None = None, bootstrap: bool = False) -> dict:
    """Attach an uploaded frame the way `translate_uploaded_frame` lowers it (contract_adapter.rs:359-455).

    `frame`: rows × frame-columns (standardised values; factor columns hold level codes).
    `columns`: {non-factor index j: frame column}  → `uploaded_data`, or {j: (frame column, proportion)} →
    `uploaded_binary`.  `factors`: {factor index f: frame column} (FactorFromFrame; strict mode copies the code).
    NORTA (`bootstrap=False`): `upload_normal[r][j]` = r-th order statistic of the mapped frame column.
    Strict (`bootstrap=True`): the raw frame + `bootstrap_frame_map` (non-factor columns first, then factors)."""
    s = copy.deepcopy(spec)
    rows = [[float(v) for v in row] for row in frame]
    u, nc = len(rows), len(rows[0]) if rows else 0
    k, nf = s["n_non_factor"], len(s["factor_n_levels"])
    s["var_params"] = (list(s.get("var_params") or []) + [0.0] * k)[:k]
    for j, m in columns.items():
        if isinstance(m, tuple):
            s["var_types"][j], s["var_params"][j] = "uploaded_binary", float(m[1])
        else:
            s["var_types"][j], s["var_params"][j] = "uploaded_data", 0.0
    s["upload_data"] = [v for row in rows for v in row]
    s["upload_data_shape"] = [u, nc]
    if bootstrap:
        fmap = [None] * (k + nf)
        for j, m in columns.items():
            fmap[j] = int(m[0] if isinstance(m, tuple) else m)
        for f, fc in (factors or {}).items():
            fmap[k + f] = int(fc)
        s["bootstrap_frame_map"] = fmap
        s["upload_normal"], s["upload_normal_shape"] = [], [0, 0]
    else:
        table = [[0.0] * k for _ in range(u)]
        for j, m in columns.items():
            fc = int(m[0] if isinstance(m, tuple) else m)
            for r, v in enumerate(sorted(row[fc] for row in rows)):
                table[r][j] = v
        s["upload_normal"] = [v for row in table for v in row]
        s["upload_normal_shape"] = [u, k]
        s["bootstrap_frame_map"] = []
    return s
