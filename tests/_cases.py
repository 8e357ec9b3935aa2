"""Shared test helpers: load the committed reference fixtures and rebuild each
validation case's `SimulationSpec` from its recorded description
(tests/golden/make_fixtures.py documents where the fixtures come from)."""
from __future__ import annotations

import gzip
import json
import math
import re
from functools import lru_cache
from pathlib import Path

import numpy as np

from mcpower_b200 import spec as S

GOLDEN = Path(__file__).resolve().parent / "golden"


@lru_cache(maxsize=1)
def validation_cases() -> dict:
    with gzip.open(GOLDEN / "validation_cases.json.gz", "rt") as f:
        return json.load(f)


def _one(v):
    return v[0] if isinstance(v, list) else v


def case_arrays(case: dict):
    n, p = case["n"], case["p"]
    x = np.array(case["design_f32_colmajor"], dtype=np.uint32).view(np.float32).reshape(p, n).T
    y = np.array(case["outcome_f32"], dtype=np.uint32).view(np.float32)
    return x, y


def golden_arrays(case: dict):
    g = case["golden"]
    return {k: np.array([float.fromhex(h) for h in g[k]]) for k in g}


def _phi_inv(p: float) -> float:
    from scipy.stats import norm
    return float(norm.ppf(p))


def case_spec(case: dict) -> dict:
    """SimulationSpec for a validation case (sim 0 of base_seed = splitmix64(seed), max_n = n)."""
    d = case["spec"]
    family = _one(d["family"])
    cols = case["columns"]
    k = sum(1 for c in cols if c.startswith("col_"))
    factors = []
    for vt in d.get("variable_types", []) or []:
        m = re.match(r"\s*(\w+)\s*=\s*\(factor,([^)]*)\)", vt)
        if m:
            factors.append((m.group(1), [float(v) for v in m.group(2).split(",")]))
    fac_names = [f[0] for f in factors]
    effects = {}
    for part in _one(d["effects"]).split(","):
        name, val = part.split("=")
        effects[name.strip()] = float(val)
    # continuous variable names in order of first appearance in the formula
    formula = _one(d["formula"]).split("~")[1]
    names = []
    for tok in re.findall(r"[A-Za-z_]\w*", formula.split("(")[0] if "(1|" in formula else formula):
        if tok not in fac_names and tok not in names:
            names.append(tok)
    assert len(names) == k, (names, k, case["label"])
    s = S.base_spec(k)
    col_of = {nm: 1 + i for i, nm in enumerate(names)}
    for nm in names:
        s["effect_sizes"][col_of[nm]] = effects.get(nm, 0.0)
    for c in d.get("correlations", []) or []:
        m = re.match(r"\s*corr\((\w+),\s*(\w+)\)\s*=\s*([-\d.eE]+)", c)
        a, b, r = col_of[m.group(1)] - 1, col_of[m.group(2)] - 1, float(m.group(3))
        s["correlation"][a * k + b] = r
        s["correlation"][b * k + a] = r
    for fname, props in factors:
        base = 1 + s["n_non_factor"] + s["n_factor_dummies"]
        effs = [effects.get(f"{fname}[{lvl}]", 0.0) for lvl in range(2, len(props) + 1)]
        s = S.add_factor(s, props, effs)
        for lvl in range(2, len(props) + 1):
            col_of[f"{fname}[{lvl}]"] = base + lvl - 2
    n_inter = sum(1 for c in cols if c.startswith("interaction_"))
    inter_keys = [key for key in effects if ":" in key]
    assert len(inter_keys) == n_inter, (inter_keys, cols)
    for key in inter_keys:
        comps = [col_of[t.strip()] for t in key.split(":")]
        s = S.add_interaction(s, comps, effects[key])
    P = S.n_predictors(s)
    assert P == case["p"], (P, case["p"])
    s["target_indices"] = list(range(1, P))
    if family == "logit":
        s["outcome_kind"], s["estimator"] = "binary", "glm"
        s["intercept"] = S.logit_intercept(_one(d["baseline_probability"]))
        s["effect_sizes"][0] = s["intercept"]
    elif family == "probit":
        s["outcome_kind"], s["estimator"], s["link"] = "binary", "glm", "probit"
        s["intercept"] = _phi_inv(_one(d["baseline_probability"]))
        s["effect_sizes"][0] = s["intercept"]
    elif family == "poisson":
        s["outcome_kind"], s["estimator"] = "count", "glm"
        s["intercept"] = math.log(_one(d["baseline_rate"]))
        s["effect_sizes"][0] = s["intercept"]
    elif family == "lme":
        cl = d["cluster"]
        icc = _one(cl["ICC"])
        s["estimator"] = "mle"
        s["cluster"] = {"sizing": {"FixedClusters": {"n_clusters": int(_one(cl["n_clusters"]))}},
                        "tau_squared": icc / (1.0 - icc), "slopes": [], "extra_groupings": []}
        for sl in d.get("slopes", []) or []:  # lmm_slope_*: random slope on generation column `column` (ColumnId)
            col = int(_one(sl["column"]))
            s["cluster"]["slopes"].append({"column": col, "variance": float(_one(sl["variance"])),
                                           "corr_with_intercept": float(_one(sl["corr_with_intercept"])),
                                           "corr_with": [float(v) for v in (sl.get("corr_with") or [])] if isinstance(sl.get("corr_with"), list)
                                           else ([float(sl["corr_with"])] if sl.get("corr_with") not in (None, [], {}) else [])})
            s["cluster_slope_design_cols"].append(1 + col)
        for eg in d.get("extra_groupings", []) or []:  # lmm_crossed_* / lmm_nested_*: GroupingSpec as the contract carries it
            rel = eg["relation"]
            if "Crossed" in rel:
                relation = {"Crossed": {"n_clusters": int(_one(rel["Crossed"]["n_clusters"]))}}
            else:
                relation = {"NestedWithin": {"n_per_parent": int(_one(rel["NestedWithin"]["n_per_parent"]))}}
            s["cluster"]["extra_groupings"].append({"relation": relation, "tau_squared": float(_one(eg["tau_squared"])), "slopes": []})
        for eg in d.get("extra", []) or []:  # lmm_slope_crossed: the composition cases' shorthand
            kind, cnt = _one(eg["kind"]), int(_one(eg["n_clusters"]))
            relation = {"Crossed": {"n_clusters": cnt}} if kind == "crossed" else {"NestedWithin": {"n_per_parent": cnt}}
            s["cluster"]["extra_groupings"].append({"relation": relation, "tau_squared": float(_one(eg["tau_squared"])), "slopes": []})
    return s


def case_seed(case: dict) -> int:
    return int(_one(case["spec"]["seed"]))


def analytic_ols_power(beta: float, n: int, P: int, alpha: float = 0.05) -> float:
    """Two-sided t-test power of one slope under a random standard-normal design with unit error variance:
    non-central t with ncp = beta * sqrt(S), S = 1/[(X'X)^-1]_jj ~ chi2(n - P + 1), averaged over S
    (midpoint rule on 4000 chi-square quantiles)."""
    import numpy as np
    from scipy import stats
    df = n - P
    tc = stats.t.ppf(1 - alpha / 2, df)
    q = (np.arange(4000) + 0.5) / 4000
    ncp = beta * np.sqrt(stats.chi2.ppf(q, n - P + 1))
    return float(np.mean(1 - stats.nct.cdf(tc, df, ncp) + stats.nct.cdf(-tc, df, ncp)))


def posthoc_spec(method: str = "none") -> dict:
    """2 correlated continuous predictors + a 3-level and a 4-level factor (exact walk), every dummy a target, one
    posthoc block per factor (all pairwise level comparisons), posthoc alpha 0.01."""
    from mcpower_b200 import spec as S
    s = S.base_spec(2)
    s["correlation"] = [1.0, 0.3, 0.3, 1.0]
    s["n_factor_dummies"] = 2 + 3
    s["factor_n_levels"] = [3, 4]
    s["factor_proportions"] = [0.5, 0.3, 0.2, 0.4, 0.3, 0.2, 0.1]
    s["effect_sizes"] = [0.0, 0.25, 0.1, 0.5, -0.4, 0.3, 0.0, 0.6]
    s["target_indices"] = [1, 3, 4, 5, 6, 7]
    s["contrast_pairs"] = [[4, 3]]
    s["correction_method"] = method
    s["crit_values"] = {"alpha": 0.05, "posthoc_alpha": 0.01}
    s["posthoc"] = [{"factor_index": 0, "target_indices": [3, 4]}, {"factor_index": 1, "target_indices": [5, 6, 7]}]
    s["factor_min_level_count"] = 5
    s["report_overall"] = True
    return s


def crossing_goldens() -> dict:
    """End-to-end goldens of the shipped engine (validation/data/cross_*.golden.rds via tests/golden/make_crossing_goldens.py)."""
    return json.loads((GOLDEN / "crossing_goldens.json").read_text())


def crossing_spec(case: dict) -> dict:
    """Lowered SimulationSpec of a CROSS_CASES entry as the public API builds it (validation/common.R:60-72): standardized
    Normal predictors, every effect a target, alpha 0.05, no correction, optimistic scenario, overall test reported."""
    from mcpower_b200 import spec as S
    effects = list(case["effects"].values())
    s = S.base_spec(len(effects))
    s["effect_sizes"] = [0.0] + [float(e) for e in effects]
    s["target_indices"] = list(range(1, len(effects) + 1))
    s["report_overall"] = True
    if case["family"] == "logit":
        s["effect_sizes"][0] = S.logit_intercept(case["baseline_probability"])
        s["outcome_kind"] = "binary"
        s["estimator"] = "glm"
        s["max_failed_fraction"] = 0.25
    elif case["family"] == "lme":
        icc = case["cluster"]["ICC"]
        s["estimator"] = "mle"
        s["report_overall"] = False
        s["cluster"] = {"sizing": {"FixedClusters": {"n_clusters": case["cluster"]["n_clusters"]}},
                        "tau_squared": icc / (1.0 - icc), "slopes": [], "extra_groupings": []}
        s["max_failed_fraction"] = 0.25
    return s
