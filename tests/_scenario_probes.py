"""Numpy restatement of the reference's L5 scenario probes (validation/common.R:895-1075, validation_scenarios.rmd:38-260):
given a draw function (seed → design [n, P] f64, outcome [n] f64) it recomputes the statistics frozen in
validation/data/scenario_golden.rds.  Shared by the oracle (CPU) and CUDA (GPU) golden tests."""
import json
import math
from pathlib import Path

import numpy as np

from mcpower_b200 import spec as S

GOLDEN = json.loads((Path(__file__).resolve().parent / "golden" / "scenario_goldens.json").read_text())
SQRT3 = math.sqrt(3.0)


def spec_of(case: dict) -> dict:
    k = len(case["effects"])
    s = S.base_spec(k)
    s["effect_sizes"] = [0.0] + [float(e) for e in case["effects"]]
    s["target_indices"] = list(range(1, k + 1))
    if case.get("family") == "probit":
        from statistics import NormalDist
        s["effect_sizes"][0] = NormalDist().inv_cdf(case["baseline_probability"])
        s["outcome_kind"], s["link"], s["estimator"] = "binary", "probit", "glm"
    elif case.get("family") == "poisson":
        s["effect_sizes"][0] = math.log(case["baseline_rate"])
        s["outcome_kind"], s["estimator"] = "count", "glm"
    if "factor" in case:
        s = S.add_factor(s, case["factor"]["proportions"], case["factor"]["effects"])
    if "corr" in case:
        s["correlation"] = [1.0 if i == j else case["corr"] for j in range(k) for i in range(k)]
    if "residual" in case:  # set_residual_distribution PINS the residual (scenarios.rs:163-186)
        s["residual_dist"] = case["residual"]
        s["residual_pinned"] = True
    sc = S.scenario("optimistic")
    sc["name"] = "custom"
    sc.update(case["knobs"])
    s["scenario"] = sc
    return s


def sd(x):
    return float(np.std(np.asarray(x, dtype=np.float64), ddof=1))


def skew(x):
    m = x.mean()
    return float(((x - m) ** 3).mean() / np.std(x, ddof=1) ** 3)


def exkurt(x):
    m = x.mean()
    return float(((x - m) ** 4).mean() / np.std(x, ddof=1) ** 4 - 3.0)


def classify_marginal(x):  # common.R:977-984
    if np.abs(x).max() <= SQRT3 + 1e-9:
        return "uniform"
    sk = skew(x)
    if sk > 1.0:
        return "right_skewed"
    if sk < -1.0:
        return "left_skewed"
    if exkurt(x) > 1.0:
        return "high_kurtosis"
    return "normal"


def classify_residual(e):  # common.R:994-998
    if abs(skew(e)) > 0.45:
        return "skewed"
    if exkurt(e) > 0.5:
        return "heavy_tailed"
    return "normal"


def lstsq_coef(X, y):
    return np.linalg.lstsq(X, y, rcond=None)[0]


def statistics(name: str, draw) -> dict:
    """draw(seed) -> (design f64 [n, P] with the intercept in column 0, outcome f64 [n])."""
    case = GOLDEN[name]["case"]
    K, seed = case["K"], case["seed"]
    tb = np.array([0.0] + case["effects"])
    draws = (draw(seed + k) for k in range(1, K + 1))
    if name == "scen_fa_ols":  # run_fa: draw(seed, sampled) here; level counts from the dummy columns
        props, n = np.array(case["factor"]["proportions"]), case["n"]

        def counts(X):
            d = X[:, 2:2 + len(props) - 1].sum(axis=0)
            return np.concatenate([[n - d.sum()], d])

        c_fix = np.array([counts(draw(seed + k, False)[0]) for k in range(1, 21)])
        c_smp = np.array([counts(draw(seed + 500 + k, True)[0]) for k in range(1, K + 1)])
        out = {"fixed: counts identical across draws (no RNG)": float((c_fix == c_fix[0]).all()),
               "fixed: max |count − n·p| ≤ 1 (largest remainder)": float(np.abs(c_fix - n * props).max())}
        for g in range(len(props)):
            out[f"sampled: Var(count) level {g + 1} = n·p(1−p)"] = float(((c_smp[:, g] - n * props[g]) ** 2).mean())
        return out
    if name in ("scen_he_probit", "scen_he_poisson"):  # run_he_glm: mean fitted slope (oracle IRLS from the truth start)
        from oracle import orc
        fam = case["family"]
        b_true = np.array(spec_of(case)["effect_sizes"])
        B = np.array([orc.fit_glm(X, y, fam, beta_start=b_true)["betas"] for X, y in draws])
        return {"β̂[x1] = per-study pseudo-true (h=0.4)": B[:, 1].mean(), "β̂[x2] = per-study pseudo-true (h=0.4)": B[:, 2].mean()}
    if name == "scen_he":  # run_he: slope of squared true-beta residuals on the squared predictors
        sl = []
        for X, y in draws:
            e2 = (y - X @ tb) ** 2
            sl.append(lstsq_coef(np.column_stack([np.ones(len(y)), X[:, 1:] ** 2]), e2)[1:])
        sl = np.array(sl)
        return {"jitter slope on x1² = h²β²  (h=0.4)": sl[:, 0].mean(), "jitter slope on x2² = h²β²  (h=0.4)": sl[:, 1].mean()}
    if name == "scen_hs":  # run_hs: slope of log e² on the driver (x1)
        g = []
        for X, y in draws:
            e2 = (y - X @ tb) ** 2
            g.append(lstsq_coef(np.column_stack([np.ones(len(y)), X[:, 1]]), np.log(e2))[1])
        return {"log-e² slope γ = ln(λ)/4": float(np.mean(g))}
    if name in ("scen_co_low", "scen_co_high"):
        r = [np.corrcoef(X[:, 1], X[:, 2])[0, 1] for X, _ in draws]
        return {"block-r mean = ±0.8-clamp censored law": float(np.mean(r)), "block-r SD = √(s²/2 + (1−ρ²)²/n)": sd(r)}
    if name == "scen_co_psd":
        R = np.array([[np.corrcoef(X[:, 1], X[:, 2])[0, 1], np.corrcoef(X[:, 1], X[:, 3])[0, 1],
                       np.corrcoef(X[:, 2], X[:, 3])[0, 1]] for X, _ in draws])
        m = R.mean(axis=0)
        return {"PSD-repaired r ∈ [0.45, 0.60] (shrinks below ρ=0.6 input, stays positive)": float(m.min()),
                "r12_mean": m[0], "r13_mean": m[1], "r23_mean": m[2], "r12_sd": sd(R[:, 0]), "r13_sd": sd(R[:, 1]), "r23_sd": sd(R[:, 2])}
    if name == "scen_px":
        per_draw = [[(classify_marginal(X[:, j]), X[:, j].mean(), np.var(X[:, j], ddof=1)) for j in range(1, X.shape[1])]
                    for X, _ in draws]
        flat = [c for d in per_draw for c in d]
        cls = np.array([c[0] for c in flat])
        f_k = [np.mean([c[0] != "normal" for c in d]) for d in per_draw]
        out = {"swap frequency = q": float(np.mean(f_k))}
        for cnd in ("right_skewed", "left_skewed", "uniform"):
            out[f"pick share {cnd} = 1/3"] = float(np.mean(cls[cls != "normal"] == cnd))
        for cnd in ("right_skewed", "left_skewed", "uniform"):
            sel = [c for c in flat if c[0] == cnd]
            out[f"{cnd} mean = 0"] = float(np.mean([c[1] for c in sel]))
            out[f"{cnd} var = 1"] = float(np.mean([c[2] for c in sel]))
        return out
    if name == "scen_px_t3":
        res = np.array([[classify_marginal(X[:, 1]) == "high_kurtosis", X[:, 1].mean(), np.var(X[:, 1], ddof=1)] for X, _ in draws])
        return {"every block swapped (q = 1)": res[:, 0].mean(), "t3 mean = 0": res[:, 1].mean(),
                "t3 var = 1 (table-normalized)": res[:, 2].mean()}
    if name == "scen_re":
        res = []
        for X, y in draws:
            e = y - X @ tb
            res.append((classify_residual(e), skew(e), exkurt(e)))
        cls = np.array([r[0] for r in res])
        return {"swap frequency = q_r": float(np.mean(cls != "normal")),
                "pick share heavy_tailed = 1/2": float(np.mean(cls[cls != "normal"] == "heavy_tailed")),
                "pick share skewed = 1/2": float(np.mean(cls[cls != "normal"] == "skewed")),
                "skewed: skew = √(8/df), df = 10": float(np.mean([r[1] for r in res if r[0] == "skewed"])),
                "heavy: excess kurtosis = 6/(df−4), df = 10": float(np.mean([r[2] for r in res if r[0] == "heavy_tailed"]))}
    if name == "scen_re_replace":
        res = []
        for X, y in draws:
            e = y - X @ tb
            res.append((classify_residual(e) == "heavy_tailed", skew(e), e.mean(), np.var(e, ddof=1)))
        res = np.array(res, dtype=np.float64)
        return {"pin holds: residual stays high_kurtosis (right_skewed swap inert)": res[:, 0].mean(),
                "symmetric: skew = 0 (pinned t(6), not the χ² swap)": res[:, 1].mean(), "mean = 0": res[:, 2].mean(),
                "var = 1 (t(6) standardized)": res[:, 3].mean()}
    raise KeyError(name)
