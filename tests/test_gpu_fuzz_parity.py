"""Randomised differential parity: seeded random specs that combine the generator's and the estimators' features freely (correlated
block, marginals, exact / sampled factors, interactions, contrasts, posthoc blocks, corrections, scenario knobs, the three GLM
families, random intercepts / slopes / cluster-level predictors, reduced test designs) are run through the oracle and through the
CUDA path; every output byte and every statistic word must agree (the single-feature tests in test_gpu_parity.py pin the features
one at a time — this file pins their combinations, which is where a lane-dependent or order-dependent defect would hide).

The spec generator is deterministic (numpy Generator, fixed seeds), so a failure names a reproducible case; the CPU half of the file
checks that every generated spec is accepted by the oracle and by the library's own validator (no GPU needed)."""
import copy

import numpy as np
import pytest

import mcpower_b200 as M
from mcpower_b200 import spec as S
from oracle import orc

ALL = M.MODE_COUNTS | M.MODE_BYTES | M.MODE_CAPTURE
N_CASES = 72
MARGINALS = ["normal", "binary", "right_skewed", "left_skewed", "high_kurtosis", "uniform"]


def random_spec(case: int):
    """→ (spec, sample sizes, n_sims, seed, label)"""
    g = np.random.default_rng(0xF022 + 7919 * case)
    kind = ["ols", "ols", "ols", "glm", "mle"][case % 5]
    k = int(g.integers(1, 7))
    s = S.base_spec(k)
    # correlated continuous block: AR(1)-like with a random rho, sometimes an arbitrary PSD matrix from a random factor
    if k > 1 and g.random() < 0.7:
        if g.random() < 0.5:
            rho = float(g.uniform(-0.5, 0.8))
            c = np.array([[rho ** abs(i - j) for j in range(k)] for i in range(k)])
        else:
            a = g.normal(size=(k, k + 2))
            c = a @ a.T
            d = np.sqrt(np.diag(c))
            c = c / d[:, None] / d[None, :]
            c = np.clip(c, -0.85, 0.85)
            np.fill_diagonal(c, 1.0)
            if np.linalg.eigvalsh(c).min() < 0.05:
                c = 0.5 * c + 0.5 * np.eye(k)
        s["correlation"] = [float(v) for v in c.T.reshape(-1)]
    for j in range(k):
        if g.random() < 0.35:
            s["var_types"][j] = MARGINALS[int(g.integers(1, len(MARGINALS)))]
            if s["var_types"][j] == "binary":
                s["var_params"][j] = float(g.uniform(0.15, 0.85))
    s["effect_sizes"] = [float(g.uniform(-0.3, 0.3))] + [float(g.choice([0.0, 0.15, 0.3, -0.25, 0.5])) for _ in range(k)]
    # factors (exact walk or sampled), then interactions
    n_factors = int(g.integers(0, 3)) if kind != "mle" or g.random() < 0.5 else 0
    for _ in range(n_factors):
        nl = int(g.integers(2, 5))
        props = g.dirichlet(np.full(nl, 3.0))
        s = S.add_factor(s, props, [float(g.choice([0.0, 0.3, -0.4, 0.6])) for _ in range(nl - 1)])
    if n_factors:
        s["factor_sampled"] = [None if g.random() < 0.5 else bool(g.random() < 0.5) for _ in range(n_factors)]
        s["factor_min_level_count"] = int(g.choice([0, 3, 5]))
    p_main = 1 + k + s["n_factor_dummies"]
    for _ in range(int(g.integers(0, 3))):
        if p_main > 2:
            a, b = (int(v) for v in g.choice(np.arange(1, p_main), size=2, replace=False))
            s = S.add_interaction(s, [a, b], float(g.choice([0.0, 0.2, -0.15])))
    P = S.n_predictors(s)
    nt = int(g.integers(1, min(P - 1, 5) + 1))
    s["target_indices"] = sorted(int(v) for v in g.choice(np.arange(1, P), size=nt, replace=False))
    if P > 3 and g.random() < 0.4:
        a, b = (int(v) for v in g.choice(np.arange(1, P), size=2, replace=False))
        s["contrast_pairs"] = [[a, b]]
    s["correction_method"] = str(g.choice(["none", "bonferroni", "holm", "benjamini_hochberg"]))
    s["crit_values"] = {"alpha": float(g.choice([0.05, 0.01, 0.2, 5e-7])), "posthoc_alpha": None}
    s["report_overall"] = bool(g.random() < 0.5)
    s["max_failed_fraction"] = 1.0
    if kind == "ols" and n_factors and g.random() < 0.5:  # posthoc blocks over a factor whose dummies are all targets
        f = int(g.integers(0, n_factors))
        first = 1 + k + sum(nl - 1 for nl in s["factor_n_levels"][:f])
        dummies = list(range(first, first + s["factor_n_levels"][f] - 1))
        s["target_indices"] = sorted(set(s["target_indices"]) | set(dummies))
        s["posthoc"] = [{"factor_index": f, "target_indices": dummies}]
        s["crit_values"]["posthoc_alpha"] = float(g.choice([0.05, 0.01]))
        if g.random() < 0.3:
            s["correction_method"] = "tukey_hsd"
    if kind == "ols" and g.random() < 0.25 and P > 3:  # reduced test design (fit_columns)
        keep = sorted(set([0] + s["target_indices"] + [int(v) for v in g.choice(np.arange(1, P), size=2)]))
        if s["contrast_pairs"]:
            keep = sorted(set(keep) | set(s["contrast_pairs"][0]))
        if not s["posthoc"]:
            s["fit_columns"] = keep
    # scenario
    preset = str(g.choice(["optimistic", "optimistic", "realistic", "doomer", "custom"]))
    mixed = kind == "mle"
    if preset == "custom":
        sc = S.scenario("realistic", mixed=mixed)
        sc.update(name="custom", heterogeneity=float(g.choice([0.0, 0.3])), heteroskedasticity_ratio=float(g.choice([1.0, 3.0])),
                  correlation_noise_sd=float(g.choice([0.0, 0.25])), distribution_change_prob=float(g.choice([0.0, 0.4, 1.0])),
                  residual_change_prob=float(g.choice([0.0, 0.6])), residual_df=float(g.choice([4.0, 9.0])),
                  sampled_factor_proportions=bool(g.random() < 0.5), truth_start=bool(g.random() < 0.5))
        s["scenario"] = sc
    else:
        s["scenario"] = S.scenario(preset, mixed=mixed)
    if kind == "ols" and g.random() < 0.3:
        s["residual_dist"] = str(g.choice(["high_kurtosis", "right_skewed", "left_skewed", "uniform"]))
        s["scenario"]["residual_df"] = float(g.choice([5.0, 8.0]))
        s["residual_pinned"] = bool(g.random() < 0.5)
    grid = sorted(int(v) for v in g.choice(np.arange(30, 400), size=int(g.integers(1, 5)), replace=False))
    if kind == "glm":
        fam = str(g.choice(["logit", "probit", "poisson"]))
        s["estimator"] = "glm"
        s["effect_sizes"] = [float(v) * 0.6 for v in s["effect_sizes"]]
        if fam == "poisson":
            s["outcome_kind"], s["link"] = "count", None
            s["intercept"] = float(g.uniform(0.0, 1.2))
        else:
            s["outcome_kind"], s["link"] = "binary", ("probit" if fam == "probit" else None)
            s["intercept"] = float(g.uniform(-1.0, 1.0))
        s["effect_sizes"][0] = s["intercept"]
        s["posthoc"], s["contrast_pairs"], s["fit_columns"] = [], [], []
        grid = [max(n, 60) for n in grid]
    if kind == "mle":
        s["estimator"] = "mle"
        s["posthoc"], s["contrast_pairs"], s["fit_columns"] = [], [], []
        if g.random() < 0.5:
            sizing = {"FixedClusters": {"n_clusters": int(g.integers(4, 25))}}
        else:
            sizing = {"FixedSize": {"cluster_size": int(g.integers(4, 15))}}
        s["cluster"] = {"sizing": sizing, "tau_squared": float(g.uniform(0.05, 0.6)), "slopes": [], "extra_groupings": []}
        n_sl = int(g.choice([0, 0, 1, 2])) if k >= 2 else int(g.choice([0, 1]))
        cols = [int(v) for v in g.choice(np.arange(k), size=n_sl, replace=False)] if n_sl else []
        for a, c in enumerate(cols):
            s["cluster"]["slopes"].append({"column": c, "variance": float(g.uniform(0.03, 0.3)),
                                           "corr_with_intercept": float(g.uniform(-0.5, 0.5)),
                                           "corr_with": [float(g.uniform(-0.4, 0.4))] if a == 1 else []})
            s["cluster_slope_design_cols"].append(1 + c)
        if g.random() < 0.3 and k >= 2:
            s["between_var_indices"] = [int(v) for v in g.choice(np.arange(1, 1 + k), size=1)]
        if g.random() < 0.5:  # extra groupings: crossed needs FixedClusters, one nested at most, cluster_size % n_per_parent == 0
            # (few levels: the oracle's dense Cholesky is cubic in the number of random-effect columns)
            if "FixedSize" in sizing:
                sizing["FixedSize"]["cluster_size"] = int(g.integers(24, 40))
            else:
                sizing["FixedClusters"]["n_clusters"] = int(g.integers(4, 13))
            kinds = ["nested"] if "FixedSize" in sizing else [str(v) for v in g.choice(["crossed", "nested"], size=int(g.integers(1, 3)))]
            if kinds.count("nested") > 1:
                kinds = ["crossed", "nested"]
            for rel_kind in kinds:
                if rel_kind == "crossed":
                    rel = {"Crossed": {"n_clusters": int(g.integers(2, 7))}}
                else:
                    npp = int(g.integers(2, 4))
                    if "FixedSize" in sizing:
                        sizing["FixedSize"]["cluster_size"] = max(npp, (sizing["FixedSize"]["cluster_size"] // npp) * npp)
                    rel = {"NestedWithin": {"n_per_parent": npp}}
                s["cluster"]["extra_groupings"].append({"relation": rel, "tau_squared": float(g.uniform(0.0, 0.3)), "slopes": []})
        if s["scenario"]["lme"] is not None and g.random() < 0.5:
            s["scenario"]["lme"] = {"random_effect_dist": str(g.choice(["normal", "high_kurtosis", "right_skewed", "uniform"])),
                                    "random_effect_df": float(g.choice([4.0, 7.0])), "icc_noise_sd": float(g.choice([0.0, 0.2]))}
        grid = [max(n, 90) for n in grid]
    n_sims = 96 if kind != "mle" else 64
    if kind == "mle" and s["cluster"]["extra_groupings"]:
        grid, n_sims = grid[-2:], 32
    return s, sorted(set(grid)), n_sims, int(g.integers(1, 2**62)), f"case{case}:{kind}:{preset}:P{P}:f{n_factors}"


CASES = list(range(N_CASES))


@pytest.mark.parametrize("case", CASES[::3])
def test_generated_specs_are_valid_for_the_oracle(case):
    s, grid, n_sims, seed, label = random_spec(case)
    r = orc.run_batch(s, grid, 8, seed, n_threads=1)
    assert r["converged"].shape == (8, len(grid)), label


def test_the_generator_covers_every_arm():
    kinds, presets, slopes, posthoc, fitcols, between = set(), set(), set(), 0, 0, 0
    for case in CASES:
        s, *_rest, label = random_spec(case)
        kinds.add((s["estimator"], s["outcome_kind"], s["link"]))
        presets.add(s["scenario"]["name"])
        if s["cluster"]:
            slopes.add(len(s["cluster"]["slopes"]))
        posthoc += bool(s["posthoc"])
        fitcols += bool(s["fit_columns"])
        between += bool(s["between_var_indices"])
    assert {("ols", "continuous", None), ("glm", "binary", None), ("glm", "binary", "probit"), ("glm", "count", None),
            ("mle", "continuous", None)} <= kinds
    assert presets == {"optimistic", "realistic", "doomer", "custom"}
    assert slopes == {0, 1, 2} and posthoc >= 3 and fitcols >= 2 and between >= 2
    groupings = {tuple("N" if "NestedWithin" in e["relation"] else "C" for e in random_spec(c)[0]["cluster"]["extra_groupings"])
                 for c in CASES if random_spec(c)[0]["cluster"]}
    assert {(), ("C",), ("N",)} <= groupings and any(len(t) == 2 for t in groupings)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_gpu_equals_oracle_on_a_random_spec(case):
    s, grid, n_sims, seed, label = random_spec(case)
    want = orc.run_batch(s, grid, n_sims, seed, n_threads=orc.host_threads(), capture=True)
    variants = (0, 1, 4, 6) if s["estimator"] == "ols" else (0,)
    for variant in variants:  # the library's own kernel choice, then every OLS kernel family that accepts the spec
        previous = M.force_ols_variant(variant)
        try:
            got = M.run_batch(s, grid, n_sims, seed, mode=ALL)
        finally:
            M.force_ols_variant(previous)
        tag = f"{label} (kernel variant {variant})"
        for key in ("uncorrected", "corrected", "converged", "overall", "factor_excluded", "joint", "boundary_hit", "posthoc_unc", "posthoc_cor"):
            w = want.get(key)
            if w is None:
                continue
            assert np.array_equal(getattr(got, key), w), (tag, key)
        assert np.array_equal(got.t_sq, want["t_sq"], equal_nan=True), (tag, "t_sq")
        finite = np.isfinite(want["t_sq"])
        assert np.array_equal(got.t_sq[finite].view(np.uint64), want["t_sq"][finite].view(np.uint64)), (tag, "t_sq bits")
        agg = M.counters_from_bytes(got.uncorrected, got.corrected, got.converged, got.overall, got.factor_excluded, got.joint,
                                    got.boundary_hit, got.posthoc_unc, got.posthoc_cor)
        assert np.array_equal(agg, got.counters), (tag, "COUNTS != aggregate(BYTES)")
    # the generated design itself, for two studies
    for sim in (0, n_sims - 1):
        xg, yg = M.generate_sim(s, seed, sim, grid[-1])
        xo, yo, _ = orc.generate_sim(s, seed, sim, grid[-1])
        assert np.array_equal(xg.view(np.uint32), xo.view(np.uint32)) and np.array_equal(yg.view(np.uint32), yo.view(np.uint32)), (label, sim)
