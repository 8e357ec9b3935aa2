"""Extra grouping factors: crossed and nested random intercepts next to the primary grouping (which may carry its random slopes).
Reference: `GroupingRelation` / `extra_level_of_row` / `extra_n_levels_at` (engine-contract generation.rs:243-430), the draw blocks
behind the primary's on the sequential stream and `y += u_g[level]` (data_gen.rs:214-236, 818-835), one ICC-jitter draw per grouping
(data_gen.rs:1077-1091), the reduced fit keeps every grouping's intercept and strips the slopes (batch.rs:1357-1375), the fit itself
is glmm::fit_on (not in the tree) — restated as lme4's penalised least squares on dense matrices (oracle) and with a block-diagonal
primary part (CUDA), the same fma chains.

* CPU: the oracle reproduces the reference's six saved datasets with extra groupings bit for bit (tests/test_oracle_golden.py walks
  every fixture), their level vectors, and lme4::lmer's fits within validation/tolerances.R's gates.
* GPU: generated data, every run_batch output byte and the statistic words equal the oracle's."""
import copy

import numpy as np
import pytest

import _cases as CS
import mcpower_b200 as M
from mcpower_b200 import spec as S
from oracle import orc

SEED = 0x600D5EED0C0FFEE5
ALL = M.MODE_COUNTS | M.MODE_BYTES | M.MODE_CAPTURE
GROUPED = sorted(n for n in CS.validation_cases() if n.startswith(("lmm_crossed", "lmm_nested", "lmm_slope_crossed")))


def _rel(a, b):
    return np.max(np.abs(np.asarray(a) - np.asarray(b)) / np.abs(np.asarray(b)))


def grouped_spec(n_clusters=12, icc=0.2, extras=(("crossed", 6, 0.15),), slopes=0, preset="optimistic", k=2, fixed_size=None):
    s = S.base_spec(k)
    s["scenario"] = S.scenario(preset, mixed=True)
    s["effect_sizes"] = [0.0] + [0.3, 0.2, 0.15, 0.25][:k]
    s["target_indices"] = list(range(1, k + 1))
    s["estimator"] = "mle"
    sizing = {"FixedClusters": {"n_clusters": n_clusters}} if fixed_size is None else {"FixedSize": {"cluster_size": fixed_size}}
    s["cluster"] = {"sizing": sizing, "tau_squared": icc / (1.0 - icc), "slopes": [], "extra_groupings": []}
    for kind, cnt, tau2 in extras:
        rel = {"Crossed": {"n_clusters": cnt}} if kind == "crossed" else {"NestedWithin": {"n_per_parent": cnt}}
        s["cluster"]["extra_groupings"].append({"relation": rel, "tau_squared": tau2, "slopes": []})
    for a in range(slopes):
        s["cluster"]["slopes"].append({"column": a, "variance": [0.1, 0.08][a], "corr_with_intercept": [0.3, 0.1][a],
                                       "corr_with": [0.2] if a == 1 else []})
        s["cluster_slope_design_cols"].append(1 + a)
    s["max_failed_fraction"] = 1.0
    return s


# --------------------------------------------------------------------------------------------------------------- CPU
def test_fixture_set():
    assert GROUPED == ["lmm_crossed_a", "lmm_crossed_b", "lmm_crossed_nested_a", "lmm_nested_a", "lmm_nested_b", "lmm_slope_crossed"]


@pytest.mark.parametrize("name", GROUPED)
def test_level_vectors_match_the_reference(name):  # generation.rs:386-430 (the pinned within-block ordering)
    case = CS.validation_cases()[name]
    spec = CS.case_spec(case)
    cl = spec["cluster"]
    nc = cl["sizing"]["FixedClusters"]["n_clusters"]
    assert list(np.arange(case["n"]) % nc) == [c - min(case["cluster_ids"]) for c in case["cluster_ids"]]
    for g, want in enumerate(case["extra_grouping_ids"]):
        got = [orc.extra_level_of_row(spec, g, i) for i in range(case["n"])]
        assert got == [w - min(want) for w in want], (name, g)
        assert orc.extra_levels_at(spec, g, case["n"]) == max(got) + 1


@pytest.mark.parametrize("name", GROUPED)
def test_oracle_reml_matches_lmer(name):
    case = CS.validation_cases()[name]
    spec = CS.case_spec(case)
    x, y = CS.case_arrays(case)
    g = CS.golden_arrays(case)
    ids = np.asarray(case["cluster_ids"]) - min(case["cluster_ids"])
    extra = [np.asarray(e) - min(e) for e in case["extra_grouping_ids"]]
    fit = orc.fit_lmm_general(x.astype(np.float64), y.astype(np.float64), ids, spec["cluster_slope_design_cols"], extra)
    assert fit["converged"] and fit["boundary"] == 0
    assert _rel(fit["betas"], g["beta"]) < 1e-4                       # tolerances.R: estimate_rel_iter
    assert _rel(fit["betas"] / fit["se"], g["z"]) < 1e-4
    assert _rel(np.concatenate([fit["variances"], fit["extra_variances"]]), g["vars"]) < 1e-3  # lmer's order: primary terms, then extras
    assert _rel(fit["sigma_sq"], g["sigma2"][0]) < 1e-3
    if len(g["corr"]):
        assert np.max(np.abs(fit["corr"] - g["corr"])) < 2e-3


def test_general_fit_without_extras_is_the_slope_fit():
    case = CS.validation_cases()["lmm_slope_b"]
    spec = CS.case_spec(case)
    x, y = CS.case_arrays(case)
    ids = np.asarray(case["cluster_ids"]) - min(case["cluster_ids"])
    a = orc.fit_lmm_general(x.astype(np.float64), y.astype(np.float64), ids, spec["cluster_slope_design_cols"], [])
    b = orc.fit_lmm_slopes(x.astype(np.float64), y.astype(np.float64), ids, spec["cluster_slope_design_cols"])
    # (the two forms round differently — fma chains here, mul / sub there — so the simplex paths part after a while)
    assert abs(a["n_eval"] - b["n_eval"]) < 40
    assert np.allclose(a["betas"], b["betas"], rtol=1e-6) and np.allclose(a["variances"], b["variances"], rtol=1e-4)


def test_validation_of_extra_groupings():  # validate.rs:720-745
    ok = grouped_spec()
    orc.run_batch(ok, [144], 2, SEED)
    bad = grouped_spec(fixed_size=12)  # crossed needs FixedClusters
    with pytest.raises(ValueError):
        orc.run_batch(bad, [144], 2, SEED)
    bad = grouped_spec(extras=(("nested", 2, 0.1), ("nested", 3, 0.1)))
    with pytest.raises(ValueError):
        orc.run_batch(bad, [144], 2, SEED)
    bad = grouped_spec(extras=(("nested", 5, 0.1),), fixed_size=12)  # 12 % 5 != 0
    with pytest.raises(ValueError):
        orc.run_batch(bad, [144], 2, SEED)
    three = grouped_spec(extras=(("crossed", 3, 0.1), ("crossed", 2, 0.1), ("nested", 2, 0.1)))
    with pytest.raises(M.Unsupported):
        M.run_batch(three, [144], 2, SEED)
    sl = grouped_spec()
    sl["cluster"]["extra_groupings"][0]["slopes"] = [{"column": 0, "variance": 0.05, "corr_with_intercept": 0.0, "corr_with": []}]
    sl["extra_slope_cols"] = [[1]]
    with pytest.raises(M.Unsupported):
        M.run_batch(sl, [144], 2, SEED)


def test_oracle_crossed_batch_is_calibrated():
    s = grouped_spec(20, 0.2, extras=(("crossed", 12, 0.15),))
    s["effect_sizes"] = [0.0, 0.0, 0.25]
    r = orc.run_batch(s, [480], 600, SEED, n_threads=orc.host_threads())
    conv = r["converged"][:, 0].astype(bool)
    assert conv.mean() > 0.97
    assert 0.02 < r["uncorrected"][conv, 0, 0].mean() < 0.09
    assert r["uncorrected"][conv, 0, 1].mean() > 0.95


# --------------------------------------------------------------------------------------------------------------- GPU
def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.gpu
@pytest.mark.parametrize("name", GROUPED)
def test_gpu_generator_reproduces_the_reference_datasets(name):
    case = CS.validation_cases()[name]
    spec = CS.case_spec(case)
    xg, yg = M.generate_sim(spec, orc.splitmix64(CS.case_seed(case)), 0, case["n"])
    xr, yr = CS.case_arrays(case)
    assert np.array_equal(_bits(xg), _bits(xr)) and np.array_equal(_bits(yg), _bits(yr))


LAYOUTS = {
    "crossed": dict(n_clusters=10, extras=(("crossed", 6, 0.15),)),
    "nested": dict(n_clusters=8, extras=(("nested", 3, 0.1),)),
    "nested_fixed_size": dict(fixed_size=12, extras=(("nested", 3, 0.1),)),
    "crossed_and_nested": dict(n_clusters=6, extras=(("crossed", 4, 0.15), ("nested", 2, 0.08))),
    "slope_and_crossed": dict(n_clusters=12, extras=(("crossed", 5, 0.16),), slopes=1),
    "two_slopes_and_nested_doomer": dict(n_clusters=10, extras=(("nested", 2, 0.1),), slopes=2, preset="doomer"),
    "crossed_realistic_p4": dict(n_clusters=9, extras=(("crossed", 4, 0.2),), preset="realistic", k=3),
}


@pytest.mark.gpu
@pytest.mark.parametrize("layout", sorted(LAYOUTS))
def test_gpu_generated_outcome_is_bit_exact(layout):
    s = grouped_spec(**LAYOUTS[layout])
    for sim in (0, 7, 424242):
        xg, yg = M.generate_sim(s, SEED, sim, 250)
        xo, yo, _ = orc.generate_sim(s, SEED, sim, 250)
        assert np.array_equal(_bits(xg), _bits(xo)) and np.array_equal(_bits(yg), _bits(yo)), (layout, sim)


@pytest.mark.gpu
@pytest.mark.parametrize("layout", sorted(LAYOUTS))
def test_gpu_run_batch_matches_the_oracle(layout):
    s = grouped_spec(**LAYOUTS[layout])
    grid = [120, 240] if "fixed_size" not in layout else [96, 250]
    n = 96
    got = M.run_batch(s, grid, n, SEED, mode=ALL)
    want = orc.run_batch(s, grid, n, SEED, n_threads=orc.host_threads(), capture=True)
    for key in ("uncorrected", "corrected", "converged", "joint", "boundary_hit"):
        assert np.array_equal(getattr(got, key), want[key]), (layout, key)
    assert np.array_equal(got.t_sq, want["t_sq"], equal_nan=True)
    fin = np.isfinite(want["t_sq"])
    assert np.array_equal(got.t_sq[fin].view(np.uint64), want["t_sq"][fin].view(np.uint64))
    agg = M.counters_from_bytes(got.uncorrected, got.corrected, got.converged, got.overall, got.factor_excluded, got.joint,
                                got.boundary_hit)
    assert np.array_equal(agg, got.counters)
    assert got.n_converged.min() > 0.8 * n
    part = M.run_batch(s, grid, n - 40, SEED, mode=ALL, sim_begin=40, wave_sims=128)
    assert np.array_equal(part.uncorrected, got.uncorrected[40:])


@pytest.mark.gpu
def test_gpu_factor_exclusion_with_extra_groupings_matches_the_oracle():
    """A sparse factor level triggers the reduced fit: the slopes are stripped, the groupings' intercepts stay."""
    s = grouped_spec(n_clusters=12, extras=(("crossed", 4, 0.16),), slopes=1)
    s = S.add_factor(s, [0.85, 0.15], [0.3])
    s["factor_min_level_count"] = 12
    s["scenario"]["sampled_factor_proportions"] = True
    grid = [48, 96, 240]
    got = M.run_batch(s, grid, 96, SEED, mode=ALL)
    want = orc.run_batch(s, grid, 96, SEED, n_threads=orc.host_threads(), capture=True)
    assert want["factor_excluded"][:, 0].any() and not want["factor_excluded"][:, 2].all()
    for key in ("uncorrected", "corrected", "converged", "joint", "boundary_hit", "factor_excluded"):
        assert np.array_equal(getattr(got, key), want[key]), key
    assert np.array_equal(got.t_sq, want["t_sq"], equal_nan=True)


@pytest.mark.gpu
@pytest.mark.parametrize("name", GROUPED)
def test_gpu_reml_on_reference_bytes_matches_lmer_and_the_oracle(name):
    """fit_provided_data (introspect.rs:174-562) through the CUDA fit on the reference's saved bytes: lme4's answers within the
    reference's own gates, and the oracle's doubles exactly."""
    case = CS.validation_cases()[name]
    spec = CS.case_spec(case)
    x, y = CS.case_arrays(case)
    g = CS.golden_arrays(case)
    ids = (np.asarray(case["cluster_ids"]) - min(case["cluster_ids"])).astype(np.uint32)
    got = M.fit_provided_data(spec, x.astype(np.float64), y.astype(np.float64), cluster_ids=ids)
    assert got["converged"] and got["boundary_hit"] == 0
    assert _rel(got["betas"], g["beta"]) < 1e-4
    assert _rel(got["statistic"], np.abs(g["z"][1:])) < 1e-4
    ns, ne = len(spec["cluster_slope_design_cols"]), len(spec["cluster"]["extra_groupings"])
    assert _rel([got["tau_sq_hat"]] + got["slope_var_hat"][:ns] + got["extra_var_hat"][:ne], g["vars"]) < 1e-3
    assert _rel(got["sigma_sq_hat"], g["sigma2"][0]) < 1e-3
    if ns:
        assert abs(got["slope_corr_hat"][0] - g["corr"][0]) < 2e-3
    extra = [np.asarray(e) - min(e) for e in case["extra_grouping_ids"]]
    want = orc.fit_lmm_general(x.astype(np.float64), y.astype(np.float64), ids.astype(np.int32), spec["cluster_slope_design_cols"], extra)
    assert np.array_equal(got["betas"], want["betas"]) and np.array_equal(got["se"], want["se"][1:])
    assert got["sigma_sq_hat"] == want["sigma_sq"] and got["tau_sq_hat"] == want["variances"][0]
    assert got["extra_var_hat"][:ne] == list(want["extra_variances"])


def _grouped_contract():
    import test_contract_lowering as TC
    cluster = {"sizing": {"FixedClusters": {"n_clusters": 10}}, "tau_squared": 0.25,
               "slopes": [{"column": 0, "variance": 0.1, "corr_with_intercept": 0.3, "corr_with": []}],
               "extra_groupings": [{"relation": {"Crossed": {"n_clusters": 6}}, "tau_squared": 0.15, "slopes": []},
                                   {"relation": {"NestedWithin": {"n_per_parent": 2}}, "tau_squared": 0.08, "slopes": []}]}
    return TC.contract([TC.normal(), TC.normal()], [TC.CONST, TC.direct(0), TC.direct(1)], [0.0, 0.3, 0.2],
                       [TC.marginal(1), TC.marginal(2)], estimator="mle", cluster=cluster)


def test_contract_with_extra_groupings_lowers_and_runs_on_the_oracle():
    import msgpack
    low = M.lower_contracts(msgpack.packb([_grouped_contract()], use_bin_type=True))[0]
    eg = low["spec"]["cluster"]["extra_groupings"]
    assert [list(e["relation"]) for e in eg] == [["Crossed"], ["NestedWithin"]]
    assert low["spec"]["cluster_slope_design_cols"] == [1] and low["spec"]["extra_slope_cols"] == [[], []]
    r = orc.run_batch(low["spec"], [240], 8, 5)
    assert r["converged"].shape == (8, 1)


@pytest.mark.gpu
def test_gpu_contract_with_extra_groupings_runs_end_to_end():
    """The hosts' payload for `y ~ x1 + x2 + (1 + x1 | g) + (1 | item) + (1 | g:half)` through the contract entry point."""
    import msgpack
    blob = msgpack.packb([_grouped_contract()], use_bin_type=True)
    low = M.lower_contracts(blob)[0]
    got = M.find_power_contracts(blob, 240, 256, 2137)
    pr = got["scenarios"][0][1]
    want = orc.run_batch(low["spec"], [240], 256, M.splitmix64(2137), n_threads=orc.host_threads())
    assert pr["success_counts_uncorrected"] == [int(v) for v in want["uncorrected"].sum(axis=0)[0]]
    assert pr["convergence_count"] == int(want["converged"].sum())
    ex = pr["estimator_extras"]  # four variance components: intercept, slope, item, half (result.rs:362-398)
    conv = want["converged"][:, 0] != 0
    pins = want["boundary_hit"][:, 0] >> 2
    assert ex["boundary_component_counts"] == [int((((pins >> k) & 1) != 0)[conv].sum()) for k in range(4)]
    assert sum(ex["boundary_component_counts"]) > 0
