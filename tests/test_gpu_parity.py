"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs, against the committed reference fixtures, and — at full size — through
size-independent properties.  Bar: bit-exact for every integer/byte output and for the f32 data
plane; the f64 statistics are computed in the same operation order as the oracle, so they are
compared bit-for-bit too (tolerance 0), with the 1e-10 relative bar of BASELINE.json's north_star
checked against the frozen R goldens."""
import math

import numpy as np
import pytest

import _cases as CS
import mcpower_b200 as M
from mcpower_b200 import spec as S
from oracle import orc

pytestmark = pytest.mark.gpu

SEED = 0x9E3779B97F4A7C15

# All OLS power kernels (one thread per study / one lane quad per study / the lean quad + solve pair for plain linear
# models, which falls back to the general quad kernel where a spec is outside its family / the general quad
# generate + accumulate kernel with the separate solve kernel) must produce the same bits
# on the same inputs: every run_batch test below runs once per forced kernel choice.  Tests that never reach the OLS power
# kernel (generated-design dumps, supplied-bytes fits, GLM / LMM arms, host-side errors) run once.
_NOT_OLS_POWER = ("cross_lme", "cross_glm", "generated_design", "marginal_transforms", "pinned_residual", "sampled_factors_and_cluster",
                  "prefix_stability", "fit_on_reference", "design_draw", "logit", "probit", "glm", "lmm", "errors", "contract")  # substrings of test names


@pytest.fixture(autouse=True, params=["thread", "quad", "lean", "split"])
def ols_kernel_choice(request):
    if request.param != "thread" and any(t in request.node.name for t in _NOT_OLS_POWER):
        pytest.skip("does not reach the OLS power kernel")
    previous = M.force_ols_variant({"thread": 1, "quad": 4, "lean": 5, "split": 6}[request.param])
    yield request.param
    M.force_ols_variant(previous)


def assert_batch_equal(got: M.BatchResult, want: dict):
    for key in ("uncorrected", "corrected", "converged"):
        assert np.array_equal(getattr(got, key), want[key]), key
    if want["overall"] is not None:
        assert np.array_equal(got.overall, want["overall"]), "overall"
    if want["factor_excluded"] is not None:
        assert np.array_equal(got.factor_excluded, want["factor_excluded"]), "factor_excluded"
    if want["t_sq"] is not None and got.t_sq is not None:
        assert np.array_equal(got.t_sq.view(np.uint64), want["t_sq"].view(np.uint64)) or \
            np.array_equal(got.t_sq, want["t_sq"], equal_nan=True), "t_sq bits"
    if want.get("joint") is not None:
        assert np.array_equal(got.joint, want["joint"]), "joint"
        assert np.array_equal(got.boundary_hit, want["boundary_hit"]), "boundary_hit"
    if want.get("posthoc_unc") is not None:
        assert np.array_equal(got.posthoc_unc, want["posthoc_unc"]), "posthoc_unc"
        assert np.array_equal(got.posthoc_cor, want["posthoc_cor"]), "posthoc_cor"
    agg = M.counters_from_bytes(got.uncorrected, got.corrected, got.converged, got.overall, got.factor_excluded,
                                got.joint, got.boundary_hit, got.posthoc_unc, got.posthoc_cor)
    assert np.array_equal(agg, got.counters), "COUNTS != aggregate(BYTES)"


ALL = M.MODE_COUNTS | M.MODE_BYTES | M.MODE_CAPTURE

_ORACLE_CACHE = {}


def oracle_once(key, fn):
    """The oracle result of a case does not depend on which GPU kernel is under test: compute it once per session."""
    if key not in _ORACLE_CACHE:
        _ORACLE_CACHE[key] = fn()
    return _ORACLE_CACHE[key]


# ------------------------------------------------------------------ data plane vs the reference's saved datasets
@pytest.mark.parametrize("name", sorted(CS.validation_cases()))
def test_generated_design_bit_exact_vs_reference_fixture(name):
    """sim 0 of base_seed = splitmix64(seed): X and y equal the reference engine's saved bytes."""
    case = CS.validation_cases()[name]
    spec = CS.case_spec(case)
    x, y = M.generate_sim(spec, M.splitmix64(CS.case_seed(case)), 0, case["n"])
    xr, yr = CS.case_arrays(case)
    assert np.array_equal(x.view(np.uint32), xr.view(np.uint32)), "design bits"
    assert np.array_equal(y.view(np.uint32), yr.view(np.uint32)), "outcome bits"


@pytest.mark.parametrize("dist", ["binary", "right_skewed", "left_skewed", "high_kurtosis", "uniform"])
def test_marginal_transforms_bit_exact_vs_oracle(dist):
    s = S.base_spec(3)
    s["correlation"] = [1, 0.4, 0.2, 0.4, 1, 0.3, 0.2, 0.3, 1]
    s["var_types"] = ["normal", dist, dist]
    s["var_params"] = [0.0, 0.3, 0.7]
    s["effect_sizes"] = [0.1, 0.2, -0.3, 0.4]
    for sim in (0, 7):
        x, y = M.generate_sim(s, SEED, sim, 257)
        xo, yo, _ = orc.generate_sim(s, SEED, sim, 257)
        assert np.array_equal(x.view(np.uint32), xo.view(np.uint32))
        assert np.array_equal(y.view(np.uint32), yo.view(np.uint32))


@pytest.mark.parametrize("resid", ["uniform", "high_kurtosis", "right_skewed", "left_skewed"])
def test_pinned_residual_dists_bit_exact_vs_oracle(resid):
    s = S.headline_spec()
    s["residual_dist"], s["residual_pinned"] = resid, True
    s["scenario"]["residual_df"] = 5.0
    x, y = M.generate_sim(s, SEED, 11, 130)
    xo, yo, _ = orc.generate_sim(s, SEED, 11, 130)
    assert np.array_equal(x.view(np.uint32), xo.view(np.uint32))
    assert np.array_equal(y.view(np.uint32), yo.view(np.uint32))


def test_sampled_factors_and_cluster_intercepts_bit_exact_vs_oracle():
    s = S.add_factor(S.base_spec(2), [0.2, 0.3, 0.5], [0.3, -0.2])
    s = S.add_factor(s, [0.6, 0.4], [0.5])
    s["factor_sampled"] = [True, None]
    s["effect_sizes"][1] = 0.25
    s["cluster"] = {"sizing": {"FixedClusters": {"n_clusters": 7}}, "tau_squared": 0.3, "slopes": [],
                    "extra_groupings": []}
    for sim in (0, 5):
        x, y = M.generate_sim(s, SEED, sim, 211)
        xo, yo, _ = orc.generate_sim(s, SEED, sim, 211)
        assert np.array_equal(x.view(np.uint32), xo.view(np.uint32))
        assert np.array_equal(y.view(np.uint32), yo.view(np.uint32))


def test_prefix_stability_across_max_n():
    s = CS.case_spec(CS.validation_cases()["ols_cf_a"])
    x1, y1 = M.generate_sim(s, 99, 5, 61)
    x2, y2 = M.generate_sim(s, 99, 5, 300)
    assert np.array_equal(x1, x2[:61]) and np.array_equal(y1, y2[:61])


# ------------------------------------------------------------------ solver vs frozen R lm goldens (1e-10, north_star)
@pytest.mark.parametrize("name", sorted(n for n in CS.validation_cases() if n.startswith("ols_")))
def test_fit_on_reference_bytes_matches_r_lm(name):
    case = CS.validation_cases()[name]
    x, y = CS.case_arrays(case)
    g = CS.golden_arrays(case)
    fit = M.fit_data(x.astype(np.float64), y.astype(np.float64))
    assert fit["converged"]
    stat = fit["betas"] / fit["se"]
    for got, want in ((fit["betas"], g["beta"]), (fit["se"], g["se"]), (stat, g["stat"])):
        assert np.max(np.abs(got - want) / np.abs(want)) < 1e-10
    o = orc.fit_ols(x.astype(np.float64), y.astype(np.float64))
    assert np.array_equal(fit["betas"], o["betas"]) and np.array_equal(fit["t_sq"], o["t_sq"])  # same op order


# ------------------------------------------------------------------ run_batch vs oracle: byte tensors identical
def test_headline_shape_small():
    s = S.headline_spec()
    got = M.run_batch(s, [500], 300, M.splitmix64(2137), mode=ALL)
    want = orc.run_batch(s, [500], 300, M.splitmix64(2137), n_threads=orc.host_threads(), capture=True)
    assert_batch_equal(got, want)


@pytest.mark.parametrize("method", ["none", "bonferroni", "holm", "benjamini_hochberg"])
def test_sweep_with_exclusion_and_corrections(method):
    s = S.sweep_spec()
    s["correction_method"] = method
    grid = [20, 24, 40, 100, 333]
    got = M.run_batch(s, grid, 400, SEED, mode=ALL)
    want = orc.run_batch(s, grid, 400, SEED, n_threads=orc.host_threads(), capture=True)
    assert_batch_equal(got, want)
    assert got.factor_excluded[:, 0, 0].all() and not got.factor_excluded[:, 2, 0].any()


@pytest.mark.parametrize("method", ["none", "bonferroni", "holm", "benjamini_hochberg", "tukey_hsd"])
def test_posthoc_blocks_match_oracle(method):
    """Posthoc pairwise contrasts (evaluate_posthoc, posthoc.rs:40-148): two factor blocks (3 and 4 levels), posthoc
    alpha != alpha, every correction method, sparse-factor exclusion live at the small sample sizes."""
    s = CS.posthoc_spec(method)
    grid = [24, 40, 60, 150]
    got = M.run_batch(s, grid, 500, SEED, mode=ALL)
    want = oracle_once(("posthoc", method), lambda: orc.run_batch(s, grid, 500, SEED, n_threads=orc.host_threads(), capture=True))
    assert got.n_posthoc == 3 + 6
    assert_batch_equal(got, want)
    assert got.posthoc_unc[:, -1].any() and got.factor_excluded[:, 0].any()
    assert np.array_equal(got.success_posthoc_uncorrected, want["posthoc_unc"].sum(axis=0))


@pytest.mark.parametrize("name", ["ols_cf_a", "ols_ff_b", "ols_interaction_a", "ols_corr_b", "ols_simple_a"])
def test_validation_specs_run_batch(name):
    case = CS.validation_cases()[name]
    s = CS.case_spec(case)
    s["report_overall"] = True
    s["contrast_pairs"] = [[1, 2]] if case["p"] > 2 else []
    grid = [case["p"] + 2, 50, 200]
    got = M.run_batch(s, grid, 256, SEED + 1, mode=ALL)
    want = orc.run_batch(s, grid, 256, SEED + 1, n_threads=orc.host_threads(), capture=True)
    assert_batch_equal(got, want)


def test_sampled_factor_exclusion_varies_per_sim():
    s = S.add_factor(S.base_spec(2), [0.7, 0.2, 0.1], [0.4, 0.6])
    s["effect_sizes"][1] = 0.3
    s["target_indices"] = [1, 3, 4]
    s["factor_sampled"] = [True]
    s["factor_min_level_count"] = 5
    s["correction_method"] = "holm"
    grid = [30, 45, 90]
    got = M.run_batch(s, grid, 512, SEED, mode=ALL)
    want = orc.run_batch(s, grid, 512, SEED, n_threads=orc.host_threads(), capture=True)
    assert_batch_equal(got, want)
    frac = got.factor_excluded[:, 0, 0].mean()
    assert 0.05 < frac < 0.95  # both the reduced and the full fit were exercised


def test_fit_columns_reduced_test_design():
    s = S.headline_spec()
    s["fit_columns"] = [0, 1, 2, 5]
    s["target_indices"] = [1, 2, 7]
    got = M.run_batch(s, [60, 120], 256, SEED, mode=ALL)
    want = orc.run_batch(s, [60, 120], 256, SEED, n_threads=orc.host_threads(), capture=True)
    assert_batch_equal(got, want)
    assert np.isnan(got.t_sq[:, :, 2]).all()


def test_rank_deficient_design_reports_not_converged():
    s = S.add_factor(S.base_spec(1), [0.5, 0.5], [0.3])
    s = S.add_interaction(s, [2], 0.0)  # a copy of the dummy column → singular X'X
    s["target_indices"] = [1]
    got = M.run_batch(s, [40], 64, SEED, mode=ALL)
    want = orc.run_batch(s, [40], 64, SEED, capture=True)
    assert_batch_equal(got, want)
    # the pivot of the duplicated column is exactly 0 in some sims and rounding-level (>1e-12 relative,
    # batch.rs:40-41) in others; both paths must agree with the oracle (checked above), and the flagged
    # fits must report nothing significant
    assert not got.converged.all()
    assert not got.uncorrected[got.converged[:, 0] == 0].any()


def test_every_template_width():
    for k in range(0, 16):
        s = S.base_spec(k)
        s["target_indices"] = list(range(1, 1 + k))
        s["effect_sizes"] = [0.0] + [0.2] * k
        s["report_overall"] = True
        got = M.run_batch(s, [k + 3, 64], 96, SEED + k, mode=ALL)
        want = orc.run_batch(s, [k + 3, 64], 96, SEED + k, n_threads=orc.host_threads(), capture=True)
        assert_batch_equal(got, want)


@pytest.mark.parametrize("k", [2, 8, 9])
def test_plain_grid_points_around_step_boundaries(k):
    """The plain kernels run a 16-row step without per-row tests when it lies inside the design and ahead of the next grid point
    (kernels_lean.cuh, interior steps) and multiply every column into the linear predictor, zero effects included: grid points on,
    one before and one after every kind of step boundary, zero / negative-zero / non-zero effects and a non-zero intercept."""
    s = S.base_spec(k)
    s["target_indices"] = list(range(1, 1 + k))
    eff = [0.35] + [0.0] * k
    eff[1] = 0.3
    eff[k] = -0.25
    if k >= 8:
        eff[4] = -0.0
        eff[6] = 0.15
    s["effect_sizes"] = eff
    s["report_overall"] = True
    grid = [k + 2, 15, 16, 17, 31, 32, 33, 48, 63, 65, 80, 81, 96, 111, 113, 128, 130]
    grid = sorted(set(g for g in grid if g >= k + 2))
    got = M.run_batch(s, grid, 160, SEED ^ k, mode=ALL)
    want = oracle_once(("step_boundaries", k), lambda: orc.run_batch(s, grid, 160, SEED ^ k, n_threads=orc.host_threads(), capture=True))
    assert_batch_equal(got, want)
    # a single distant grid point: every step but the last is interior
    got = M.run_batch(s, [203], 96, SEED ^ k, mode=ALL)
    want = oracle_once(("step_boundaries_single", k), lambda: orc.run_batch(s, [203], 96, SEED ^ k, n_threads=orc.host_threads(), capture=True))
    assert_batch_equal(got, want)


# ------------------------------------------------------------------ robustness scenarios (BASELINE config 3)
def _scenario_spec(preset: str, alpha: float = 0.05) -> dict:
    s = S.sweep_spec()
    s["scenario"] = S.scenario(preset)
    s["crit_values"]["alpha"] = alpha
    s["heteroskedasticity_driver"] = None
    return s


@pytest.mark.parametrize("preset", ["realistic", "doomer"])
def test_scenario_design_draw_bit_exact_vs_oracle(preset):
    """Per-sim perturbed correlation (+ Jacobi PSD repair), marginal swaps, residual swaps (t / chi-square
    slots), per-study beta jitter, heteroskedastic residuals, sampled factor levels: X and y bit-exact."""
    s = _scenario_spec(preset)
    repaired = 0
    for sim in range(12):
        x, y = M.generate_sim(s, SEED, sim, 150)
        xo, yo, _ = orc.generate_sim(s, SEED, sim, 150)
        assert np.array_equal(x.view(np.uint32), xo.view(np.uint32)), (preset, sim)
        assert np.array_equal(y.view(np.uint32), yo.view(np.uint32)), (preset, sim)
        repaired += 1
    assert repaired == 12


@pytest.mark.parametrize("preset", ["realistic", "doomer"])
def test_scenario_run_batch_matches_oracle(preset):
    s = _scenario_spec(preset)
    grid = [30, 90, 240]
    got = M.run_batch(s, grid, 384, SEED, mode=ALL)
    want = orc.run_batch(s, grid, 384, SEED, n_threads=orc.host_threads(), capture=True)
    assert_batch_equal(got, want)


def test_scenario_single_knobs_match_oracle():
    base = S.headline_spec()
    base["report_overall"] = True
    knobs = [("heterogeneity", 0.4), ("heteroskedasticity_ratio", 4.0), ("correlation_noise_sd", 0.3),
             ("distribution_change_prob", 0.8), ("residual_change_prob", 0.8)]
    for key, val in knobs:
        s = dict(base, scenario=dict(S.scenario("optimistic"), **{key: val, "residual_df": 5.0}))
        got = M.run_batch(s, [60, 200], 256, SEED, mode=ALL)
        want = orc.run_batch(s, [60, 200], 256, SEED, n_threads=orc.host_threads(), capture=True)
        assert_batch_equal(got, want)
    s = dict(base, heteroskedasticity_driver=2,
             scenario=dict(S.scenario("optimistic"), heteroskedasticity_ratio=3.0))
    got = M.run_batch(s, [80], 256, SEED, mode=ALL)
    want = orc.run_batch(s, [80], 256, SEED, n_threads=orc.host_threads(), capture=True)
    assert_batch_equal(got, want)


def test_five_sigma_alpha_joint_power_doomer():
    """Config 3: all-tests-significant joint power at alpha = 5e-7 — the last histogram bucket."""
    s = _scenario_spec("doomer", alpha=5e-7)
    s["effect_sizes"] = [0.0, 0.6, 0.5, 0.5, 0.4, 0.4, 0.4, 0.4, 0.4, 0.8, 1.0]
    n_sims = 4096
    got = M.run_batch(s, [400, 1000], n_sims, SEED, mode=M.MODE_COUNTS | M.MODE_BYTES)
    want = orc.run_batch(s, [400, 1000], n_sims, SEED, n_threads=orc.host_threads())
    assert np.array_equal(got.uncorrected, want["uncorrected"])
    T = got.n_targets
    joint = got.hist_uncorrected[:, T] / n_sims
    assert np.array_equal(got.hist_uncorrected[:, T], (want["uncorrected"].sum(axis=2) == T).sum(axis=0))
    assert joint[1] >= joint[0]


# ------------------------------------------------------------------ GLM IRLS (BASELINE config 4)
def test_logit_power_curve_matches_oracle():
    s = S.logit_spec()
    grid = [50, 100, 200, 500, 1000, 2000]
    got = M.run_batch(s, grid, 512, SEED, mode=ALL)
    want = orc.run_batch(s, grid, 512, SEED, n_threads=orc.host_threads(), capture=True)
    assert_batch_equal(got, want)
    p = got.power()[:, 0]
    assert p[-1] > 0.95 and np.all(np.diff(p) > -0.05)


@pytest.mark.parametrize("family", ["probit", "poisson"])
def test_probit_and_poisson_match_oracle(family):
    s = S.logit_spec()
    if family == "probit":
        s["link"] = "probit"
        s["effect_sizes"][0] = -0.5
    else:
        s["outcome_kind"] = "count"
        s["effect_sizes"] = [0.7, 0.3, 0.2, 0.1, 0.1]
    s["correction_method"] = "holm"
    got = M.run_batch(s, [60, 150, 400], 384, SEED + 3, mode=ALL)
    want = orc.run_batch(s, [60, 150, 400], 384, SEED + 3, n_threads=orc.host_threads(), capture=True)
    assert_batch_equal(got, want)


def test_glm_factor_exclusion_and_separation_fallback_match_oracle():
    """Rare outcome + sparse 4-level factor at small N: sparse-level exclusion (code 1), non-convergence from
    quasi-separation and the fallback refit without the sparsest factor (code 2), reduced-df LRT."""
    s = S.add_factor(S.base_spec(1), [0.55, 0.25, 0.12, 0.08], [0.8, 2.5, -2.5])
    s["effect_sizes"][0] = S.logit_intercept(0.12)
    s["effect_sizes"][1] = 0.6
    s["target_indices"] = [1, 2, 3, 4]
    s["outcome_kind"], s["estimator"] = "binary", "glm"
    s["report_overall"] = True
    s["factor_min_level_count"] = 5
    s["factor_sampled"] = [True]
    s["correction_method"] = "benjamini_hochberg"
    grid = [40, 70, 120, 300]
    got = M.run_batch(s, grid, 768, SEED, mode=ALL)
    want = orc.run_batch(s, grid, 768, SEED, n_threads=orc.host_threads(), capture=True)
    assert_batch_equal(got, want)
    codes = set(np.unique(got.factor_excluded).tolist())
    assert {0, 1}.issubset(codes)          # sparse exclusion exercised
    assert got.converged.mean() < 1.0 or 2 in codes  # fallback / non-convergence exercised


def test_glm_fit_columns_and_scenario_match_oracle():
    s = S.logit_spec()
    s["fit_columns"] = [0, 1, 2]
    s["target_indices"] = [1, 2]
    s["scenario"] = dict(S.scenario("realistic"))
    got = M.run_batch(s, [120, 360], 256, SEED, mode=ALL)
    want = orc.run_batch(s, [120, 360], 256, SEED, n_threads=orc.host_threads(), capture=True)
    assert_batch_equal(got, want)


def test_glm_wave_split_is_invariant():
    s = S.logit_spec()
    a = M.run_batch(s, [100, 300], 700, SEED, mode=ALL)
    b = M.run_batch(s, [100, 300], 700, SEED, mode=ALL, wave_sims=256)
    assert np.array_equal(a.uncorrected, b.uncorrected) and np.array_equal(a.counters, b.counters)
    assert np.array_equal(a.t_sq, b.t_sq, equal_nan=True) and b.gpu_launches == 9  # 3 waves x (generate, fit, transpose the wave's per-fit outputs)


# ------------------------------------------------------------------ random-intercept LMM, REML (BASELINE config 5)
def test_lmm_random_intercept_matches_oracle():
    s = S.lmm_spec(30, 0.2)
    grid = [60, 120, 300, 600]
    got = M.run_batch(s, grid, 512, SEED, mode=ALL)
    want = orc.run_batch(s, grid, 512, SEED, n_threads=orc.host_threads(), capture=True)
    assert_batch_equal(got, want)
    assert got.boundary_hit[:, 0].mean() > 0.0 and got.success_joint[-1] > 0.95 * 512
    p = got.power()[:, 0]
    assert p[-1] > 0.95 and np.all(np.diff(p) > -0.05)


def test_lmm_and_glm_under_scenarios_match_oracle():
    """Robustness scenarios on the mixed-model and GLM arms: the per-study designs come from scen_design_kernel and the
    quad generators read them (correlated predictors + doomer noise → the Jacobi PSD repair runs for most studies)."""
    m = S.lmm_spec(12, 0.2)
    m["correlation"] = [1.0, 0.5, 0.5, 1.0]
    m["scenario"] = dict(S.scenario("doomer"), sampled_factor_proportions=False)
    got = M.run_batch(m, [48, 96, 240], 384, SEED, mode=ALL)
    want = orc.run_batch(m, [48, 96, 240], 384, SEED, n_threads=orc.host_threads(), capture=True)
    assert_batch_equal(got, want)
    assert got.gpu_launches == 4  # designs, generate, fit, transpose (BYTES / CAPTURE on a multi-point grid)
    g = S.logit_spec()
    g["correlation"] = [1, 0.4, 0.2, 0.4, 1, 0.3, 0.2, 0.3, 1]
    g["scenario"] = S.scenario("doomer")
    got = M.run_batch(g, [80, 200, 500], 384, SEED + 1, mode=ALL)
    want = orc.run_batch(g, [80, 200, 500], 384, SEED + 1, n_threads=orc.host_threads(), capture=True)
    assert_batch_equal(got, want)
    assert got.gpu_launches == 4


def test_lmm_fixed_size_unbalanced_prefixes_match_oracle():
    """FixedSize sizing with sample sizes that are not multiples of the cluster size: two cluster-size groups."""
    s = S.lmm_spec(10, 0.3)
    s["cluster"]["sizing"] = {"FixedSize": {"cluster_size": 8}}
    s["correction_method"] = "holm"
    got = M.run_batch(s, [40, 100, 160], 384, SEED + 9, mode=ALL)
    want = orc.run_batch(s, [40, 100, 160], 384, SEED + 9, n_threads=orc.host_threads(), capture=True)
    assert_batch_equal(got, want)
    s2 = S.lmm_spec(7, 0.1)   # FixedClusters with N % n_clusters != 0
    got = M.run_batch(s2, [30, 95], 384, SEED + 9, mode=ALL)
    want = orc.run_batch(s2, [30, 95], 384, SEED + 9, n_threads=orc.host_threads(), capture=True)
    assert_batch_equal(got, want)


def test_lmm_with_factor_exclusion_and_interaction_matches_oracle():
    s = S.add_factor(S.lmm_spec(12, 0.25), [0.6, 0.25, 0.15], [0.4, 0.7])
    s = S.add_interaction(s, [1, 2], 0.15)
    s["target_indices"] = [1, 2, 3, 4, 5]
    s["factor_min_level_count"] = 5
    s["factor_sampled"] = [True]
    got = M.run_batch(s, [24, 48, 120], 384, SEED, mode=ALL)
    want = orc.run_batch(s, [24, 48, 120], 384, SEED, n_threads=orc.host_threads(), capture=True)
    assert_batch_equal(got, want)
    assert 0.02 < got.factor_excluded[:, 0, 0].mean() < 0.98


# ------------------------------------------------------------------ determinism contract (batch.rs:2715-2786)
def test_results_independent_of_wave_size_and_shard_offset(ols_kernel_choice):
    s = S.sweep_spec()
    grid = [40, 80]
    a = M.run_batch(s, grid, 1000, SEED, mode=ALL)
    b = M.run_batch(s, grid, 1000, SEED, mode=ALL, wave_sims=137)
    assert np.array_equal(a.uncorrected, b.uncorrected) and np.array_equal(a.counters, b.counters)
    # the lean and split paths round a wave down to whole 128-study groups and launch two kernels per wave (accumulate, solve)
    # per wave: accumulate + solve on the lean / split paths (one launch on the cross-check kernels), + the transposition of the
    # wave's per-fit outputs (BYTES / CAPTURE on a multi-point grid)
    assert b.gpu_launches == (3 * math.ceil(1000 / 128) if ols_kernel_choice in ("lean", "split") else 2 * math.ceil(1000 / 137))  # optimistic spec
    c1 = M.run_batch(s, grid, 600, SEED, mode=ALL, sim_begin=0)
    c2 = M.run_batch(s, grid, 400, SEED, mode=ALL, sim_begin=600)
    assert np.array_equal(np.concatenate([c1.uncorrected, c2.uncorrected]), a.uncorrected)
    assert np.array_equal(c1.counters + c2.counters, a.counters)


def test_progress_and_cancel():
    s = S.headline_spec()
    seen = []
    M.run_batch(s, [100], 5000, SEED, lambda cur, tot: seen.append((cur, tot)) or True)
    # the kernels publish ~50 progress values per call; the waiting host thread reports the ones it sees (a 0.1 ms batch is
    # over before most of them), always ending with (n, n) — tests/test_gpu_hostpath.py covers the long-running case
    assert seen[-1] == (5000, 5000) and 1 <= len(seen) <= 51
    assert [c for c, _ in seen] == sorted({c for c, _ in seen})
    with pytest.raises(M.Cancelled):
        M.run_batch(s, [100], 5000, SEED, lambda cur, tot: cur < 1000)


def test_errors_match_reference_classes():
    s = S.headline_spec()
    with pytest.raises(M.InvalidSpec):
        M.run_batch(s, [9], 10, SEED)  # df = N - P < 1 (critvals.rs:874-879)
    bad = S.base_spec(2)
    bad["correlation"] = [1.0, 1.2, 1.2, 1.0]
    with pytest.raises(M.CorrelationNotPSD):
        M.run_batch(bad, [50], 10, SEED)
    glmm = S.logit_spec()
    glmm["cluster"] = {"sizing": {"FixedClusters": {"n_clusters": 10}}, "tau_squared": 0.25, "slopes": [],
                       "extra_groupings": []}
    with pytest.raises(M.Unsupported):
        M.run_batch(glmm, [50], 10, SEED)
    r = M.run_batch(s, [50], 0, SEED, mode=ALL)  # n_sims == 0 → empty-shaped result (batch.rs:390-421)
    assert r.uncorrected.shape == (0, 1, 2) and r.counters.sum() == 0


# ------------------------------------------------------------------ uploaded-data arms
def _upload_frame(n=60):
    rng = np.random.default_rng(17)
    x = rng.exponential(size=n)
    x = (x - x.mean()) / x.std()
    return [[float(x[i]), float(x[i] ** 2 - 1.0), (1.0 if rng.random() < 0.35 else 0.0) - 0.35, float(rng.integers(0, 3))] for i in range(n)]


def _upload_spec(kind: str) -> dict:
    """NORTA (Gaussian-copula marginal through the sorted table) and strict bootstrap (whole-row copies) specs:
    correlated block, uploaded continuous + binary columns, a frame factor next to a sampled synthetic factor
    (so the bootstrap pick is the THIRD sequential-stream word of a row), an interaction on copied columns."""
    s = S.base_spec(3)
    s["correlation"] = [1, 0.4, 0.2, 0.4, 1, 0.3, 0.2, 0.3, 1]
    s["effect_sizes"] = [0.1, 0.3, 0.25, 0.4]
    s = S.add_factor(s, [0.5, 0.3, 0.2], [0.3, -0.4])
    s = S.add_factor(s, [0.6, 0.4], [0.2])
    s["factor_sampled"] = [None, True]
    s = S.add_interaction(s, [1, 3], 0.2)
    s["target_indices"] = [1, 2, 3, 4, 7]
    s["correction_method"] = "holm"
    s["report_overall"] = True
    s["factor_min_level_count"] = 5
    frame = _upload_frame()
    if kind.startswith("norta"):
        s = S.with_uploaded_frame(s, frame, {0: 0, 2: (2, 0.35)})
    else:
        s = S.with_uploaded_frame(s, frame, {0: 0, 1: 1, 2: (2, 0.35)}, factors={0: 3}, bootstrap=True)
    if kind.endswith("doomer"):
        s["scenario"] = S.scenario("doomer")
    if kind.endswith("cluster"):
        s["cluster"] = {"sizing": {"FixedClusters": {"n_clusters": 6}}, "tau_squared": 0.2, "slopes": [], "extra_groupings": []}
    return s


_UPLOAD_KINDS = ["norta", "norta_doomer", "strict", "strict_doomer", "strict_cluster"]


@pytest.mark.parametrize("kind", _UPLOAD_KINDS)
def test_uploaded_generated_design_bit_exact_vs_oracle(kind):
    s = _upload_spec(kind)
    for sim in (0, 3, 11):
        x, y = M.generate_sim(s, SEED, sim, 203)
        xo, yo, _ = orc.generate_sim(s, SEED, sim, 203)
        assert np.array_equal(x.view(np.uint32), xo.view(np.uint32)), (kind, sim)
        assert np.array_equal(y.view(np.uint32), yo.view(np.uint32)), (kind, sim)
    if kind.startswith("strict"):  # every generated row is one whole frame row (test_upload_bootstrap.rs:104-131)
        fr = np.asarray(_upload_frame())
        src = {(np.float32(r[0]).item(), np.float32(r[1]).item(), 1.0 if r[2] + 0.35 >= 0.5 else 0.0, int(r[3])) for r in fr}
        for i in range(203):
            lvl = 1 if x[i, 4] == 1.0 else 2 if x[i, 5] == 1.0 else 0
            assert (float(x[i, 1]), float(x[i, 2]), float(x[i, 3]), lvl) in src, i


@pytest.mark.parametrize("kind", _UPLOAD_KINDS)
def test_uploaded_run_batch_matches_oracle(kind):
    s = _upload_spec(kind)
    grid = [24, 48, 150]
    got = M.run_batch(s, grid, 384, SEED, mode=ALL)
    want = oracle_once(("upload", kind), lambda: orc.run_batch(s, grid, 384, SEED, n_threads=orc.host_threads(), capture=True))
    assert_batch_equal(got, want)
    if kind.startswith("strict"):  # the copied factor's level counts differ per study: both fits were exercised at n = 24
        assert 0.0 < got.factor_excluded[:, 0, 0].mean() < 1.0


def test_uploaded_logit_and_lmm_match_oracle():
    g = S.logit_spec()
    g = S.with_uploaded_frame(g, _upload_frame(), {1: 0, 0: (2, 0.35)}, bootstrap=True)
    got = M.run_batch(g, [80, 200], 256, SEED, mode=ALL)
    want = orc.run_batch(g, [80, 200], 256, SEED, n_threads=orc.host_threads(), capture=True)
    assert_batch_equal(got, want)
    m = S.with_uploaded_frame(S.lmm_spec(10, 0.2), _upload_frame(), {0: 0})
    got = M.run_batch(m, [60, 120], 256, SEED, mode=ALL)
    want = orc.run_batch(m, [60, 120], 256, SEED, n_threads=orc.host_threads(), capture=True)
    assert_batch_equal(got, want)


def test_uploaded_errors():
    s = _upload_spec("strict")
    s["bootstrap_frame_map"][1] = 9
    with pytest.raises(M.InvalidSpec):
        M.run_batch(s, [50], 10, SEED)
    s = _upload_spec("norta")
    s["upload_normal"] = s["upload_normal"][:-1]
    with pytest.raises(M.InvalidSpec):
        M.run_batch(s, [50], 10, SEED)


# ------------------------------------------------------------------ contract entry points
def test_contract_entry_points_match_spec_entry_points():
    """The hosts' payload (msgpack `Vec<SimulationContract>`) through mcp_find_power_contracts / mcp_find_sample_size_contracts
    gives exactly what the lowered spec gives through the spec entry points with base_seed = splitmix64(user_seed)
    (find_power.rs:251-288), labelled in report space; scenarios of one call are paired runs (same seed)."""
    import msgpack
    import test_contract_lowering as TC
    terms = [TC.CONST, TC.direct(0), TC.direct(1), TC.dummy(2, 1), TC.dummy(2, 2)]
    base = TC.contract([TC.normal(), TC.normal(), TC.factor(3, [0.5, 0.3, 0.2])], terms, [0.0, 0.4, 0.25, 0.3, 0.5],
                       [TC.marginal(2), TC.marginal(3), TC.marginal(4), {"kind": "joint", "terms": [1, 2, 3]}],
                       design_test=[TC.CONST, TC.direct(1), TC.dummy(2, 1), TC.dummy(2, 2)], correction="holm",
                       correlations={"Matrix": {"continuous_columns": [0, 1], "values": [1.0, 0.3, 0.3, 1.0]}})
    base["test"]["targets"][0] = TC.marginal(1)  # design_test term positions: 1 = x2, 2/3 = the dummies
    base["test"]["targets"][1] = TC.marginal(2)
    base["test"]["targets"][2] = TC.marginal(3)
    base["scenario"]["name"] = "optimistic"
    rough = __import__("copy").deepcopy(base)
    rough["scenario"] = TC.scenario(name="realistic", heterogeneity=0.2, heteroskedasticity_ratio=2.0, correlation_noise_sd=0.15,
                                    distribution_change_prob=0.5, new_distributions=["RightSkewed", "LeftSkewed", "Uniform"],
                                    residual_change_prob=0.5, residual_dists=["high_kurtosis", "right_skewed"], residual_df=8.0,
                                    sampled_factor_proportions=True)
    blob = msgpack.packb([base, rough], use_bin_type=True)
    lowered = M.lower_contracts(blob)
    assert [sc["label"] for sc in lowered] == ["optimistic", "realistic"]
    assert lowered[0]["spec"]["target_indices"] == [2, 3, 4] and lowered[0]["report_target_indices"] == [1, 2, 3]
    assert lowered[0]["spec"]["fit_columns"] == [0, 2, 3, 4] and lowered[0]["spec"]["report_overall"] is True
    got = M.find_power_contracts(blob, 60, 2048, 2137)
    assert [label for label, _ in got["scenarios"]] == ["optimistic", "realistic"]
    for (label, pr), sc in zip(got["scenarios"], lowered):
        want = M.find_power(sc["spec"], [60], 2048, M.splitmix64(2137))[0]
        assert pr["target_indices"] == sc["report_target_indices"]
        for key in ("n", "n_sims", "success_counts_uncorrected", "success_counts_corrected", "convergence_count",
                    "overall_significant_count", "success_count_histogram_uncorrected", "factor_exclusion_counts"):
            assert pr[key] == want[key], (label, key)
    assert got["scenarios"][0][1]["success_counts_uncorrected"] != got["scenarios"][1][1]["success_counts_uncorrected"]
    ss = M.find_sample_size_contracts(blob, 20, 200, 1024, 2137, by=20)
    for (label, res), sc in zip(ss["scenarios"], lowered):
        want = M.find_sample_size(sc["spec"], 20, 200, 1024, M.splitmix64(2137), by=20)
        assert res["first_achieved"] == want["first_achieved"] and res["fitted"] == want["fitted"], label
        assert [p["success_counts_uncorrected"] for p in res["grid_or_trace"]] == [p["success_counts_uncorrected"] for p in want["grid_or_trace"]]
        assert res["grid_or_trace"][0]["target_indices"] == sc["report_target_indices"]
    with pytest.raises(M.InvalidSpec, match="contract 1"):
        bad = __import__("copy").deepcopy(rough)
        bad["test"]["alpha"] = 2.0
        M.find_power_contracts(msgpack.packb([base, bad]), 60, 16)


# ------------------------------------------------------------------ full-size properties (BASELINE.json sizes)
def test_full_size_counts_are_consistent_and_match_analytic_power():
    s = S.headline_spec()
    n, n_sims = 500, 2_000_000
    r = M.run_batch(s, [n], n_sims, M.splitmix64(2137))
    T = 2
    assert r.n_converged[0] == n_sims
    assert r.hist_uncorrected[0].sum() == n_sims and r.hist_corrected[0].sum() == n_sims
    # checksum of checksums: Σ_k k·hist[k] = Σ_t successes[t]
    assert (np.arange(T + 1) * r.hist_uncorrected[0]).sum() == r.success_uncorrected[0].sum()
    power = CS.analytic_ols_power(0.18, n, 9)
    for t in range(T):
        emp = r.power()[0, t]
        assert abs(emp - power) < 4 * math.sqrt(power * (1 - power) / n_sims) + 1e-4, (emp, power)
    # a shard of the same run reproduces the oracle's counts exactly (same sim ids)
    sub = M.run_batch(s, [n], 3000, M.splitmix64(2137), sim_begin=1_000_000, mode=M.MODE_COUNTS | M.MODE_BYTES)
    want = orc.run_batch(s, [n], 3000, M.splitmix64(2137), sim_begin=1_000_000, n_threads=orc.host_threads())
    assert np.array_equal(sub.uncorrected, want["uncorrected"])


def test_power_is_monotone_in_n_full_grid():
    s = S.sweep_spec()
    grid = list(range(20, 1001, 20))
    r = M.run_batch(s, grid, 20000, SEED)
    p = r.power()[:, 0]
    assert p[-1] > 0.99 and np.all(np.diff(p) > -0.02)
    assert np.all(r.hist_uncorrected.sum(axis=1) == 20000)


# ------------------------------------------------------------------ end-to-end vs the SHIPPED ENGINE's own outputs
@pytest.mark.parametrize("name", ["cross_ols", "cross_partial", "cross_lme", "cross_glm"])
def test_find_sample_size_reproduces_the_engines_frozen_required_n(name):
    """The reference engine's own find_sample_size results (validation/data/cross_*.golden.rds: 20..300 by 2, 100000 sims per
    cell, seed 2137, public API) reproduced through mcp_find_sample_size: OLS (lean quad + solve kernels), logistic GLM
    (IRLS kernel), random-intercept LMM (REML kernel, cluster atom 10) and the not-reached case."""
    g = CS.crossing_goldens()[name]
    r = M.find_sample_size(CS.crossing_spec(g["case"]), g["from_size"], g["to_size"], g["n_sims"], M.splitmix64(g["seed"]),
                           by=g["by"])
    assert r["cluster_atom"] == g["case"].get("cluster", {}).get("n_clusters", 1)
    for fit, want in zip(r["fitted"], g["dense"].values()):
        assert fit["status"] == want["status"], (name, fit)
        if want["status"] == "fitted":
            assert fit["n_achievable"] == want["n_achievable"], (name, fit)
    pr = r["grid_or_trace"][-1]
    assert pr["n"] == 300 and pr["n_sims"] == g["n_sims"] and sum(pr["success_count_histogram_corrected"]) == g["n_sims"]
    lo, hi = M.wilson_ci(pr["success_counts_corrected"][0], pr["n_sims"])
    assert pr["ci_corrected"][0] == {"lo": lo, "hi": hi}
