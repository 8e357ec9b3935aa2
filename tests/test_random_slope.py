"""Random slope on the primary grouping, q = 2 (reference: the slope draws `draw_cluster_effects_with_slopes`,
data_gen.rs:1075-1145; `y += b0[c] + b1[c]·x_s`, data_gen.rs:955-1010; the general-q REML fit, lmm/general.rs; the reduced fit
strips the slope, batch.rs:1357-1375).  The slice restated here: ONE or TWO slopes (q <= 3), no extra groupings.

* CPU: the oracle's generator reproduces the reference's saved `lmm_slope_*` sim-0 datasets bit for bit
  (tests/test_oracle_golden.py walks every fixture); here its 3-parameter profiled REML is held against lme4::lmer on those bytes
  at the reference's own gates (validation/tolerances.R: 1e-4 fixed effects and z, 1e-3 variance components, 2e-3 correlation).
* GPU: generated designs, every run_batch output byte and the statistic words equal the oracle's; fit_provided_data returns the
  oracle's doubles.
"""
import copy

import numpy as np
import pytest

import _cases as CS
import mcpower_b200 as M
from mcpower_b200 import spec as S
from oracle import orc

SEED = 0x5EED5107E5EED510
ALL = M.MODE_COUNTS | M.MODE_BYTES | M.MODE_CAPTURE
SLOPE = sorted(n for n in CS.validation_cases() if n in ("lmm_slope_a", "lmm_slope_b", "lmm_multislope"))


def _rel(a, b):
    return np.max(np.abs(np.asarray(a) - np.asarray(b)) / np.abs(np.asarray(b)))


def slope_spec(n_clusters=24, icc=0.2, var=0.1, corr=0.3, col=0, preset="optimistic", k=2, second=None):
    """y ~ x1 + ... + xk + (1 + x_{col+1} | cluster); second = (column, variance, corr with the intercept, corr with the first
    slope) adds a second slope."""
    s = S.lmm_spec(n_clusters, icc, preset=preset) if k == 2 else None
    if s is None:
        s = S.base_spec(k)
        s["scenario"] = S.scenario(preset, mixed=True)
        s["effect_sizes"] = [0.0] + [0.3, 0.2, 0.15, 0.1, 0.25][:k]
        s["target_indices"] = list(range(1, k + 1))
        s["estimator"] = "mle"
        s["cluster"] = {"sizing": {"FixedClusters": {"n_clusters": n_clusters}}, "tau_squared": icc / (1.0 - icc),
                        "slopes": [], "extra_groupings": []}
        s["max_failed_fraction"] = 0.25
    s["cluster"]["slopes"] = [{"column": col, "variance": var, "corr_with_intercept": corr, "corr_with": []}]
    s["cluster_slope_design_cols"] = [1 + col]
    if second is not None:
        c2, v2, r02, r12 = second
        s["cluster"]["slopes"].append({"column": c2, "variance": v2, "corr_with_intercept": r02, "corr_with": [r12]})
        s["cluster_slope_design_cols"].append(1 + c2)
    return s


# --------------------------------------------------------------------------------------------------------------- CPU
def test_fixture_set_holds_the_primary_slope_cases():
    assert SLOPE == ["lmm_multislope", "lmm_slope_a", "lmm_slope_b"]


@pytest.mark.parametrize("name", SLOPE)
def test_oracle_slope_reml_matches_lmer(name):
    case = CS.validation_cases()[name]
    spec = CS.case_spec(case)
    x, y = CS.case_arrays(case)
    g = CS.golden_arrays(case)
    ids = np.asarray(case["cluster_ids"], dtype=np.int32)
    ids = ids - ids.min()
    cols = spec["cluster_slope_design_cols"]
    fit = orc.fit_lmm_slopes(x.astype(np.float64), y.astype(np.float64), ids, cols)
    assert fit["converged"] and fit["boundary"] == 0
    assert _rel(fit["betas"], g["beta"]) < 1e-4                  # tolerances.R: estimate_rel_iter
    assert _rel(fit["betas"] / fit["se"], g["z"]) < 1e-4
    assert _rel(fit["variances"], g["vars"]) < 1e-3              # variance components (intercept, slopes)
    assert _rel(fit["sigma_sq"], g["sigma2"][0]) < 1e-3
    assert np.max(np.abs(fit["corr"] - g["corr"])) < 2e-3        # (1,0) [(2,0) (2,1)]
    # the simplex search has a hard budget (300 evaluations per parameter); a fit that ran out is flagged unconverged
    assert fit["n_eval"] < 300 * (len(cols) + 1) * (len(cols) + 2) // 2


def test_oracle_slope_draws_have_the_specified_covariance():
    """b0 ~ (0, tau²), b1 ~ (0, var), corr(b0, b1) = rho: recovered from y − X beta − noise-free part over many sims is awkward, so
    use the estimator itself: mean REML components over sims sit near the DGP's."""
    s = slope_spec(40, 0.2, var=0.15, corr=0.4)
    got = {"vi": [], "vs": [], "c": []}
    for sim in range(60):
        x, y, ids = orc.generate_sim(s, SEED, sim, 1200)
        f = orc.fit_lmm_slope(x.astype(np.float64), y.astype(np.float64), ids, 1)
        if f["converged"] and f["boundary"] == 0:
            got["vi"].append(f["var_intercept"]); got["vs"].append(f["var_slope"]); got["c"].append(f["corr"])
    assert len(got["vi"]) > 45
    assert abs(np.mean(got["vi"]) - 0.25) < 0.04
    assert abs(np.mean(got["vs"]) - 0.15) < 0.03
    assert abs(np.mean(got["c"]) - 0.4) < 0.12


def test_oracle_slope_validation_and_the_untouched_intercept_model():
    s = slope_spec()
    bad = copy.deepcopy(s)
    bad["cluster"]["tau_squared"] = 0.0  # the slope draw conditions on z0 = u0 / tau (data_gen.rs:1075-1145)
    with pytest.raises(ValueError):
        orc.run_batch(bad, [240], 4, SEED)
    bad = copy.deepcopy(s)
    bad["cluster_slope_design_cols"] = [7]
    with pytest.raises(ValueError):
        orc.run_batch(bad, [240], 4, SEED)
    # the slope arm leaves the intercept-only path untouched: dropping the slope gives back lmm_spec's bytes
    plain = S.lmm_spec(24, 0.2)
    x0, y0, _ = orc.generate_sim(plain, SEED, 2, 96)
    x1, y1, _ = orc.generate_sim(s, SEED, 2, 96)
    assert np.array_equal(x0, x1)           # the design does not depend on the random effects
    assert not np.array_equal(y0, y1)       # the outcome does


def test_oracle_run_batch_with_a_slope_has_calibrated_size_and_sane_power():
    s = slope_spec(30, 0.2, var=0.1, corr=0.3)
    s["effect_sizes"] = [0.0, 0.0, 0.25]   # x1 (the slope column) has no fixed effect: its rejection rate is the test's size
    r = orc.run_batch(s, [600], 1500, SEED, n_threads=orc.host_threads())
    conv = r["converged"][:, 0].astype(bool)
    assert conv.mean() > 0.95
    size = r["uncorrected"][conv, 0, 0].mean()
    assert 0.02 < size < 0.095  # nominal 0.05, Wald z with 30 clusters runs slightly liberal
    assert r["uncorrected"][conv, 0, 1].mean() > 0.9


def test_oracle_two_slopes_batch_converges_and_keeps_size():
    s = slope_spec(30, 0.2, var=0.1, corr=0.3, second=(1, 0.08, 0.1, 0.2))
    s["effect_sizes"] = [0.0, 0.0, 0.25]
    r = orc.run_batch(s, [600], 600, SEED, n_threads=orc.host_threads())
    conv = r["converged"][:, 0].astype(bool)
    assert conv.mean() > 0.95
    assert 0.02 < r["uncorrected"][conv, 0, 0].mean() < 0.1
    assert r["uncorrected"][conv, 0, 1].mean() > 0.8


def test_more_than_two_slopes_are_refused_without_a_gpu():
    s = slope_spec(k=3, second=(1, 0.1, 0.0, 0.0))
    s["cluster"]["slopes"].append({"column": 2, "variance": 0.1, "corr_with_intercept": 0.0, "corr_with": [0.0, 0.0]})
    s["cluster_slope_design_cols"] = [1, 2, 3]
    with pytest.raises(M.Unsupported, match="random slope"):
        M.run_batch(s, [240], 4, SEED)


# --------------------------------------------------------------------------------------------------------------- GPU
def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.gpu
@pytest.mark.parametrize("sizing", [{"FixedClusters": {"n_clusters": 12}}, {"FixedSize": {"cluster_size": 7}}])
@pytest.mark.parametrize("preset", ["optimistic", "doomer"])
def test_gpu_generated_outcome_with_a_slope_is_bit_exact(sizing, preset):
    s = slope_spec(12, 0.25, var=0.2, corr=-0.4, col=1, preset=preset, k=3)
    s["cluster"]["sizing"] = sizing
    for sim in (0, 3, 987654):
        xg, yg = M.generate_sim(s, SEED, sim, 101)
        xo, yo, _ = orc.generate_sim(s, SEED, sim, 101)
        assert np.array_equal(_bits(xg), _bits(xo)) and np.array_equal(_bits(yg), _bits(yo)), sim


@pytest.mark.gpu
@pytest.mark.parametrize("name", SLOPE)
def test_gpu_generator_reproduces_the_reference_slope_datasets(name):
    case = CS.validation_cases()[name]
    spec = CS.case_spec(case)
    xg, yg = M.generate_sim(spec, orc.splitmix64(CS.case_seed(case)), 0, case["n"])
    xr, yr = CS.case_arrays(case)
    assert np.array_equal(_bits(xg), _bits(xr)) and np.array_equal(_bits(yg), _bits(yr))


@pytest.mark.gpu
@pytest.mark.parametrize("layout", ["clusters_24", "size_10", "p5_col3", "doomer", "two_slopes", "two_slopes_p5_size"])
def test_gpu_run_batch_with_a_slope_matches_the_oracle(layout):
    s = slope_spec(24, 0.2, var=0.1, corr=0.3)
    grid = [240, 480]
    if layout == "size_10":
        s["cluster"]["sizing"] = {"FixedSize": {"cluster_size": 10}}
        grid = [150, 333]
    elif layout == "p5_col3":
        s = slope_spec(20, 0.3, var=0.25, corr=-0.5, col=2, k=4)
        grid = [200, 400, 600]
    elif layout == "doomer":
        s = slope_spec(30, 0.2, var=0.1, corr=0.3, preset="doomer")
    elif layout == "two_slopes":
        s = slope_spec(30, 0.2, var=0.1, corr=0.3, second=(1, 0.08, 0.1, 0.2))
        grid = [300, 600]
    elif layout == "two_slopes_p5_size":
        s = slope_spec(20, 0.25, var=0.2, corr=-0.3, col=3, k=4, preset="realistic", second=(0, 0.15, 0.2, -0.4))
        s["cluster"]["sizing"] = {"FixedSize": {"cluster_size": 12}}
        grid = [240, 350]
    n = 256
    got = M.run_batch(s, grid, n, SEED, mode=ALL)
    want = orc.run_batch(s, grid, n, SEED, n_threads=orc.host_threads(), capture=True)
    for key in ("uncorrected", "corrected", "converged", "joint", "boundary_hit"):
        assert np.array_equal(getattr(got, key), want[key]), key
    assert np.array_equal(got.t_sq.view(np.uint64), want["t_sq"].view(np.uint64))
    agg = M.counters_from_bytes(got.uncorrected, got.corrected, got.converged, got.overall, got.factor_excluded, got.joint,
                                got.boundary_hit)
    assert np.array_equal(agg, got.counters)
    assert got.n_converged.min() > 0.7 * n
    # waves, an offset first sim and (when the box has them) several devices leave every byte alone
    part = M.run_batch(s, grid, n - 100, SEED, mode=ALL, sim_begin=100, wave_sims=64)
    assert np.array_equal(part.uncorrected, got.uncorrected[100:]) and np.array_equal(part.t_sq.view(np.uint64), got.t_sq[100:].view(np.uint64))
    if M.device_count() > 1:
        many = M.run_batch(s, grid, n, SEED, mode=ALL, n_devices=-1)
        assert np.array_equal(many.counters, got.counters) and np.array_equal(many.t_sq.view(np.uint64), got.t_sq.view(np.uint64))


@pytest.mark.gpu
@pytest.mark.parametrize("name", SLOPE)
def test_gpu_slope_reml_on_reference_bytes_matches_lmer_and_the_oracle(name):
    case = CS.validation_cases()[name]
    spec = CS.case_spec(case)
    x, y = CS.case_arrays(case)
    g = CS.golden_arrays(case)
    ids = np.asarray(case["cluster_ids"], dtype=np.int64)
    ids = (ids - ids.min()).astype(np.uint32)
    got = M.fit_provided_data(spec, x.astype(np.float64), y.astype(np.float64), cluster_ids=ids)
    assert got["converged"] and got["boundary_hit"] == 0
    assert _rel(got["betas"], g["beta"]) < 1e-4
    assert _rel(got["statistic"], np.abs(g["z"][1:])) < 1e-4
    ns = len(spec["cluster_slope_design_cols"])
    nc = 1 if ns == 1 else 3
    assert _rel([got["tau_sq_hat"]] + got["slope_var_hat"][:ns], g["vars"]) < 1e-3
    assert _rel(got["sigma_sq_hat"], g["sigma2"][0]) < 1e-3
    assert np.max(np.abs(np.array(got["slope_corr_hat"][:nc]) - g["corr"])) < 2e-3
    want = orc.fit_lmm_slopes(x.astype(np.float64), y.astype(np.float64), ids.astype(np.int32), spec["cluster_slope_design_cols"])
    assert np.array_equal(got["betas"], want["betas"]) and np.array_equal(got["se"], want["se"][1:])
    assert got["sigma_sq_hat"] == want["sigma_sq"] and got["tau_sq_hat"] == want["variances"][0]
    assert got["slope_var_hat"][:ns] == list(want["variances"][1:]) and got["slope_corr_hat"][:nc] == list(want["corr"])


def _slope_contract(n_slopes):
    import test_contract_lowering as TC
    slopes = [{"column": 0, "variance": 0.1, "corr_with_intercept": 0.3, "corr_with": []}]
    if n_slopes == 2:
        slopes.append({"column": 1, "variance": 0.08, "corr_with_intercept": 0.1, "corr_with": [0.2]})
    return TC.contract([TC.normal(), TC.normal()], [TC.CONST, TC.direct(0), TC.direct(1)], [0.0, 0.3, 0.2],
                       [TC.marginal(1), TC.marginal(2)], estimator="mle",
                       cluster={"sizing": {"FixedClusters": {"n_clusters": 30}}, "tau_squared": 0.25, "slopes": slopes, "extra_groupings": []})


@pytest.mark.parametrize("n_slopes", [1, 2])
def test_contract_with_primary_slopes_lowers_to_design_columns(n_slopes):  # contract_adapter.rs:132-140, 2267-2340
    import msgpack
    low = M.lower_contracts(msgpack.packb([_slope_contract(n_slopes)], use_bin_type=True))[0]
    assert low["spec"]["cluster_slope_design_cols"] == [1, 2][:n_slopes]
    assert len(low["spec"]["cluster"]["slopes"]) == n_slopes
    r = orc.run_batch(low["spec"], [300], 16, 5)
    assert r["converged"].shape == (16, 1)


@pytest.mark.gpu
@pytest.mark.parametrize("n_slopes", [1, 2])
def test_gpu_contract_with_primary_slopes_runs_end_to_end(n_slopes):
    """The hosts' payload (msgpack contract) for `y ~ x1 + x2 + (1 + x1 [+ x2] | g)` through the contract entry point: it used to
    come back UNSUPPORTED; the counts are the oracle's."""
    import msgpack
    blob = msgpack.packb([_slope_contract(n_slopes)], use_bin_type=True)
    low = M.lower_contracts(blob)[0]
    got = M.find_power_contracts(blob, 600, 512, 2137)
    pr = got["scenarios"][0][1]
    want = orc.run_batch(low["spec"], [600], 512, M.splitmix64(2137), n_threads=orc.host_threads())
    assert pr["success_counts_uncorrected"] == [int(v) for v in want["uncorrected"].sum(axis=0)[0]]
    assert pr["convergence_count"] == int(want["converged"].sum())
    # EstimatorExtras::Mle (result.rs:341-398): per-component pin counts over the converged fits, one entry per variance component
    ex = pr["estimator_extras"]
    conv = want["converged"][:, 0] != 0
    code, pins = want["boundary_hit"][:, 0] & 3, want["boundary_hit"][:, 0] >> 2
    assert ex["boundary_component_counts"] == [int((((pins >> k) & 1) != 0)[conv].sum()) for k in range(1 + n_slopes)]
    assert ex["singular_count"] == int((code == 1)[conv].sum()) and ex["boundary_hits"] == int((code != 0).sum())
    assert len(ex["boundary_rate_per_component"]) == 1 + n_slopes
