"""`fit_provided_data` (crates/engine-core/src/introspect.rs:174-562) — north_star's first correctness check: on identical design
matrices and responses exported from the reference, betas / SEs / statistics of the shipped fit code must match.

The datasets are the reference's own `validation/data/<case>.rds` sim-0 exports with their frozen R goldens (`lm`, `glm`,
`lmer`): 15 OLS, 8 logit, 4 probit, 4 Poisson, 8 random-intercept LME (tests/golden/validation_cases.json.gz).  Gates are the
reference's (`validation/tolerances.R:24-37`): 1e-11 relative for the closed-form OLS fit (tested here at north_star's 1e-10),
1e-4 relative for the iterative fits (IRLS / REML), 1e-9 / 1e-8 absolute for the critical values.

The CPU half checks the oracle's REML restatement against the `lmer` goldens; the GPU half (marked `gpu`) runs the CUDA
IRLS / REML / OLS kernels through `mcp_fit_provided_data` and additionally requires bit-equality with the oracle."""
import numpy as np
import pytest

import _cases as CS
import mcpower_b200 as M
from oracle import orc

GLM = sorted(n for n in CS.validation_cases() if n.startswith(("glm_", "probit_", "poisson_")))
LME = sorted(n for n in CS.validation_cases() if n.startswith("lme_"))
OLS = sorted(n for n in CS.validation_cases() if n.startswith("ols_"))


def _rel(a, b):
    return np.max(np.abs(np.asarray(a) - np.asarray(b)) / np.abs(np.asarray(b)))


@pytest.mark.parametrize("name", LME)
def test_oracle_reml_matches_lmer(name):
    case = CS.validation_cases()[name]
    x, y = CS.case_arrays(case)
    g = CS.golden_arrays(case)
    ids = np.asarray(case["cluster_ids"], dtype=np.int32)
    assert np.array_equal(ids, np.arange(case["n"]) % (int(ids.max()) + 1))  # generation.rs cluster_of_row, FixedClusters
    fit = orc.fit_lmm(x.astype(np.float64), y.astype(np.float64), ids)
    assert fit["converged"]
    assert _rel(fit["betas"], g["beta"]) < 1e-4
    assert _rel(fit["se"], g["se"]) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("name", GLM)
def test_gpu_irls_on_reference_bytes_matches_r_glm_and_the_oracle(name):
    case = CS.validation_cases()[name]
    spec = CS.case_spec(case)
    x, y = CS.case_arrays(case)
    g = CS.golden_arrays(case)
    got = M.fit_provided_data(spec, x.astype(np.float64), y.astype(np.float64))
    assert got["converged"] and got["iterations"] > 1
    assert got["target_cols"] == list(range(1, case["p"]))
    assert _rel(got["betas"], g["beta"]) < 1e-4                       # tolerances.R: estimate_rel_iter
    assert _rel(got["se"], g["se"][1:]) < 1e-4
    assert _rel(got["statistic"], np.abs(g["stat"][1:])) < 1e-4
    assert abs(np.sqrt(got["crit_sq_uncorrected"]) - g["crit"][0]) < 1e-8  # crit_abs_iter (z quantile)
    assert got["df_resid"] == case["n"] - case["p"]
    fam = {"glm": "logit", "probit": "probit", "poisson": "poisson"}[name.split("_")[0]]
    want = orc.fit_glm(x.astype(np.float64), y.astype(np.float64), fam, beta_start=spec["effect_sizes"])
    assert np.array_equal(got["betas"], want["betas"]) and np.array_equal(got["se"], want["se"][1:])
    assert got["iterations"] == want["iters"] and got["deviance"] == want["deviance"] and got["deviance_null"] == want["deviance_null"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", LME)
def test_gpu_reml_on_reference_bytes_matches_lmer_and_the_oracle(name):
    case = CS.validation_cases()[name]
    spec = CS.case_spec(case)
    x, y = CS.case_arrays(case)
    g = CS.golden_arrays(case)
    ids = np.asarray(case["cluster_ids"], dtype=np.uint32)
    got = M.fit_provided_data(spec, x.astype(np.float64), y.astype(np.float64), cluster_ids=ids)
    assert got["converged"] and got["boundary_hit"] == 0
    assert _rel(got["betas"], g["beta"]) < 1e-4
    assert _rel(got["se"], g["se"][1:]) < 1e-4
    assert _rel(got["statistic"], np.abs(g["stat"][1:])) < 1e-4
    assert abs(np.sqrt(got["crit_sq_uncorrected"]) - g["crit"][0]) < 1e-8
    want = orc.fit_lmm(x.astype(np.float64), y.astype(np.float64), ids.astype(np.int32))
    assert np.array_equal(got["betas"], want["betas"]) and np.array_equal(got["se"], want["se"][1:])
    assert got["sigma_sq_hat"] == want["sigma_sq"] and got["tau_sq_hat"] == want["theta"] ** 2 * want["sigma_sq"]
    # ICC of the DGP was 0.2: the variance component is in the right place (a loose sanity band, not a gate)
    assert 0.02 < got["tau_sq_hat"] / (got["tau_sq_hat"] + got["sigma_sq_hat"]) < 0.6
    # ids = NULL means "the spec's sizing"; ids that contradict it are refused
    again = M.fit_provided_data(spec, x.astype(np.float64), y.astype(np.float64))
    assert np.array_equal(again["betas"], got["betas"])
    with pytest.raises(M.Unsupported, match="cluster ids"):
        M.fit_provided_data(spec, x.astype(np.float64), y.astype(np.float64), cluster_ids=ids[::-1].copy())


@pytest.mark.gpu
@pytest.mark.parametrize("name", OLS)
def test_gpu_ols_on_reference_bytes_with_spec_matches_r_lm(name):
    case = CS.validation_cases()[name]
    spec = CS.case_spec(case)
    x, y = CS.case_arrays(case)
    g = CS.golden_arrays(case)
    got = M.fit_provided_data(spec, x.astype(np.float64), y.astype(np.float64))
    assert got["converged"]
    assert _rel(got["betas"], g["beta"]) < 1e-10                       # north_star: 1e-10 relative (reference gate 1e-11)
    assert _rel(got["se"], g["se"][1:]) < 1e-10
    assert _rel(got["statistic"], np.abs(g["stat"][1:])) < 1e-10
    assert abs(np.sqrt(got["crit_sq_uncorrected"]) - g["crit"][0]) < 1e-9  # crit_abs (t quantile vs R qt)
    assert got["df_resid"] == case["n"] - case["p"]
    want = orc.fit_ols(x.astype(np.float64), y.astype(np.float64))
    assert np.array_equal(got["betas"], want["betas"])


def test_fit_provided_data_shape_errors_without_a_gpu():
    case = CS.validation_cases()["glm_simple_a"]
    spec = CS.case_spec(case)
    x, y = CS.case_arrays(case)
    with pytest.raises(M.InvalidSpec, match="columns"):
        M.fit_provided_data(spec, x[:, :-1].astype(np.float64), y.astype(np.float64))
    with pytest.raises(M.InvalidSpec, match="df"):
        M.fit_provided_data(spec, x[:2].astype(np.float64), y[:2].astype(np.float64))
