"""Designs wider than the 16 register-resident template widths (the reference has no cap: bin/throughput.rs's
`ols_wide` is k = 15, and one six-level factor next to eleven predictors is already 17 columns).  Columns 17..32 run on the split
path's kernels at a padded width (csrc/plan.h, mcp_template_p): zero padding columns, fits through the runtime-p solvers on the
kept (real) columns.  Every output must equal the oracle's bit for bit, like at the exact widths."""
import numpy as np
import pytest

import mcpower_b200 as M
from mcpower_b200 import spec as S
from oracle import orc

SEED = 0x1DEA5EED0BADF00D
ALL = M.MODE_COUNTS | M.MODE_BYTES | M.MODE_CAPTURE


def wide_spec(k, factor_levels=(), n_inter=0, rho=0.2):
    s = S.base_spec(k)
    s["correlation"] = [1.0 if i == j else rho ** abs(i - j) for j in range(k) for i in range(k)]
    s["effect_sizes"] = [0.1] + [0.25 / (1 + j % 4) * (-1) ** j for j in range(k)]
    for nl in factor_levels:
        props = [1.0 / nl + (0.03 if g == 0 else -0.03 / (nl - 1)) for g in range(nl)]
        s = S.add_factor(s, props, [0.2 + 0.05 * g for g in range(nl - 1)])
    for j in range(n_inter):
        s = S.add_interaction(s, [1 + j, 2 + j], 0.1)
    P = S.n_predictors(s)
    s["target_indices"] = list(range(1, min(P, 31)))
    s["correction_method"] = "holm"
    s["report_overall"] = True
    s["factor_min_level_count"] = 5
    return s


def test_specs_up_to_32_columns_are_accepted_and_33_is_refused_without_a_gpu():
    s = wide_spec(24, (5, 5))  # 1 + 24 + 8 = 33
    with pytest.raises(M.Unsupported, match="n_predictors"):
        M.run_batch(s, [100], 4, 1)
    s = wide_spec(25)
    with pytest.raises(M.Unsupported, match="n_non_factor"):
        M.run_batch(s, [100], 4, 1)


def _check(spec, grid, n, seed=SEED):
    got = M.run_batch(spec, grid, n, seed, mode=ALL)
    want = orc.run_batch(spec, grid, n, seed, n_threads=orc.host_threads(), capture=True)
    for key in ("uncorrected", "corrected", "converged", "overall", "factor_excluded", "joint", "boundary_hit"):
        if want.get(key) is not None:
            assert np.array_equal(getattr(got, key), want[key]), key
    assert np.array_equal(got.t_sq, want["t_sq"], equal_nan=True)
    agg = M.counters_from_bytes(got.uncorrected, got.corrected, got.converged, got.overall, got.factor_excluded, got.joint,
                                got.boundary_hit)
    assert np.array_equal(agg, got.counters)
    return got


@pytest.mark.gpu
@pytest.mark.parametrize("P", [17, 20, 24, 25, 32])
def test_gpu_wide_ols_matches_the_oracle(P):
    spec = {17: wide_spec(11, (6,)), 20: wide_spec(10, (5, 6)), 24: wide_spec(10, (5, 6), 4), 25: wide_spec(24),
            32: wide_spec(20, (4, 5), 4)}[P]
    assert S.n_predictors(spec) == P
    xg, yg = M.generate_sim(spec, SEED, 2, 61)
    xo, yo, _ = orc.generate_sim(spec, SEED, 2, 61)
    assert xg.shape == (61, P)
    assert np.array_equal(xg.view(np.uint32), xo.view(np.uint32)) and np.array_equal(yg.view(np.uint32), yo.view(np.uint32))
    got = _check(spec, [P + 8, 90, 200], 192)
    assert got.n_converged[-1] == 192


@pytest.mark.gpu
def test_gpu_wide_ols_under_scenarios_and_sampled_factors():
    spec = wide_spec(9, (5, 6))  # P = 19
    spec["scenario"] = S.scenario("doomer")
    got = _check(spec, [60, 150], 192)
    assert 0 < got.factor_excluded[:, 0].sum() < got.factor_excluded[:, 0].size  # per-study level counts: some studies exclude


@pytest.mark.gpu
def test_gpu_wide_glm_and_mixed_model_match_the_oracle():
    g = wide_spec(12, (6,))  # P = 18
    g["effect_sizes"] = [S.logit_intercept(0.4)] + [v * 0.6 for v in g["effect_sizes"][1:]]
    g["outcome_kind"], g["estimator"], g["correction_method"] = "binary", "glm", "bonferroni"
    _check(g, [150, 400], 128)
    m = wide_spec(12, (5,), 3)  # P = 20
    m["estimator"], m["report_overall"] = "mle", False
    m["cluster"] = {"sizing": {"FixedClusters": {"n_clusters": 12}}, "tau_squared": 0.25, "slopes": [], "extra_groupings": []}
    got = _check(m, [120, 240], 128)
    assert got.n_converged.min() > 100
