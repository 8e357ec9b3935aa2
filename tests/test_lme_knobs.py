"""Mixed-model scenario knobs and cluster-level predictors.

Reference: random-effect draws through `draw_residual` with `scenario.lme.random_effect_dist / random_effect_df`
(crates/engine-core/src/data_gen.rs:146-161, 1289-1320), per-study ICC jitter from the scenario stream after the residual pick,
clamped at 0 (data_gen.rs:1065-1075), `between_var_indices` broadcast of the representative row (data_gen.rs:605-640; the
properties below restate the reference's own tests DGEN-CL1..CL4, data_gen.rs:2593-2830).

The CPU half pins the oracle: the cluster effects it adds are reconstructed from two streams that are themselves pinned to the
reference's golden vectors (`draw_residual`, golden data_gen.rs:2446-2505; the scenario stream, golden_rng.rs:177-206).  The GPU
half (marked `gpu`) requires the CUDA path to equal the oracle bit for bit."""
import numpy as np
import pytest

import mcpower_b200 as M
from mcpower_b200 import spec as S
from oracle import orc

SEED = 0x5EEDFACE12345678
ALL = M.MODE_COUNTS | M.MODE_BYTES | M.MODE_CAPTURE
SCEN_DOMAIN = 0x5C5C5C5C5C5C5C5C  # scenarios.rs:19


def lme_only(spec, dist="high_kurtosis", df=5.0, icc_sd=0.0):
    """A scenario whose only live knobs are the mixed-model ones."""
    s = dict(spec)
    s["scenario"] = dict(S.scenario("optimistic"), name="lme_probe",
                         lme={"random_effect_dist": dist, "random_effect_df": df, "icc_noise_sd": icc_sd})
    return s


def cluster_effects(spec, seed, sim, n):
    """u_c recovered from the generated data: the same study with and without tau² differs by f32 rounding of y only."""
    with_u = orc.generate_sim(spec, seed, sim, n)[1].astype(np.float64)
    flat = dict(spec, cluster=dict(spec["cluster"], tau_squared=0.0))
    if flat["scenario"]["lme"]:
        flat["scenario"] = dict(flat["scenario"], lme=dict(flat["scenario"]["lme"], icc_noise_sd=0.0))
    without = orc.generate_sim(flat, seed, sim, n)[1].astype(np.float64)
    return with_u - without


@pytest.mark.parametrize("dist,df", [("normal", 0.0), ("high_kurtosis", 5.0), ("high_kurtosis", 10.0), ("right_skewed", 6.0),
                                      ("left_skewed", 3.0), ("uniform", 0.0)])
def test_oracle_random_effects_are_draw_residual_times_tau(dist, df):
    nc, n, tau2 = 12, 48, 0.7
    spec = lme_only(S.lmm_spec(nc, 0.2), dist, df)
    spec["cluster"]["tau_squared"] = tau2
    spec["effect_sizes"] = [0.0, 0.0, 0.0]
    for sim in (0, 7):
        diff = cluster_effects(spec, SEED, sim, n)
        draws = orc.draw_residual(SEED, sim, dist, df, nc)  # the cluster draws open the study's sequential stream
        u = (draws * np.sqrt(tau2)).astype(np.float32).astype(np.float64)
        assert np.allclose(diff, u[np.arange(n) % nc], atol=2e-6), (dist, sim)


def test_oracle_normal_lme_block_with_zero_jitter_is_byte_identical_to_the_optimistic_draw():  # data_gen.rs:1063-1064
    base = S.lmm_spec(10, 0.3)
    x0, y0, _ = orc.generate_sim(base, SEED, 3, 50)
    x1, y1, _ = orc.generate_sim(lme_only(base, "normal", 0.0, 0.0), SEED, 3, 50)
    assert np.array_equal(x0.view(np.uint32), x1.view(np.uint32)) and np.array_equal(y0.view(np.uint32), y1.view(np.uint32))


def test_oracle_icc_jitter_comes_from_the_scenario_stream_and_is_clamped():
    nc, n, tau2, sd = 8, 32, 0.25, 0.5
    spec = lme_only(S.lmm_spec(nc, 0.2), "normal", 0.0, sd)
    spec["cluster"]["tau_squared"] = tau2
    spec["effect_sizes"] = [0.0, 0.0, 0.0]
    clamped = 0
    for sim in range(24):
        # no other scenario knob is live, so the jitter is the first normal of the scenario stream (data_gen.rs:1070-1075)
        z = orc.seq_stream(SEED, sim ^ SCEN_DOMAIN, "normal", 1).view(np.float32)[0]
        eff = max(tau2 + float(z) * sd, 0.0)
        clamped += eff == 0.0
        draws = orc.draw_residual(SEED, sim, "normal", 0.0, nc)
        u = (draws * np.sqrt(eff)).astype(np.float32).astype(np.float64)
        diff = cluster_effects(spec, SEED, sim, n)
        assert np.allclose(diff, u[np.arange(n) % nc], atol=2e-6), sim
    assert clamped >= 1  # P(z < -0.5) = 0.31 per study: the clamp branch is exercised


def test_oracle_heavy_tailed_random_effects_have_t_kurtosis_and_unit_variance():
    nc, n, df = 30, 60, 10.0
    spec = lme_only(S.lmm_spec(nc, 0.2), "high_kurtosis", df)
    spec["cluster"]["tau_squared"] = 1.0
    spec["effect_sizes"] = [0.0, 0.0, 0.0]
    u = np.concatenate([cluster_effects(spec, SEED, sim, n)[:nc] for sim in range(400)])
    assert abs(u.var() - 1.0) < 0.06                     # draw_residual standardises: tau² is preserved
    kurt = np.mean((u - u.mean()) ** 4) / u.var() ** 2 - 3.0
    assert 0.45 < kurt < 1.9, kurt                        # t(10): 6 / (df - 4) = 1


def _clustered(sizing, between, inter=False, factor=False):
    s = S.base_spec(2)
    s["effect_sizes"] = [0.0, 0.5, 0.3]
    if factor:
        s = S.add_factor(s, [0.4, 0.35, 0.25], [0.3, -0.2])
    if inter:
        s = S.add_interaction(s, [1, 2], 0.2)
    s["target_indices"] = list(range(1, S.n_predictors(s)))
    s["estimator"] = "mle"
    s["cluster"] = {"sizing": sizing, "tau_squared": 0.2, "slopes": [], "extra_groupings": []}
    s["between_var_indices"] = list(between)
    return s


def _check_broadcast(x, sizing, between, n):
    if "FixedClusters" in sizing:
        cid = np.arange(n) % sizing["FixedClusters"]["n_clusters"]
    else:
        cid = np.arange(n) // sizing["FixedSize"]["cluster_size"]
    for col in range(1, x.shape[1]):
        within_const = all(len(set(x[cid == c, col].tolist())) == 1 for c in set(cid.tolist()))
        if col in between:
            assert within_const, f"marked column {col} must be constant within a cluster"
        elif x[:, col].std() > 0:
            pass
    return cid


@pytest.mark.parametrize("sizing", [{"FixedClusters": {"n_clusters": 4}}, {"FixedSize": {"cluster_size": 7}}])
def test_oracle_between_columns_follow_the_reference_properties(sizing):
    n = 40
    # DGEN-CL1: an empty list broadcasts nothing
    x, _, _ = orc.generate_sim(_clustered(sizing, []), SEED, 0, n)
    cid = _check_broadcast(x, sizing, [], n)
    assert len(set(x[cid == 0, 1].tolist())) > 1
    # DGEN-CL2: marked continuous column constant within, varying across clusters; the unmarked one still varies within
    x, _, _ = orc.generate_sim(_clustered(sizing, [1]), SEED, 0, n)
    _check_broadcast(x, sizing, [1], n)
    assert len({x[cid == c, 1][0] for c in set(cid.tolist())}) > 1
    assert len(set(x[cid == 0, 2].tolist())) > 1
    # the representative row keeps its own draw: it equals the unbroadcast design's value
    x_plain, _, _ = orc.generate_sim(_clustered(sizing, []), SEED, 0, n)
    rep = np.array([np.flatnonzero(cid == c)[0] for c in cid])
    assert np.array_equal(x[:, 1], x_plain[rep, 1]) and np.array_equal(x[:, 2], x_plain[:, 2])
    # DGEN-CL3: a cross-level interaction multiplies the BROADCAST value
    x, _, _ = orc.generate_sim(_clustered(sizing, [1], inter=True), SEED, 0, n)
    assert np.array_equal(x[:, 3], (x[:, 1].astype(np.float64) * x[:, 2].astype(np.float64)).astype(np.float32))
    # DGEN-CL4: a marked factor broadcasts one valid one-hot pattern per cluster
    x, _, _ = orc.generate_sim(_clustered(sizing, [3, 4], factor=True), SEED, 0, n)
    _check_broadcast(x, sizing, [3, 4], n)
    assert set(np.unique(x[:, 3:5]).tolist()) <= {0.0, 1.0} and (x[:, 3] + x[:, 4]).max() <= 1.0


def test_between_index_out_of_range_is_rejected_without_a_gpu():
    s = _clustered({"FixedClusters": {"n_clusters": 4}}, [3])
    with pytest.raises(M.InvalidSpec, match="between_var_indices"):
        M.run_batch(s, [40], 4, 1)
    with pytest.raises(orc.OracleError):
        orc.generate_sim(s, 1, 0, 40)


# ---------------------------------------------------------------------------------------------------------- GPU parity
def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.gpu
@pytest.mark.parametrize("dist,df,icc", [("high_kurtosis", 5.0, 0.3), ("high_kurtosis", 10.0, 0.15), ("right_skewed", 4.0, 0.0),
                                          ("left_skewed", 7.0, 0.2), ("uniform", 0.0, 0.4), ("normal", 0.0, 0.5)])
def test_gpu_generated_design_with_lme_knobs_is_bit_exact(dist, df, icc):
    spec = lme_only(S.lmm_spec(9, 0.25), dist, df, icc)
    spec["cluster"]["sizing"] = {"FixedSize": {"cluster_size": 6}}
    for sim in (0, 5, 1234567):
        xg, yg = M.generate_sim(spec, SEED, sim, 57)   # trailing partial cluster: rows 54..56 have no drawn effect
        xo, yo, _ = orc.generate_sim(spec, SEED, sim, 57)
        assert np.array_equal(_bits(xg), _bits(xo)) and np.array_equal(_bits(yg), _bits(yo)), (dist, sim)


@pytest.mark.gpu
@pytest.mark.parametrize("preset", ["realistic", "doomer"])
@pytest.mark.parametrize("layout", ["fixed_clusters_30", "fixed_size_8", "many_clusters"])
def test_gpu_config5_under_the_stock_presets_matches_the_oracle(preset, layout):
    """BASELINE config 5 under the reference's default non-optimistic presets (configs/scenarios.json): heavy-tailed random
    effects (t kernel, df 10 / 5), ICC jitter 0.15 / 0.3 — on the shared-memory generator, and (many clusters) on the
    one-thread-per-study generator."""
    s = S.lmm_spec(30, 0.2, preset=preset)
    grid = [60, 120, 300]
    if layout == "fixed_size_8":
        s["cluster"]["sizing"] = {"FixedSize": {"cluster_size": 8}}
        grid = [40, 100, 164]
    elif layout == "many_clusters":
        s["cluster"]["sizing"] = {"FixedClusters": {"n_clusters": 150}}
        grid = [300, 450]
    assert s["scenario"]["lme"]["random_effect_dist"] == "high_kurtosis"
    got = M.run_batch(s, grid, 320, SEED, mode=ALL)
    want = orc.run_batch(s, grid, 320, SEED, n_threads=orc.host_threads(), capture=True)
    for key in ("uncorrected", "corrected", "converged", "joint", "boundary_hit"):
        assert np.array_equal(getattr(got, key), want[key]), key
    assert np.array_equal(got.t_sq, want["t_sq"], equal_nan=True)
    agg = M.counters_from_bytes(got.uncorrected, got.corrected, got.converged, got.overall, got.factor_excluded, got.joint,
                                got.boundary_hit)
    assert np.array_equal(agg, got.counters)
    assert got.n_converged.min() > 0.8 * 320


@pytest.mark.gpu
def test_gpu_clustered_contract_under_the_realistic_preset_runs():
    """The hosts' payload for a random-intercept model carries the preset's `lme` block (project_contract.rs:972-980); it used
    to come back UNSUPPORTED."""
    import msgpack
    import test_contract_lowering as TC
    scen = TC.scenario(name="realistic", heterogeneity=0.2, heteroskedasticity_ratio=2.0, correlation_noise_sd=0.15,
                       distribution_change_prob=0.5, new_distributions=["RightSkewed", "LeftSkewed", "Uniform"],
                       residual_change_prob=0.5, residual_dists=["high_kurtosis", "right_skewed"], residual_df=8.0,
                       sampled_factor_proportions=True,
                       lme={"random_effect_dist": "high_kurtosis", "random_effect_df": 10.0, "icc_noise_sd": 0.15})
    c = TC.contract([TC.normal(), TC.normal()], [TC.CONST, TC.direct(0), TC.direct(1)], [0.0, 0.4, 0.25],
                    [TC.marginal(1), TC.marginal(2)], estimator="mle", scen=scen,
                    cluster={"sizing": {"FixedClusters": {"n_clusters": 30}}, "tau_squared": 0.25, "slopes": [], "extra_groupings": []})
    blob = msgpack.packb([c], use_bin_type=True)
    low = M.lower_contracts(blob)[0]
    assert low["spec"]["scenario"]["lme"]["icc_noise_sd"] == 0.15
    got = M.find_power_contracts(blob, 300, 1024, 2137)
    pr = got["scenarios"][0][1]
    want = orc.run_batch(low["spec"], [300], 1024, M.splitmix64(2137), n_threads=orc.host_threads())
    assert pr["success_counts_uncorrected"] == [int(v) for v in want["uncorrected"].sum(axis=0)[0]]
    assert pr["convergence_count"] == int(want["converged"].sum())


@pytest.mark.gpu
@pytest.mark.parametrize("sizing", [{"FixedClusters": {"n_clusters": 5}}, {"FixedSize": {"cluster_size": 6}}])
@pytest.mark.parametrize("shape", ["continuous", "factor_and_interaction"])
def test_gpu_between_columns_bit_exact(sizing, shape):
    if shape == "continuous":
        s = _clustered(sizing, [1])
    else:
        s = _clustered(sizing, [2, 3, 4], inter=True, factor=True)
        s["factor_sampled"] = [True]
        s["factor_min_level_count"] = 5
    for sim in (0, 3):
        xg, yg = M.generate_sim(s, SEED, sim, 53)
        xo, yo, _ = orc.generate_sim(s, SEED, sim, 53)
        assert np.array_equal(_bits(xg), _bits(xo)) and np.array_equal(_bits(yg), _bits(yo))
    grid = [30, 53] if "FixedClusters" in sizing else [30, 54]
    got = M.run_batch(s, grid, 320, SEED, mode=ALL)   # mixed-model arm: the quad generator with cluster sums in shared memory
    want = orc.run_batch(s, grid, 320, SEED, n_threads=orc.host_threads(), capture=True)
    for key in ("uncorrected", "corrected", "converged", "joint", "boundary_hit", "factor_excluded"):
        if want[key] is not None:
            assert np.array_equal(getattr(got, key), want[key]), key
    # generate clustered, solve OLS: every OLS power kernel sees the same broadcast design
    o = dict(s, estimator="ols")
    want = orc.run_batch(o, grid, 320, SEED, n_threads=orc.host_threads(), capture=True)
    for variant in (1, 4, 6):
        prev = M.force_ols_variant(variant)
        try:
            got = M.run_batch(o, grid, 320, SEED, mode=ALL)
        finally:
            M.force_ols_variant(prev)
        for key in ("uncorrected", "corrected", "converged", "factor_excluded"):
            if want[key] is not None:
                assert np.array_equal(getattr(got, key), want[key]), (variant, key)
        assert np.array_equal(got.t_sq, want["t_sq"], equal_nan=True), variant
