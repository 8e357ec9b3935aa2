"""Pin the CPU oracle against every reference vector this path has.

Bit-exact: Philox / splitmix64 / u32→f32 KATs, the sequential and scenario stream
goldens, draw_residual goldens, the exact-form marginals, and the sim-0 design
matrices + outcomes of all 39 validation datasets (OLS, logit, probit, Poisson, LME).
Value-pinned: R lm goldens (≤1e-10 rel; the reference's own gate is 1e-11,
validation/tolerances.R:24-28), R contrast t² (posthoc.rs:300-317), quantile goldens
(critvals.rs:1807-1925), phi-composed transforms (≤2 ULP — the reference's exp kernel
lives in the un-vendored glmm crate and is only specified to "≤1 ULP of libm").
"""
import ctypes as C
import json
import math

import numpy as np
import pytest

import _cases as CS
from oracle import orc

G = json.loads((CS.GOLDEN / "rng_goldens.json").read_text())


def ulp_diff64(a: float, b_bits: int) -> int:
    a_bits = int(np.float64(a).view(np.uint64))

    def key(u):
        return u if u < (1 << 63) else (1 << 63) - u - 1 if False else -(u - (1 << 63))
    return abs(key(a_bits) - key(b_bits))


# ---------------------------------------------------------------- RNG primitives
def test_philox_known_answers():  # tests/rng_byte_identity.rs:17-33
    assert orc.philox([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert orc.philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert orc.philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_splitmix64_pinned():  # tests/rng_byte_identity.rs:36-42
    assert orc.splitmix64(0) == 0
    assert orc.splitmix64(1) == 0x5692161D100B05E5
    assert orc.splitmix64(0x0123456789ABCDEF) == 0xB2C058E4EBB5112C
    assert orc.splitmix64(0xFFFFFFFFFFFFFFFF) == 0xB4D055FCF2CBBD7B


def test_u32_to_unit_f32_pinned():  # tests/rng_byte_identity.rs:45-50
    f = orc.lib().orc_u32_to_unit_f32
    for w, bits in ((0, 0x33800000), (0xFFFFFFFF, 0x3F7FFFFF), (0x9E3779B9, 0x3F1E3779)):
        assert int(np.float32(f(w)).view(np.uint32)) == bits


def test_golden_streams():  # tests/golden_rng.rs:177-206
    assert orc.seq_stream(42, 0, "normal", 64).tolist() == G["golden_normals"]
    assert orc.seq_stream(42, 1, "uniform", 64).tolist() == G["golden_uniforms"]
    assert orc.seq_categorical(42, 2, [0.2, 0.3, 0.5], 64).tolist() == G["golden_categoricals"]
    assert orc.seq_stream(42, 0 ^ 0x5C5C5C5C5C5C5C5C, "uniform", 64).tolist() == G["golden_scenario_u32s"]


def test_golden_sample_t():  # tests/golden_rng.rs:226-235
    assert orc.sample_t(42, 3, 5.0, 16).view(np.uint64).tolist() == G["golden_sample_t_df5"]
    assert orc.sample_t(42, 4, 30.0, 16).view(np.uint64).tolist() == G["golden_sample_t_df30"]


def test_golden_draw_residual():  # data_gen.rs:2446-2505
    names = {"Normal": "normal", "Uniform": "uniform", "HighKurtosis": "high_kurtosis",
             "RightSkewed": "right_skewed", "LeftSkewed": "left_skewed"}
    assert len(G["draw_residual"]["rows"]) == 5
    for row in G["draw_residual"]["rows"]:
        got = orc.draw_residual(42, row["sim_id"], names[row["dist"]], 8.0, 8).view(np.uint64).tolist()
        assert got == row["bits"], row["dist"]


def test_planar_matches_sequential_kernel():
    # rng.rs:230-261: planar fill is bit-identical to the scalar kernel on word(class, col, row)
    key = orc.lib().orc_pcg_mix64(7, 3)
    z = orc.planar_normals(7, 3, 1, 2, 13)
    for i in range(13):
        w = orc.philox([i >> 2, 2, 1, 0], [key & 0xFFFFFFFF, key >> 32])[i & 3]
        u = orc.lib().orc_u32_to_unit_f32(w)
        assert np.float32(orc.lib().orc_norm_inv_cdf_f32(u)) == z[i]


# ---------------------------------------------------------------- transforms
def test_golden_phi_within_2ulp():  # tests/golden_rng.rs:208-212
    exact = 0
    for z, bits in zip(G["z_grid"], G["golden_phi"]):
        d = ulp_diff64(orc.lib().orc_phi(z), bits)
        # 1 - erf cancels at z = -8: the golden there is 2^-51-quantised; allow the same absolute slack
        assert d <= 2 or abs(orc.lib().orc_phi(z) - np.uint64(bits).view(np.float64)) < 3e-16, (z, d)
        exact += d == 0
    assert exact >= 8


def test_golden_marginal_uniform_within_2ulp():  # tests/golden_rng.rs:214-223
    for z, bits in zip(G["z_grid"], G["golden_marginal_uniform"]):
        got = orc.lib().orc_apply_marginal(z, orc.DIST["uniform"], 0.0)
        want = float(np.uint64(bits).view(np.float64))
        assert abs(got - want) <= 4e-16 * max(1.0, abs(want)), (z, got, want)


def test_golden_apply_marginal():  # data_gen.rs:2282-2400
    names = {"Normal": "normal", "Binary": "binary", "RightSkewed": "right_skewed", "LeftSkewed": "left_skewed",
             "HighKurtosis": "high_kurtosis", "Uniform": "uniform", "UploadedData": "uploaded_data",
             "UploadedBinary": "uploaded_binary", "UploadedFactor": "uploaded_factor"}
    zs = G["apply_marginal"]["z_grid"]
    # the fabricated sorted upload table of the reference test, shape (5, 2); var_idx = 1 reads [-2, -1, 0, 1, 2]
    table = np.array([-1.5, -2.0, -0.5, -1.0, 0.0, 0.0, 0.5, 1.0, 1.5, 2.0])
    tp = table.ctypes.data_as(C.POINTER(C.c_double))
    seen = 0
    for row in G["apply_marginal"]["rows"]:
        seen += 1
        for z, bits in zip(zs, row["bits"]):
            got = orc.lib().orc_apply_marginal_upload(z, orc.DIST[names[row["dist"]]], row["param"], tp, 5, 2, 1)
            want = float(np.uint64(bits).view(np.float64))
            if row["dist"] in ("Normal", "Binary", "UploadedBinary", "UploadedFactor"):
                assert int(np.float64(got).view(np.uint64)) == bits or got == want
            else:
                # phi/ln-composed: value-pinned; |z|=8 rows sit in phi's cancellation region
                assert abs(got - want) <= 1e-9 * max(1.0, abs(want)), (row["dist"], z, got, want)
    assert seen == 9


def test_own_exp_log_accuracy():
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.uniform(-700, 700, 4000), rng.uniform(-2, 2, 4000), [0.0, -0.0, 1e-300, -745.0, 709.0]])
    for x in xs:
        got, want = orc.lib().orc_exp(float(x)), math.exp(float(x))
        assert abs(got - want) <= 2.5e-16 * want + 5e-324, x
    for x in np.concatenate([np.exp(rng.uniform(-700, 700, 4000)), rng.uniform(0.5, 2, 4000), [1.0, 5e-324, 1e308]]):
        got, want = orc.lib().orc_log(float(x)), math.log(float(x))
        assert abs(got - want) <= 2.5e-16 * abs(want) + 1e-17, x


# ---------------------------------------------------------------- critical values
def test_quantile_goldens():  # critvals.rs:1807-1925 (R qnorm/qt/qchisq/qf/qtukey)
    L = orc.lib()
    for p, want in ((0.975, 1.959963985), (0.95, 1.644853627), (0.025, -1.959963985), (0.999, 3.090232306)):
        assert abs(L.orc_norm_ppf(p) - want) < 2e-9
    for p, df, want in ((0.975, 30.0, 2.042272), (0.975, 5.0, 2.570582), (0.975, 1000.0, 1.962339),
                        (0.995, 10.0, 3.169273)):
        assert abs(L.orc_t_ppf(p, df) - want) / want < 1e-5
    for p, k, want in ((0.95, 1.0, 3.841459), (0.95, 2.0, 5.991465), (0.95, 5.0, 11.070498), (0.99, 3.0, 11.344867)):
        assert abs(L.orc_chi2_ppf(p, k) - want) / want < 1e-5
    for p, a, b, want in ((0.95, 3.0, 96.0, 2.699393), (0.95, 1.0, 30.0, 4.170877), (0.99, 2.0, 50.0, 5.056611)):
        assert abs(L.orc_f_ppf(p, a, b) - want) / want < 1e-5
    assert abs(L.orc_q_tukey_ppf(0.95, 3.0, 1e6) - 3.314493) < 0.01
    assert abs(L.orc_q_tukey_ppf(0.95, 4.0, 30.0) - 3.845401) < 0.01


def test_quantiles_against_scipy():
    from scipy import stats
    L = orc.lib()
    for df in (1, 2, 3, 9, 11, 30, 201, 491, 989, 5000):
        for p in (0.975, 0.9975, 1 - 2.5e-7, 1 - 0.05 / 22):
            assert abs(L.orc_t_ppf(p, float(df)) / stats.t.ppf(p, df) - 1) < 1e-9, (p, df)
    for dfn, dfd in ((1, 10), (8, 491), (10, 9), (10, 989)):
        assert abs(L.orc_f_ppf(0.95, float(dfn), float(dfd)) / stats.f.ppf(0.95, dfn, dfd) - 1) < 1e-9
    assert abs(L.orc_chi2_ppf(0.95, 10.0) / stats.chi2.ppf(0.95, 10) - 1) < 1e-9


def test_reduced_crit_table_known_quantiles():  # batch.rs:5035-5055
    from mcpower_b200 import spec as S
    s = S.add_factor(S.base_spec(0), [0.92, 0.02, 0.02, 0.02, 0.02], [0.0] * 4)
    s["target_indices"] = [1]
    ct = orc.crit_table(s, [12], n_predictors=2)
    assert abs(ct["t_crit_sq_uncorrected"][0] - 4.9646027) < 1e-5
    assert abs(ct["overall_crit"][0] - ct["t_crit_sq_uncorrected"][0]) < 1e-7  # t²(ν) = F(1, ν)


def test_crit_table_df_guard():  # critvals.rs:874-879
    from mcpower_b200 import spec as S
    with pytest.raises(orc.OracleError):
        orc.crit_table(S.base_spec(3), [4])


def test_correction_rows_monotone():
    from mcpower_b200 import spec as S
    s = S.sweep_spec()
    for method in ("bonferroni", "holm", "benjamini_hochberg"):
        s["correction_method"] = method
        ct = orc.crit_table(s, [50, 200])
        row = ct["correction_t_crit_sq"][0]
        assert np.all(row >= ct["t_crit_sq_uncorrected"][0] - 1e-12)
        if method == "holm":
            assert np.all(np.diff(row) <= 0)  # alpha/(2(m-k)) grows with k
        if method == "benjamini_hochberg":
            assert np.all(np.diff(row) <= 0) and abs(row[-1] - ct["t_crit_sq_uncorrected"][0]) < 1e-9


# ---------------------------------------------------------------- corrections (correction.rs tests)
def test_apply_correction_semantics():
    t = np.array([9.0, np.nan, 4.5, 4.5, 1.0])
    assert orc.apply_correction("none", t, 4.0, np.zeros(5)).tolist() == [1, 0, 1, 1, 0]
    assert orc.apply_correction("bonferroni", t, 4.0, np.full(5, 5.0)).tolist() == [1, 0, 0, 0, 0]
    # Holm: ranks 9.0 > 8, 4.5 > 4.4, 4.5 !> 4.6 → stop
    assert orc.apply_correction("holm", t, 4.0, np.array([8.0, 4.4, 4.6, 0.5, 0.1])).tolist() == [1, 0, 1, 0, 0]
    # BH: last rank passing is k=3 (1.0 > 0.5) → ranks 0..3 pass even though rank 2 failed its own threshold
    assert orc.apply_correction("benjamini_hochberg", t, 4.0, np.array([8.0, 4.4, 4.6, 0.5, 0.1])).tolist() == \
        [1, 0, 1, 1, 1]


# ---------------------------------------------------------------- DGP: the 39 saved sim-0 datasets
@pytest.mark.parametrize("name", sorted(CS.validation_cases()))
def test_dgp_bit_exact_vs_reference_dataset(name):
    case = CS.validation_cases()[name]
    spec = CS.case_spec(case)
    seed = orc.splitmix64(CS.case_seed(case))  # debug.rs:343-349
    x, y, ids = orc.generate_sim(spec, seed, 0, case["n"])
    xr, yr = CS.case_arrays(case)
    assert np.array_equal(x.view(np.uint32), xr.view(np.uint32)), "design matrix bits differ"
    assert np.array_equal(y.view(np.uint32), yr.view(np.uint32)), "outcome bits differ"
    if case.get("cluster_ids"):
        assert ids.tolist() == [c - min(case["cluster_ids"]) for c in case["cluster_ids"]]


def test_dgp_prefix_stable_in_max_n():  # rng.rs:3-6 / data_gen.rs:3-6
    case = CS.validation_cases()["ols_cf_a"]
    spec = CS.case_spec(case)
    x1, y1, _ = orc.generate_sim(spec, 99, 5, 64)
    x2, y2, _ = orc.generate_sim(spec, 99, 5, 301)
    assert np.array_equal(x1, x2[:64]) and np.array_equal(y1, y2[:64])


# ---------------------------------------------------------------- OLS solve vs frozen R lm
@pytest.mark.parametrize("name", sorted(n for n in CS.validation_cases() if n.startswith("ols_")))
def test_ols_fit_matches_r_lm(name):
    case = CS.validation_cases()[name]
    x, y = CS.case_arrays(case)
    g = CS.golden_arrays(case)
    fit = orc.fit_ols(x.astype(np.float64), y.astype(np.float64))
    assert fit["converged"]
    stat = fit["betas"] / fit["se"]
    for got, want, what in ((fit["betas"], g["beta"], "beta"), (fit["se"], g["se"], "se"), (stat, g["stat"], "t")):
        rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-300)
        assert rel.max() < 1e-10, (what, rel.max())
    spec = CS.case_spec(case)
    ct = orc.crit_table(spec, [case["n"]])
    assert abs(math.sqrt(ct["t_crit_sq_uncorrected"][0]) / g["crit"][0] - 1) < 1e-11  # tolerances.R:32-37


def test_posthoc_contrast_golden():  # posthoc.rs:283-437 (R lm contrast t², ±1e-5)
    f = G["posthoc_lm"]
    n = 12
    x = np.zeros((n, 4))
    x[:, 0] = 1.0
    x[:, 1] = f["x_cont"]
    x[4:8, 2] = 1.0
    x[8:12, 3] = 1.0
    # fit_ols narrows to f32 (introspect.rs:195-203); the fixture is f64, so use a tolerance that
    # covers the narrowing (1e-6) as well as the reference's own 1e-5 sandwich
    fit = orc.fit_ols(x, np.array(f["y"]))
    assert abs(fit["t_sq"][2] / f["tsq_b"] - 1) < 1e-5
    assert abs(fit["t_sq"][3] / f["tsq_c"] - 1) < 1e-5


def test_posthoc_contrasts_reduce_to_target_and_contrast_tests():
    """build_posthoc_contrasts (batch.rs:2183-2227): level-vs-reference is the unit contrast on the dummy, level-vs-level
    is e_b - e_a.  With posthoc_alpha = alpha and equal family sizes the posthoc decisions must therefore equal the
    marginal-target / contrast_pairs decisions on the same columns (same t², same thresholds)."""
    s = CS.posthoc_spec("bonferroni")
    s["factor_n_levels"], s["factor_proportions"], s["n_factor_dummies"] = [3], [0.5, 0.3, 0.2], 2
    s["effect_sizes"] = [0.0, 0.25, 0.1, 0.5, -0.4]
    s["target_indices"], s["contrast_pairs"] = [3, 4], [[4, 3]]
    s["posthoc"] = [{"factor_index": 0, "target_indices": [3, 4]}]
    s["crit_values"] = {"alpha": 0.05, "posthoc_alpha": None}
    r = orc.run_batch(s, [30, 90], 400, 0x1234, n_threads=orc.host_threads())
    assert r["posthoc_unc"].shape == (400, 2, 3)
    assert np.array_equal(r["posthoc_unc"], r["uncorrected"]) and np.array_equal(r["posthoc_cor"], r["corrected"])
    assert r["posthoc_unc"].any() and not r["posthoc_unc"].all()


def test_posthoc_correction_rows_per_block():
    """build_posthoc_correction_crit_with_levels (batch.rs:2252-2310): one table per block at posthoc alpha, family =
    the block's C(k,2) contrasts, Tukey k = the block's level count; rows concatenated in block order."""
    import ctypes as C
    for method in ("bonferroni", "holm", "benjamini_hochberg", "tukey_hsd"):
        ps = orc.PackedSpec(CS.posthoc_spec(method))
        ss = np.array([40, 200], dtype=np.uint32)
        out = np.zeros((2, 9))
        assert orc.lib().orc_posthoc_corr_rows(C.byref(ps.c), ss.ctypes.data_as(C.POINTER(C.c_uint32)), 2, ps.P,
                                               out.ctypes.data_as(C.POINTER(C.c_double))) == 0
        for i, n in enumerate((40, 200)):
            df = n - ps.P
            for off, m, k in ((0, 3, 3), (3, 6, 4)):
                row = out[i, off:off + m]
                if method == "bonferroni":
                    want = [orc.lib().orc_t_ppf(1 - 0.01 / (2 * m), df) ** 2] * m
                elif method == "holm":
                    want = [orc.lib().orc_t_ppf(1 - 0.01 / (2 * (m - j)), df) ** 2 for j in range(m)]
                elif method == "benjamini_hochberg":
                    want = [orc.lib().orc_t_ppf(1 - (j + 1) / m * 0.01 / 2, df) ** 2 for j in range(m)]
                else:
                    want = [(orc.lib().orc_q_tukey_ppf(1 - 0.01, k, df) / math.sqrt(2)) ** 2] * m
                assert np.allclose(row, want, rtol=1e-13), (method, n, off)


# ---------------------------------------------------------------- GLM IRLS vs frozen R glm (1e-4 gate, tolerances.R:29-31)
@pytest.mark.parametrize("name", sorted(n for n in CS.validation_cases()
                                        if n.startswith(("glm_", "probit_", "poisson_"))))
def test_glm_fit_matches_r_glm(name):
    case = CS.validation_cases()[name]
    x, y = CS.case_arrays(case)
    g = CS.golden_arrays(case)
    fam = {"glm": "logit", "probit": "probit", "poisson": "poisson"}[name.split("_")[0]]
    fit = orc.fit_glm(x.astype(np.float64), y.astype(np.float64), fam)
    assert fit["converged"]
    assert np.max(np.abs(fit["betas"] - g["beta"]) / np.abs(g["beta"])) < 1e-4
    assert np.max(np.abs(fit["se"] - g["se"]) / np.abs(g["se"])) < 1e-4


# ---------------------------------------------------------------- batch driver behaviour (batch.rs tests)
def test_run_batch_validation_errors():  # batch.rs:236-253
    from mcpower_b200 import spec as S
    s = S.headline_spec()
    for bad in ([], [0, 10], [50, 50], [60, 40]):
        with pytest.raises(orc.OracleError):
            orc.run_batch(s, bad, 4, 1)


def test_run_batch_threads_and_offsets_are_invariant():  # batch.rs:2715-2786
    from mcpower_b200 import spec as S
    s = S.sweep_spec()
    a = orc.run_batch(s, [20, 60, 200], 96, 123, n_threads=1, capture=True)
    b = orc.run_batch(s, [20, 60, 200], 96, 123, n_threads=4, capture=True)
    c = orc.run_batch(s, [20, 60, 200], 40, 123, sim_begin=56, n_threads=2, capture=True)
    for k in ("uncorrected", "corrected", "converged", "overall", "factor_excluded"):
        assert np.array_equal(a[k], b[k]), k
        assert np.array_equal(a[k][56:], c[k]), k
    assert np.array_equal(a["t_sq"], b["t_sq"], equal_nan=True)


def test_sparse_factor_exclusion_at_small_n():  # batch.rs:4164-4286, SURVEY §7 "hard parts"
    from mcpower_b200 import spec as S
    s = S.sweep_spec()
    r = orc.run_batch(s, [20, 40], 8, 5, capture=True)
    # exact walk (0.5,0.3,0.2) at n=20 → counts (10,6,4): level 3 < 5 ⇒ factor dropped, its targets report 0/NaN
    assert r["factor_excluded"][:, 0, 0].tolist() == [1] * 8
    assert r["factor_excluded"][:, 1, 0].tolist() == [0] * 8
    assert np.all(np.isnan(r["t_sq"][:, 0, 8:10])) and not np.any(np.isnan(r["t_sq"][:, 1, :]))
    assert np.all(r["uncorrected"][:, 0, 8:10] == 0)


def test_power_matches_noncentral_t():
    """Headline spec, n=500: empirical power of x1 vs the analytic non-central t power (3 SE band)."""
    from scipy import stats
    from mcpower_b200 import spec as S
    s = S.headline_spec()
    n, n_sims = 500, 4000
    r = orc.run_batch(s, [n], n_sims, orc.splitmix64(2137), n_threads=orc.host_threads())
    power = CS.analytic_ols_power(0.18, n, 9)
    emp = r["uncorrected"][:, 0, 0].mean()
    assert abs(emp - power) < 4 * math.sqrt(power * (1 - power) / n_sims) + 1e-4, (emp, power)


# ---------------------------------------------------------------- end-to-end vs the SHIPPED ENGINE's own outputs
@pytest.mark.parametrize("name", ["cross_ols", "cross_partial", "cross_lme", "cross_glm"])
def test_oracle_reproduces_the_engines_frozen_required_n(name):
    """validation/data/cross_*.golden.rds: `find_sample_size(20..300 by 2, 100000 sims, seed 2137)` run by the reference
    engine through its public API (validation_crossing.rmd:112-135).  The oracle's per-cell corrected counts, folded by the
    product's host-side crossing fit (fit.rs twin, itself pinned by the reference's KATs in test_orchestrator_host.py),
    must give the same status and required N for every target — RNG, data generation, OLS / IRLS / REML fits, decisions,
    aggregation, grid snapping and crossing fit, end to end."""
    import mcpower_b200 as M
    g = CS.crossing_goldens()[name]
    s = CS.crossing_spec(g["case"])
    atom = g["case"].get("cluster", {}).get("n_clusters", 1)
    grid, _ = M.build_grid(g["from_size"], g["to_size"], by=g["by"], atom=atom, hard_min=2 * atom if atom > 1 else 1)
    r = orc.run_batch(s, grid, g["n_sims"], orc.splitmix64(g["seed"]), n_threads=orc.host_threads())
    counts = r["corrected"].sum(axis=0)
    for t, want in enumerate(g["dense"].values()):
        fit = M.fit_crossing(grid, counts[:, t], g["n_sims"], 80.0, atom)
        assert fit["status"] == want["status"], (name, t, fit)
        if want["status"] == "fitted":
            assert fit["n_achievable"] == want["n_achievable"], (name, t, fit)


# ---------------------------------------------------------------- uploaded-data arms
def _nonlinear_frame():  # tests/test_upload_bootstrap.rs:88-96 — x2 = x1², impossible under a Gaussian copula
    return [[-2.0 + 4.0 * i / 39.0, (-2.0 + 4.0 * i / 39.0) ** 2] for i in range(40)]


def _strict_2cont_spec(frame):  # tests/test_upload_bootstrap.rs:29-83
    from mcpower_b200 import spec as S
    s = S.base_spec(2)
    s["effect_sizes"] = [0.0, 0.5, 0.3]
    return S.with_uploaded_frame(s, frame, {0: 0, 1: 1}, bootstrap=True)


def test_strict_bootstrap_copies_whole_rows():  # tests/test_upload_bootstrap.rs:104-131
    frame = _nonlinear_frame()
    x, _, _ = orc.generate_sim(_strict_2cont_spec(frame), 2137, 0, 500)
    f32 = np.asarray(frame, dtype=np.float64).astype(np.float32)
    for i in range(500):
        assert np.any((f32[:, 0] == x[i, 1]) & (f32[:, 1] == x[i, 2])), i
    assert len({float(v) for v in x[:, 1]}) > 20  # rows are picked all over the frame, not one row repeated


def test_strict_bootstrap_is_row_stable():  # tests/test_upload_bootstrap.rs:136-160
    s = _strict_2cont_spec(_nonlinear_frame())
    xs, ys, _ = orc.generate_sim(s, 2137, 5, 100)
    xb, yb, _ = orc.generate_sim(s, 2137, 5, 1000)
    assert np.array_equal(xs, xb[:100]) and np.array_equal(ys, yb[:100])


def test_strict_bootstrap_factor_reexpands_level():  # tests/test_upload_bootstrap.rs:229-300
    from mcpower_b200 import spec as S
    frame = [[i * 0.1 - 1.5, float(i % 3)] for i in range(30)]
    s = S.add_factor(S.base_spec(1), [0.5, 0.3, 0.2], [0.0, 0.0])
    s["effect_sizes"][1], s["target_indices"] = 0.5, [1]
    s = S.with_uploaded_frame(s, frame, {0: 0}, factors={0: 1}, bootstrap=True)
    x, _, _ = orc.generate_sim(s, 2137, 0, 500)
    src = {(np.float32(r[0]).item(), int(r[1])) for r in frame}
    for i in range(500):
        d0, d1 = float(x[i, 2]), float(x[i, 3])
        assert (d0, d1) in ((0.0, 0.0), (1.0, 0.0), (0.0, 1.0))
        assert (float(x[i, 1]), 1 if d0 == 1.0 else 2 if d1 == 1.0 else 0) in src, i  # both from the SAME source row


def test_strict_bootstrap_binary_copies_joint_with_continuous():  # tests/test_upload_bootstrap.rs:308-444
    from mcpower_b200 import spec as S
    p = 0.4
    frame = [[-1.0 + 0.05 * i, (1.0 if i % 5 < 2 else 0.0) - p] for i in range(40)]  # binary stored centred: x − p
    s = S.base_spec(2)
    s["effect_sizes"] = [0.0, 0.4, 0.3]
    s = S.with_uploaded_frame(s, frame, {0: 0, 1: (1, p)}, bootstrap=True)
    x, _, _ = orc.generate_sim(s, 2137, 3, 400)
    src = {(np.float32(r[0]).item(), 1.0 if r[1] + p >= 0.5 else 0.0) for r in frame}
    assert set(np.unique(x[:, 2])) == {0.0, 1.0}
    for i in range(400):
        assert (float(x[i, 1]), float(x[i, 2])) in src, i


def test_norta_uploaded_marginal_reproduces_the_empirical_distribution():  # data_gen.rs:1156-1175, 1262-1277
    from mcpower_b200 import spec as S
    rng = np.random.default_rng(5)
    col = rng.exponential(size=200)
    col = (col - col.mean()) / col.std()
    frame = [[float(v), float(i % 2) - 0.5] for i, v in enumerate(col)]
    s = S.base_spec(3)
    s["correlation"] = [1, 0.5, 0, 0.5, 1, 0, 0, 0, 1]
    s["effect_sizes"] = [0.0, 0.2, 0.2, 0.1]
    s = S.with_uploaded_frame(s, frame, {1: 0, 2: (1, 0.5)})
    x, _, _ = orc.generate_sim(s, 2137, 0, 20000)
    z = x[:, 2].astype(np.float64)
    assert col.min() - 1e-6 <= z.min() and z.max() <= col.max() + 1e-6
    for q in (0.1, 0.5, 0.9):
        assert abs(np.quantile(z, q) - np.quantile(col, q)) < 0.06
    assert abs(np.corrcoef(x[:, 1], x[:, 2])[0, 1] - 0.5) < 0.06  # the Gaussian copula carries the correlation (attenuated)
    assert set(np.unique(x[:, 3])) == {0.0, 1.0} and abs(x[:, 3].mean() - 0.5) < 0.02


def test_uploaded_tables_are_validated():
    from mcpower_b200 import spec as S
    s = _strict_2cont_spec(_nonlinear_frame())
    s["bootstrap_frame_map"] = [0, 5]
    with pytest.raises(orc.OracleError):
        orc.generate_sim(s, 1, 0, 10)
    s["bootstrap_frame_map"] = [0]
    with pytest.raises(orc.OracleError):
        orc.generate_sim(s, 1, 0, 10)


# ---------------------------------------------------------------- PSD repair at k = 8 against an independent eigensolver
def test_psd_repair_k8_matches_numpy_eigh():
    """correlation.rs:82-189: clip eigenvalues at 1e-10, rebuild, rescale to a unit diagonal, factor.  The reference uses faer's
    self-adjoint EVD; the oracle (and the CUDA kernel, bit for bit) a cyclic Jacobi.  The reference's vectors only reach k = 3,
    BASELINE config 3 runs k = 8 through the repair for most studies — so the k = 8 result is checked here against LAPACK
    (numpy.linalg.eigh): the repaired correlation L L' must agree to 1e-9 (eigenvector bases may differ, the repaired matrix is
    unique), and untouched studies must reproduce their perturbed correlation to rounding."""
    from mcpower_b200 import spec as S
    s = dict(S.sweep_spec(), scenario=S.scenario("doomer"))
    k, sd = 8, s["scenario"]["correlation_noise_sd"]
    base = np.array(s["correlation"]).reshape(k, k)
    seed, repaired, worst = 0x0BADC0FFEE, 0, 0.0
    for sim in range(200):
        z = orc.seq_stream(seed, sim ^ 0x5C5C5C5C5C5C5C5C, "normal", k * k).view(np.float32).astype(np.float64)
        noise = z.reshape(k, k).T * sd             # noise[i, j] = draw j*k + i (column-major fill, scenarios.rs:67-120)
        a = np.clip(base + 0.5 * (noise + noise.T), -0.8, 0.8)
        np.fill_diagonal(a, 1.0)
        lam, vec = np.linalg.eigh(a)
        if lam.min() > 1e-8:                         # comfortably positive definite: no repair
            want = a
        elif lam.min() < -1e-8:
            rep = (vec * np.maximum(lam, 1e-10)) @ vec.T
            d = np.sqrt(np.diag(rep))
            want = rep / np.outer(d, d)
            np.fill_diagonal(want, 1.0)
            repaired += 1
        else:
            continue                                 # borderline: either branch is legitimate
        L = orc.design_factor(s, seed, sim)
        assert np.allclose(np.triu(L, 1), 0.0)
        got = L @ L.T
        worst = max(worst, float(np.abs(got - want).max()))
    assert repaired >= 60                            # doomer noise (sd 0.3) breaks positive definiteness for most 8 x 8 draws
    assert worst < 1e-9, worst
