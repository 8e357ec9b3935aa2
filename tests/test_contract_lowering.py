"""Contract lowering: the hosts' `Vec<SimulationContract>` payload → the kernel `SimulationSpec`.

Every case restates a known-answer test of the reference's adapter (crates/engine-core/src/contract_adapter.rs:900-2440)
or validator (crates/engine-contract/src/validate.rs): same contract in, same spec fields out.  The contracts are built
as plain dicts in serde's shape and sent both as the msgpack blob `rmp_serde::to_vec_named` produces
(engine-py/src/orchestrator_bridge.rs:23-30) and as JSON; both decoders must lower to the same spec.  Host-side only:
no GPU is needed (the GPU leg lives in tests/test_gpu_parity.py::test_contract_entry_points_match_spec_entry_points)."""
import copy
import json

import msgpack
import pytest

import mcpower_b200 as M

T3 = [1.0 / 3.0] * 3
CONST = "Const"


def direct(c):
    return {"Direct": {"column": c}}


def dummy(c, lv):
    return {"DummyOf": {"column": c, "level_index": lv}}


def inter(*comps):
    return {"Interaction": {"components": list(comps)}}


def normal(pinned=False):
    return {"Synthetic": {"kind": "Normal", "pinned": pinned}}


def factor(n, props, sampled=None):
    return {"FactorSynthetic": {"n_levels": n, "proportions": list(props), "sampled_proportions": sampled}}


def marginal(t):
    return {"kind": "marginal", "term": t}


def scenario(**kw):  # ScenarioPerturbations::default() (scenarios.rs:27-44)
    s = {"name": "", "heterogeneity": 0.0, "heteroskedasticity_ratio": 1.0, "correlation_noise_sd": 0.0,
         "distribution_change_prob": 0.0, "new_distributions": [], "residual_change_prob": 0.0, "residual_dists": [],
         "residual_df": 0.0, "sampled_factor_proportions": False, "truth_start": False, "lme": None}
    s.update(kw)
    return s


def contract(columns, gen_terms, coefficients, targets, *, design_test=None, kind="continuous", intercept=0.0,
             estimator="ols", correction="none", alpha=0.05, posthoc=(), correlations="Identity", cluster=None,
             frame=None, cluster_level_columns=(), residual=("normal", False), hsk_driver=None, link=None, scen=None):
    return {
        "generation": {"columns": list(columns), "correlations": correlations, "cluster": cluster, "uploaded_frame": frame,
                       "cluster_level_columns": list(cluster_level_columns)},
        "design_generation": {"terms": list(gen_terms)},
        "outcome": {"kind": kind, "intercept": intercept, "coefficients": list(coefficients),
                    "residual": {"distribution": residual[0], "pinned": residual[1]},
                    "heteroskedasticity_driver": hsk_driver, "link": link},
        "design_test": None if design_test is None else {"terms": list(design_test)},
        "estimator": estimator, "wald_se": "rx", "nagq": 1,
        "test": {"targets": list(targets), "correction": correction, "alpha": alpha},
        "posthoc": list(posthoc), "scenario": scen or scenario(), "max_failed_fraction": 0.03,
    }


def example1():  # engine-contract/src/fixtures.rs:14-69
    return contract([normal(), normal()], [CONST, direct(0), direct(1)], [0.0, 0.5, 0.3], [marginal(1), marginal(2)])


def lower(c):
    """Lower through BOTH wire formats and insist they agree; return the single scenario."""
    via_json = M.lower_contracts([c])
    via_msgpack = M.lower_contracts(msgpack.packb([c], use_bin_type=True))
    assert via_json == via_msgpack
    assert len(via_json) == 1
    return via_json[0]


def test_example1_round_trips_generation_block():  # contract_adapter.rs:906-921
    s = lower(example1())["spec"]
    assert s["n_non_factor"] == 2 and s["n_factor_dummies"] == 0
    assert s["var_types"] == ["normal", "normal"] and s["var_params"] == [0.0, 0.0]
    assert s["factor_n_levels"] == [] and s["factor_proportions"] == []
    assert s["correlation"] == [1.0, 0.0, 0.0, 1.0]
    assert s["upload_normal_shape"] == [0, 0] and s["upload_data_shape"] == [0, 0]
    assert s["target_indices"] == [1, 2] and s["correction_method"] == "none"  # :1067-1074
    assert s["crit_values"] == {"alpha": 0.05, "posthoc_alpha": None}
    assert s["fit_columns"] == []  # :1149-1158
    assert s["factor_min_level_count"] == 5  # lower_contracts forces the configured limit (find_power.rs:271)


def test_effect_sizes_layout_matches_design_matrix_columns():  # contract_adapter.rs:924-1001
    terms = [CONST, dummy(0, 1), dummy(0, 2), direct(1)]
    c = contract([factor(3, T3), normal()], terms, [0.0, 0.3, 0.5, 0.4], [marginal(3)], design_test=terms)
    sc = lower(c)
    assert sc["spec"]["effect_sizes"] == [0.0, 0.4, 0.3, 0.5]
    assert sc["spec"]["target_indices"] == [1] and sc["report_target_indices"] == [3]


def test_logit_intercept_folds_into_effect_sizes_slot0():  # contract_adapter.rs:1004-1064, 1447-1455
    b0 = -1.3863
    c = contract([normal()], [CONST, direct(0)], [0.0, 0.5], [marginal(1)], kind="binary", intercept=b0, estimator="glm")
    s = lower(c)["spec"]
    assert s["effect_sizes"] == [b0, 0.5] and s["intercept"] == b0 and s["estimator"] == "glm" and s["outcome_kind"] == "binary"


def test_fit_columns_derived_from_reduced_design_test():  # contract_adapter.rs:1077-1146
    c = contract([normal(), normal()], [CONST, direct(0), direct(1)], [0.0, 0.1, 0.8], [marginal(1)], design_test=[CONST, direct(0)])
    assert lower(c)["spec"]["fit_columns"] == [0, 1]


def test_report_indices_when_leading_term_dropped():  # contract_adapter.rs:1228-1256
    c = contract([normal(), normal()], [CONST, direct(0), direct(1)], [0.0, 0.3, 0.0], [marginal(1)], design_test=[CONST, direct(1)])
    sc = lower(c)
    assert sc["spec"]["target_indices"] == [2] and sc["report_target_indices"] == [1] and sc["report_contrast_pairs"] == []
    assert sc["spec"]["fit_columns"] == [0, 2]


def test_report_indices_handle_factor_target_after_dropped_leading_continuous():  # contract_adapter.rs:1258-1352
    c = contract([normal(), factor(3, T3)], [CONST, direct(0), dummy(1, 1), dummy(1, 2)], [0.0, 0.0, 0.3, 0.5],
                 [marginal(1), marginal(2)], design_test=[CONST, dummy(1, 1), dummy(1, 2)])
    sc = lower(c)
    assert sc["spec"]["target_indices"] == [2, 3] and sc["report_target_indices"] == [1, 2]


def test_posthoc_splits_alpha_and_indices():  # contract_adapter.rs:1355-1444
    terms = [CONST, direct(0), dummy(1, 1), dummy(1, 2)]
    c = contract([normal(), factor(3, T3)], terms, [0.0, 0.2, 0.3, 0.5], [marginal(1)], correction="holm",
                 posthoc=[{"factor_column": 1, "target_term_indices": [2, 3], "posthoc_alpha": 0.025}])
    s = lower(c)["spec"]
    assert s["correction_method"] == "holm" and s["crit_values"]["posthoc_alpha"] == 0.025
    assert s["posthoc"] == [{"factor_index": 0, "target_indices": [2, 3]}]


def test_heteroskedasticity_driver_residual_and_scenario_pass_through():  # contract_adapter.rs:1458-1532
    c = example1()
    c["outcome"]["heteroskedasticity_driver"] = 1
    c["outcome"]["residual"] = {"distribution": "high_kurtosis", "pinned": True}
    s = lower(c)["spec"]
    assert s["heteroskedasticity_driver"] == 2 and s["residual_dist"] == "high_kurtosis" and s["residual_pinned"] is True
    lme = contract([normal(), normal()], [CONST, direct(0), direct(1)], [0.0, 0.5, 0.3], [marginal(1)], estimator="mle",
                   cluster={"sizing": {"FixedClusters": {"n_clusters": 20}}, "tau_squared": 0.25},
                   scen=scenario(name="realistic", heterogeneity=0.1, new_distributions=["RightSkewed", {"Binary": {"p": 0.3}}],
                                 residual_dists=["high_kurtosis"],
                                 lme={"random_effect_dist": "normal", "random_effect_df": 5.0, "icc_noise_sd": 0.05}))
    sc = lower(lme)
    s = sc["spec"]
    assert sc["label"] == "realistic" and s["scenario"]["name"] == "realistic"
    assert s["cluster"] == {"sizing": {"FixedClusters": {"n_clusters": 20}}, "tau_squared": 0.25, "slopes": [], "extra_groupings": []}
    assert s["scenario"]["new_distributions"] == ["right_skewed", "binary"] and s["scenario"]["residual_dists"] == ["high_kurtosis"]
    assert s["scenario"]["lme"] == {"random_effect_dist": "normal", "random_effect_df": 5.0, "icc_noise_sd": 0.05}


def test_joint_targets_route_to_report_overall_or_are_rejected():  # contract_adapter.rs:1535-1655
    c = example1()
    c["test"]["targets"] = [{"kind": "joint", "terms": [1, 2]}]
    s = lower(c)["spec"]
    assert s["report_overall"] is True and s["target_indices"] == []
    c["test"]["targets"] = [{"kind": "joint", "terms": [2, 1]}, marginal(1), marginal(2)]
    s = lower(c)["spec"]
    assert s["report_overall"] is True and s["target_indices"] == [1, 2]
    three = contract([normal(), normal(), normal()], [CONST, direct(0), direct(1), direct(2)], [0.0, 0.5, 0.3, 0.1],
                     [{"kind": "joint", "terms": [1, 2]}])
    with pytest.raises(M.InvalidSpec, match=r"omnibus shape.*\[1, 2\]"):
        M.lower_contracts([three])


def test_contrasts_collapse_on_the_reference_level_or_route_to_pairs():  # contract_adapter.rs:1676-1853
    terms = [CONST, dummy(0, 1), dummy(0, 2)]
    ref = contract([factor(3, T3)], terms, [0.0, 0.3, 0.5], [{"kind": "contrast", "positive": 2, "negative": 0}])
    sc = lower(ref)
    assert sc["spec"]["target_indices"] == [2] and sc["spec"]["contrast_pairs"] == [] and sc["report_target_indices"] == [2]
    pair = contract([factor(3, T3)], terms, [0.0, 0.3, 0.5], [{"kind": "contrast", "positive": 1, "negative": 2}])
    sc = lower(pair)
    assert sc["spec"]["target_indices"] == [] and sc["spec"]["contrast_pairs"] == [[1, 2]] and sc["report_contrast_pairs"] == [[1, 2]]


def test_continuous_interaction_maps_to_appended_kernel_column():  # contract_adapter.rs:1879-1962
    c = contract([normal(), normal()], [CONST, direct(0), direct(1), inter(direct(0), direct(1))], [0.0, 0.5, 0.3, 0.2], [marginal(3)])
    s = lower(c)["spec"]
    assert (s["n_non_factor"], s["n_factor_dummies"]) == (2, 0)
    assert s["interactions"] == [[1, 2]] and s["effect_sizes"] == [0.0, 0.5, 0.3, 0.2] and s["target_indices"] == [3]


def _uploaded(data, n_rows, bootstrap):  # contract_adapter.rs:1970-2021
    return contract([{"Resampled": {"frame_column": 0}}], [CONST, direct(0)], [0.0, 0.4], [marginal(1)],
                    frame={"data": list(data), "n_rows": n_rows, "n_cols": 1, "bootstrap": bootstrap})


def test_norta_fills_the_sorted_table_and_strict_the_frame_map():  # contract_adapter.rs:2024-2098
    s = lower(_uploaded([0.3, -1.2, 0.9], 3, False))["spec"]
    assert s["var_types"] == ["uploaded_data"] and s["upload_normal_shape"] == [3, 1] and s["upload_normal"] == [-1.2, 0.3, 0.9]
    assert s["upload_data_shape"] == [3, 1] and s["upload_data"] == [0.3, -1.2, 0.9] and s["bootstrap_frame_map"] == []
    s = lower(_uploaded([0.3, -1.2, 0.9], 3, True))["spec"]
    assert s["upload_normal"] == [] and s["upload_normal_shape"] == [0, 0]
    assert s["upload_data"] == [0.3, -1.2, 0.9] and s["upload_data_shape"] == [3, 1] and s["bootstrap_frame_map"] == [0]


def test_strict_mixed_columns_frame_map_has_none_and_both_passes():  # contract_adapter.rs:2186-2268
    cols = [normal(), {"ResampledBinary": {"frame_column": 0, "proportion": 0.4}},
            {"FactorFromFrame": {"frame_column": 1, "n_levels": 3, "proportions": T3, "sampled_proportions": None}}]
    c = contract(cols, [CONST, direct(0), direct(1), dummy(2, 1), dummy(2, 2)], [0.0, 0.2, 0.3, 0.1, 0.1], [marginal(1)],
                 frame={"data": [0.6, 0.0, -0.4, 1.0, 0.6, 2.0], "n_rows": 3, "n_cols": 2, "bootstrap": True})
    s = lower(c)["spec"]
    assert s["bootstrap_frame_map"] == [None, 0, 1] and s["upload_normal_shape"] == [0, 0]
    assert s["var_types"] == ["normal", "uploaded_binary"] and s["var_params"] == [0.0, 0.4]


def test_cluster_level_columns_and_slopes_resolve_to_kernel_columns():  # contract_adapter.rs:2102-2183, 2272-2351
    cols = [normal(), normal(), factor(3, T3)]
    c = contract(cols, [CONST, direct(0), direct(1), dummy(2, 1), dummy(2, 2)], [0.0, 0.2, 0.3, 0.1, 0.1], [marginal(1)],
                 estimator="mle", cluster_level_columns=[1, 2],
                 cluster={"sizing": {"FixedClusters": {"n_clusters": 20}}, "tau_squared": 0.25,
                          "slopes": [{"column": 0, "variance": 0.1, "corr_with_intercept": 0.0, "corr_with": []},
                                     {"column": 1, "variance": 0.1, "corr_with_intercept": 0.0, "corr_with": [0.0]}]})
    s = lower(c)["spec"]
    assert s["between_var_indices"] == [2, 3, 4] and s["cluster_slope_design_cols"] == [1, 2]


def test_per_factor_sampled_flags_are_lowered():  # contract_adapter.rs:2357-2431
    c = contract([factor(2, [0.5, 0.5], False), factor(2, [0.5, 0.5], True), factor(2, [0.5, 0.5])],
                 [CONST, dummy(0, 1), dummy(1, 1), dummy(2, 1)], [0.0, 0.1, 0.2, 0.3], [marginal(1)])
    assert lower(c)["spec"]["factor_sampled"] == [False, True, None]


def test_correlation_matrix_passes_through():
    c = example1()
    c["generation"]["correlations"] = {"Matrix": {"continuous_columns": [0, 1], "values": [1.0, 0.3, 0.3, 1.0]}}
    assert lower(c)["spec"]["correlation"] == [1.0, 0.3, 0.3, 1.0]


def test_serde_default_fields_may_be_absent():  # contract.rs:71-90 (older producers)
    c = example1()
    del c["wald_se"], c["nagq"], c["posthoc"], c["design_test"]
    del c["generation"]["cluster_level_columns"], c["outcome"]["link"], c["outcome"]["heteroskedasticity_driver"]
    del c["scenario"]["truth_start"], c["scenario"]["sampled_factor_proportions"]
    c["generation"]["columns"][0] = {"Synthetic": {"kind": "Normal"}}
    s = lower(c)["spec"]
    assert s["wald_se"] == "rx" and s["nagq"] == 1 and s["scenario"]["truth_start"] is False


@pytest.mark.parametrize("mutate, message", [  # validate.rs invariants, in the reference's order
    (lambda c: c["outcome"]["coefficients"].pop(), "coefficients length"),                                  # 01 (:1858-1876)
    (lambda c: c["test"]["targets"].append(marginal(7)), "out of range"),                                   # 02
    (lambda c: c["test"]["targets"].append(marginal(1)), "duplicate Marginal"),                             # 03
    (lambda c: c["design_generation"]["terms"].__setitem__(1, direct(5)), "column id 5 out of range"),       # 04
    (lambda c: c["generation"].__setitem__("correlations", {"Matrix": {"continuous_columns": [0, 1], "values": [1.0]}}), "correlation values length"),  # 05
    (lambda c: c["generation"]["columns"].__setitem__(1, factor(3, [0.5, 0.3, 0.3])), "sum to 1.0"),         # 07
    (lambda c: c["generation"]["columns"].__setitem__(1, factor(3, T3)), "Direct term on factor"),          # 09
    (lambda c: c.__setitem__("estimator", "glm"), "binary or count"),                                       # 12
    (lambda c: c.__setitem__("estimator", "mle"), "requires generation.cluster"),                           # 12
    (lambda c: c.__setitem__("max_failed_fraction", 1.5), "max_failed_fraction"),                           # 14
    (lambda c: c["test"].__setitem__("alpha", 1.0), "alpha"),                                               # 15
    (lambda c: c["generation"].__setitem__("correlations", {"Matrix": {"continuous_columns": [0, 1], "values": [1.0, 1.2, 1.2, 1.0]}}), "positive semi-definite"),  # 16
    (lambda c: c["scenario"].__setitem__("distribution_change_prob", 1.5), "distribution_change_prob"),     # 22
    (lambda c: c["generation"]["columns"].__setitem__(0, {"Synthetic": {"kind": {"Binary": {"p": 1.5}}, "pinned": False}}), "binary probability"),  # 23
    (lambda c: c["outcome"].__setitem__("link", "probit"), "probit"),                                       # 24
    (lambda c: c.__setitem__("nagq", 4), "nagq"),                                                           # 25
])
def test_invalid_contracts_are_rejected_with_the_scenario_label(mutate, message):
    c = example1()
    c["scenario"]["name"] = "optimistic"
    mutate(c)
    with pytest.raises(M.InvalidSpec, match=message) as e:
        M.lower_contracts([c])
    assert 'contract 0 ("optimistic")' in str(e.value)  # OrchestratorError::InvalidScenarios (find_power.rs:264-269)


def test_malformed_payloads():  # orchestrator_bridge.rs:23-30
    with pytest.raises(M.InvalidSpec, match="malformed contracts bytes"):
        M.lower_contracts(b"\x91\xc1")
    with pytest.raises(M.InvalidSpec, match="malformed contracts bytes"):
        M.lower_contracts(msgpack.packb([example1()])[:-3])
    with pytest.raises(M.InvalidSpec, match="cannot be empty"):
        M.lower_contracts(msgpack.packb([]))
    with pytest.raises(M.InvalidSpec, match="missing field"):
        M.lower_contracts([{"generation": {}}])


def test_lowered_spec_is_accepted_by_the_spec_reader():
    """The lowered JSON is exactly what mcp_run_batch parses: crit tables can be built from it without a GPU."""
    sc = lower(example1())
    import numpy as np
    t = M.crit_table(sc["spec"], [50, 100], 3)
    assert np.all(np.isfinite(t["t_crit_sq_uncorrected"])) and t["t_crit_sq_uncorrected"][0] > t["t_crit_sq_uncorrected"][1] > 3.8
    assert json.loads(json.dumps(sc["spec"])) == sc["spec"]
    two = copy.deepcopy(example1())
    two["scenario"]["name"] = "b"
    assert [s["label"] for s in M.lower_contracts([example1(), two])] == ["", "b"]


# ------------------------------------------------------------------ find_sample_size guards that fire before any device work
def _clustered(sizing, scen=None):
    return contract([normal(), normal()], [CONST, direct(0), direct(1)], [0.0, 0.5, 0.3], [marginal(1), marginal(2)],
                    estimator="mle", scen=scen,
                    cluster={"sizing": sizing, "tau_squared": 0.25, "slopes": [], "extra_groupings": []})


def test_scenarios_that_disagree_on_the_cluster_atom_are_rejected():  # find_sample_size.rs:33-44, result.rs:606-607
    a = _clustered({"FixedClusters": {"n_clusters": 10}}, scenario(name="a"))
    b = _clustered({"FixedClusters": {"n_clusters": 12}}, scenario(name="b"))
    with pytest.raises(M.InvalidSpec, match=r"disagree on the cluster atom \(10 vs 12\)"):
        M.find_sample_size_contracts([a, b], 40, 400, 64, by=40)
    # an unclustered scenario next to a clustered one shares the clustered grid: the guard looks at clustered scenarios only
    plain = example1()
    with pytest.raises(M.EngineError) as e:  # passes the host-side guards; fails only for lack of a device here
        M.find_sample_size_contracts([a, plain], 40, 400, 64, by=40)
    assert "atom" not in str(e.value)


def test_fixed_size_cluster_below_the_row_minimum_is_rejected():  # find_sample_size.rs:63-70, result.rs:608-609
    with pytest.raises(M.InvalidSpec, match="cluster_size >= 2"):
        M.find_sample_size_contracts([_clustered({"FixedSize": {"cluster_size": 1}})], 40, 400, 64, by=40)


def test_non_finite_numbers_cannot_be_lowered():
    c = example1()
    c["outcome"]["coefficients"][1] = float("inf")
    with pytest.raises(M.InvalidSpec):
        M.lower_contracts(msgpack.packb([c], use_bin_type=True))


def test_lowered_spec_carries_the_configured_factor_minimum():  # find_power.rs:271, configs/config.json:31
    assert lower(example1())["spec"]["factor_min_level_count"] == 5
