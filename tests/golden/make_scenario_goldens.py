#!/usr/bin/env python3
"""Extract the reference's frozen scenario-perturbation goldens (validation/data/scenario_golden.rds) into a fixture.

Each value is a statistic the SHIPPED engine's generated data produced in validation_scenarios.rmd (K draws, draw k = sim 0
of base_seed = splitmix64(seed + k) under a one-knob custom scenario; case list validation/formulas.R:705-830, probes
validation/common.R:895-1075), frozen with `golden_abs = 1e-9` (tolerances.R:146).  Run HERE (needs /root/reference).
"""
import json
from pathlib import Path

import make_fixtures as mf

CASES = {  # validation/formulas.R:708-806
    "scen_he": {"effects": [0.40, 0.25], "knobs": {"heterogeneity": 0.4}, "n": 4000, "K": 200, "seed": 31000},
    "scen_hs": {"effects": [0.25], "knobs": {"heteroskedasticity_ratio": 4.0}, "n": 4000, "K": 200, "seed": 32000},
    "scen_co_low": {"effects": [0.25, 0.25], "corr": 0.3, "knobs": {"correlation_noise_sd": 0.15}, "n": 4000, "K": 400, "seed": 33000},
    "scen_co_high": {"effects": [0.25, 0.25], "corr": 0.75, "knobs": {"correlation_noise_sd": 0.15}, "n": 2000, "K": 800, "seed": 34000},
    "scen_co_psd": {"effects": [0.25, 0.25, 0.25], "corr": 0.6, "knobs": {"correlation_noise_sd": 0.3}, "n": 2000, "K": 400, "seed": 35000},
    "scen_px": {"effects": [0.25, 0.25], "knobs": {"distribution_change_prob": 0.5}, "n": 4000, "K": 400, "seed": 36000},
    "scen_px_t3": {"effects": [0.25], "knobs": {"distribution_change_prob": 1.0, "new_distributions": ["high_kurtosis"]},
                   "n": 4000, "K": 200, "seed": 37000},
    "scen_re": {"effects": [0.25], "knobs": {"residual_change_prob": 0.5, "residual_df": 10.0}, "n": 4000, "K": 400, "seed": 38000},
    "scen_re_replace": {"effects": [0.25], "residual": "high_kurtosis",
                        "knobs": {"residual_change_prob": 1.0, "residual_dists": ["right_skewed"], "residual_df": 6.0},
                        "n": 4000, "K": 200, "seed": 39000},
    # g = (factor, 0.5, 0.3, 0.2); fixed walk under the optimistic scenario (20 draws, seed + k), sampled levels under
    # sampled_factor_proportions (K draws, seed + 500 + k): run_fa, validation_scenarios.rmd:262-283
    # GLM families under heterogeneity: mean fitted slope over K studies (run_he_glm, validation_scenarios.rmd:64-85; the
    # fit there is R glm — the tests use the oracle's IRLS from the truth start, see TOL_GLM in test_scenario_goldens.py)
    "scen_he_probit": {"effects": [0.5, 0.3], "family": "probit", "baseline_probability": 0.4, "knobs": {"heterogeneity": 0.4},
                       "n": 2000, "K": 300, "seed": 60000},
    "scen_he_poisson": {"effects": [0.5, 0.3], "family": "poisson", "baseline_rate": 2.0, "knobs": {"heterogeneity": 0.4},
                        "n": 2000, "K": 300, "seed": 61000},
    "scen_fa_ols": {"effects": [0.25], "factor": {"proportions": [0.5, 0.3, 0.2], "effects": [0.50, 0.80]},
                    "knobs": {"sampled_factor_proportions": True}, "n": 1000, "K": 400, "seed": 40000},
}


def main():
    g = mf.plain(mf.read_rds(Path("/root/reference/validation/data/scenario_golden.rds")))
    out = {}
    for name, case in CASES.items():
        vals = {k: (v[0] if isinstance(v, list) else v) for k, v in g[name].items()}
        out[name] = {"case": case, "golden": vals}
    path = Path(__file__).resolve().parent / "scenario_goldens.json"
    path.write_text(json.dumps(out, indent=1, ensure_ascii=False))
    print(path, len(out))


if __name__ == "__main__":
    main()
