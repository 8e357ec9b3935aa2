#!/usr/bin/env python3
"""Extract the reference's frozen END-TO-END crossing goldens (validation/data/cross_*.golden.rds) into a portable fixture.

These are outputs of the shipped Rust engine itself: `find_sample_size(from=20, to=300, by=2, n_sims=100000, seed=2137)`
through the public API for the four CROSS_CASES of validation/formulas.R:660-677 (validation_crossing.rmd:112-135);
`dense$<target>$n_achievable` is the model-based required N (PAVA crossing of the corrected power curve, ceiled to the
cluster atom).  Reproducing them exercises the whole path — RNG, data generation, fit, decisions, aggregation, grid,
crossing fit — against the real reference.  Run HERE (needs /root/reference); output: tests/golden/crossing_goldens.json.
"""
import json
from pathlib import Path

import make_fixtures as mf

REF = Path("/root/reference/validation/data")
CASES = {  # validation/formulas.R:660-677
    "cross_ols": {"formula": "y ~ x1 + x2", "effects": {"x1": 0.40, "x2": 0.25}, "family": "ols"},
    "cross_glm": {"formula": "y ~ x1 + x2", "effects": {"x1": 0.50, "x2": 0.45}, "family": "logit", "baseline_probability": 0.3},
    "cross_lme": {"formula": "y ~ x1 + (1|grp)", "effects": {"x1": 0.45}, "family": "lme",
                  "cluster": {"ICC": 0.2, "n_clusters": 10, "cluster_size": 20}},
    "cross_partial": {"formula": "y ~ x1 + x2", "effects": {"x1": 0.40, "x2": 0.05}, "family": "ols"},
}


def main():
    out = {}
    for name, case in CASES.items():
        g = mf.plain(mf.read_rds(REF / f"{name}.golden.rds"))
        one = lambda v: v[0] if isinstance(v, list) else v
        dense = {t: {"status": one(d["status"]), "n_achievable": (None if d.get("n_achievable") is None else one(d["n_achievable"]))}
                 for t, d in g["dense"].items()}
        out[name] = {"case": case, "seed": int(one(g["seed"])), "n_sims": int(one(g["n_sims"])), "by": int(one(g["by"])),
                     "from_size": int(one(g["from_size"])), "to_size": int(one(g["to_size"])), "frozen_on": one(g["frozen_on"]),
                     "dense": dense}
    path = Path(__file__).resolve().parent / "crossing_goldens.json"
    path.write_text(json.dumps(out, indent=1))
    print(path, json.dumps(out)[:600])


if __name__ == "__main__":
    main()
