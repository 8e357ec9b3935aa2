#!/usr/bin/env python3
"""Scrape the reference's in-tree golden constants into tests/golden/rng_goldens.json.

Run HERE (needs /root/reference).  Sources:
  crates/engine-core/tests/golden_rng.rs:50-147     RNG stream + phi/uniform/sample_t bit goldens
  crates/engine-core/src/data_gen.rs:2282-2400      apply_marginal goldens (z-grid × Distribution)
  crates/engine-core/src/data_gen.rs:2446-2505      draw_residual goldens (df = 8.0, SimRng::new(42, sim_id))
  crates/engine-core/src/posthoc.rs:300-336         12-row R lm contrast fixture (x_cont, y, 3 golden t²)
Only numeric literals are extracted; no reference source text is copied.
"""
import json
import re
from pathlib import Path

CORE = Path("/root/reference/crates/engine-core")
OUT = Path(__file__).resolve().parent / "rng_goldens.json"


def const_array(text, name):
    m = re.search(name + r"\s*:\s*\[[^\]]*\]\s*=\s*\[(.*?)\];", text, re.S)
    body = re.sub(r"//.*", "", m.group(1))
    return [int(v.replace("_", ""), 0) for v in re.findall(r"0x[0-9a-fA-F_]+|\b\d+\b", body)]


def main():
    g = (CORE / "tests/golden_rng.rs").read_text()
    out = {name.lower(): const_array(g, name) for name in (
        "GOLDEN_NORMALS", "GOLDEN_UNIFORMS", "GOLDEN_CATEGORICALS", "GOLDEN_SCENARIO_U32S", "GOLDEN_PHI",
        "GOLDEN_MARGINAL_UNIFORM", "GOLDEN_SAMPLE_T_DF5", "GOLDEN_SAMPLE_T_DF30")}
    out["z_grid"] = [-8.0, -3.0, -1.959963985, -1.0, -0.5, 0.0, 0.5, 1.0, 1.644853627, 3.0, 8.0]

    d = (CORE / "src/data_gen.rs").read_text()
    blk = d[d.index("fn golden_apply_marginal_all_variants"):d.index("fn apply_marginal_column_matches_scalar")]
    marg = []
    for m in re.finditer(r"Distribution::(\w+),\s*([\d.]+),\s*\[(.*?)\]", blk, re.S):
        marg.append({"dist": m.group(1), "param": float(m.group(2)),
                     "bits": [int(v, 16) for v in re.findall(r"0x[0-9a-f]+", m.group(3))]})
    out["apply_marginal"] = {"z_grid": [-8.0, -3.0, -1.0, 0.0, 1.0, 3.0, 8.0], "rows": marg}
    blk = d[d.index("fn golden_draw_residual_all_variants"):d.index("fn cluster_draws_precede_row_loop")]
    res = []
    for m in re.finditer(r"ResidualDist::(\w+),\s*(\d+),(.*?)\],", blk, re.S):
        res.append({"dist": m.group(1), "sim_id": int(m.group(2)),
                    "bits": [int(v, 16) for v in re.findall(r"0x[0-9a-f]+", m.group(3))]})
    out["draw_residual"] = {"base_seed": 42, "df": 8.0, "rows": [r for r in res if len(r["bits"]) == 8]}

    p = (CORE / "src/posthoc.rs").read_text()
    blk = p[p.index("fn posthoc_contrast_golden_three_level_factor"):]
    def floats(name):
        m = re.search(r"let " + name + r"\s*=\s*\[(.*?)\];", blk, re.S)
        return [float(v.replace("_", "")) for v in re.findall(r"-?[\d_]+\.[\d_]+", m.group(1))]
    out["posthoc_lm"] = {
        "x_cont": floats("x_cont"), "y": floats("y"),
        "tsq_b": float(re.search(r"tsq_b: f64 = ([\d.]+)", blk).group(1)),
        "tsq_c": float(re.search(r"tsq_c: f64 = ([\d.]+)", blk).group(1)),
        "tsq_b_minus_c": float(re.search(r"tsq_b_minus_c: f64 = ([\d.]+)", blk).group(1)),
    }
    OUT.write_text(json.dumps(out, indent=0))
    print({k: (len(v) if isinstance(v, list) else "…") for k, v in out.items()})
    print(len(out["apply_marginal"]["rows"]), len(out["draw_residual"]["rows"]), out["posthoc_lm"]["tsq_b"])


if __name__ == "__main__":
    main()
