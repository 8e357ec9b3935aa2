"""Scenario-perturbation goldens of the SHIPPED engine (validation/data/scenario_golden.rds, frozen with golden_abs = 1e-9,
validation/tolerances.R:146): heterogeneity jitter, heteroskedasticity, correlation noise with and without the PSD repair,
marginal swaps (preset pool and the t(3) table), residual swaps (t / chi-square) and a pinned residual.  The statistics are
recomputed from the generated designs/outcomes of K draws (draw k = sim 0 of base_seed = splitmix64(seed + k)) — by the
oracle on CPU and by the CUDA data plane (mcp_generate_sim) on the GPU."""
import numpy as np
import pytest

import _scenario_probes as SP

TOL = 1e-9  # SCENARIO_TOL$golden_abs
# The two GLM cases freeze the mean of R `glm` slopes; the tests refit with the oracle's IRLS (R's stopping rule, truth
# start), so the comparison carries R's own convergence tolerance (|ddev|/(|dev|+0.1) < 1e-8 ⇒ ~1e-7 on a probit slope).
# Measured: Poisson 5e-11, probit 1.4e-7.
TOL_GLM = {"scen_he_probit": 1e-6, "scen_he_poisson": 1e-8}


def _with_sampling(spec, sampled):
    sp = dict(spec)
    sp["scenario"] = dict(spec["scenario"], sampled_factor_proportions=bool(sampled))
    return sp


def _check(name, draw):
    got = SP.statistics(name, draw)
    for key, want in SP.GOLDEN[name]["golden"].items():
        assert abs(got[key] - want) <= TOL_GLM.get(name, TOL), (name, key, got[key], want)


@pytest.mark.parametrize("name", sorted(SP.GOLDEN))
def test_oracle_reproduces_engine_scenario_goldens(name):
    from oracle import orc
    case = SP.GOLDEN[name]["case"]
    spec = SP.spec_of(case)

    def draw(seed, sampled=None):
        sp = spec if sampled is None else _with_sampling(spec, sampled)
        x, y, _ = orc.generate_sim(sp, orc.splitmix64(seed), 0, case["n"])
        return x.astype(np.float64), y.astype(np.float64)

    _check(name, draw)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(SP.GOLDEN))
def test_cuda_data_plane_reproduces_engine_scenario_goldens(name):
    import mcpower_b200 as M
    case = SP.GOLDEN[name]["case"]
    spec = SP.spec_of(case)

    def draw(seed, sampled=None):
        sp = spec if sampled is None else _with_sampling(spec, sampled)
        x, y = M.generate_sim(sp, M.splitmix64(seed), 0, case["n"])
        return x.astype(np.float64), y.astype(np.float64)

    _check(name, draw)
