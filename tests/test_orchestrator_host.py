"""Host-side orchestrator logic against the reference's own known-answer tests: the crossing fit
(engine-orchestrator/src/fit.rs:262-491), the grid builder (grid.rs:393-638, tests/test_grid.rs) and the Wilson interval
(aggregation.rs:18-33, tests/test_aggregation.rs).  Pure host functions of the C ABI: no GPU needed."""
import math

import pytest

import mcpower_b200 as M

GRID = [50, 100, 150, 200]


def test_fitted_crossing_with_two_sided_ci():  # fit.rs:318-335
    f = M.fit_crossing(GRID, [60, 75, 85, 95], 100, 0.8)
    assert f["status"] == "fitted" and f["n_achievable"] == 125
    assert abs(f["n_star"] - 125.0) < 1e-9
    assert abs(f["ci_lo"] - 90.83640209948472) < 1e-9 and abs(f["ci_hi"] - 163.5595794316606) < 1e-9


def test_exact_hit_resolves_leftmost():  # fit.rs:337-355
    f = M.fit_crossing(GRID, [60, 70, 80, 84], 100, 0.8)
    assert f["status"] == "fitted" and f["n_achievable"] == 150 and abs(f["n_star"] - 150.0) < 1e-9
    assert abs(f["ci_lo"] - 111.06986495758443) < 1e-9 and f["ci_hi"] is None


def test_ci_lo_none_when_hi_band_at_floor():  # fit.rs:357-372
    f = M.fit_crossing(GRID, [75, 85, 95, 98], 100, 0.8)
    assert abs(f["n_star"] - 75.0) < 1e-9 and f["ci_lo"] is None and abs(f["ci_hi"] - 113.55957943166061) < 1e-9


def test_at_or_below_min_and_atom_and_percentage():  # fit.rs:374-396
    assert M.fit_crossing(GRID, [85, 90, 95, 98], 100, 0.8) == {"status": "at_or_below_min", "n_min": 50}
    assert M.fit_crossing(GRID, [60, 75, 85, 95], 100, 0.8, atom=30)["n_achievable"] == 150
    assert M.fit_crossing(GRID, [60, 75, 85, 95], 100, 80.0) == M.fit_crossing(GRID, [60, 75, 85, 95], 100, 0.8)


def test_monotonicity_gate():  # fit.rs:399-427
    assert M.fit_crossing([50, 100], [50, 40], 100, 0.9) == {"status": "not_reached", "n_approx": None}
    f = M.fit_crossing([50, 100], [50, 35], 100, 0.9)
    assert f["status"] == "non_monotone" and abs(f["max_violation"] - 0.15) < 1e-12
    f = M.fit_crossing(GRID, [50, 20, 30, 95], 100, 0.8)
    assert f["status"] == "non_monotone" and abs(f["max_violation"] - 0.3) < 1e-12


def test_probit_hint_goldens():  # fit.rs:430-481
    assert M.fit_crossing(GRID, [30, 45, 60, 72], 100, 0.9) == {"status": "not_reached", "n_approx": 352}
    assert M.fit_crossing(GRID, [30, 45, 60, 72], 100, 0.9, atom=30) == {"status": "not_reached", "n_approx": 360}
    assert M.fit_crossing(GRID, [30, 40, 55, 65], 100, 0.9) == {"status": "not_reached", "n_approx": None}
    assert M.fit_crossing(GRID, [20, 30, 40, 49], 100, 0.6) == {"status": "not_reached", "n_approx": None}
    assert M.fit_crossing(GRID, [20, 30, 40, 51], 100, 0.6) == {"status": "not_reached", "n_approx": 263}
    assert M.fit_crossing(GRID, [60, 60, 60, 60], 100, 0.7) == {"status": "not_reached", "n_approx": None}
    assert M.fit_crossing(GRID, [0, 0, 0, 0], 0, 0.8) == {"status": "not_reached", "n_approx": None}


def test_pava_through_the_fit():  # fit.rs:275-308: pooled blocks change where the fit crosses
    # p = [0.1, 0.5, 0.3, 0.4, 0.9] pools to [0.1, 0.4, 0.4, 0.4, 0.9] (gate not fired at n_sims = 20): crossing of 0.65
    # lies on the last segment, (0.65 - 0.4) / 0.5 of the way from 40 to 50
    f = M.fit_crossing([10, 20, 30, 40, 50], [2, 10, 6, 8, 18], 20, 0.65)
    assert f["status"] == "fitted" and abs(f["n_star"] - 45.0) < 1e-9


def test_build_grid_kats():  # grid.rs:398-470, tests/test_grid.rs
    g, w = M.build_grid(30, 200, count=12)
    assert g[0] == 30 and g[-1] == 200 and 11 <= len(g) <= 13 and w == 0 and all(a < b for a, b in zip(g, g[1:]))
    g, _ = M.build_grid(30, 52, by=5)
    assert g[0] == 30 and g[-1] == 52
    assert M.build_grid(30, 50, by=5)[0] == [30, 35, 40, 45, 50]
    g, w = M.build_grid(30, 200, count=12, atom=20, hard_min=40)
    assert all(n % 20 == 0 for n in g) and g[0] == 40 and g[-1] == 200 and w >= 1
    g, _ = M.build_grid(20, 300, by=2, atom=10, hard_min=20)
    assert g == list(range(20, 301, 10))  # validation_crossing: "29 points" for cross_lme
    assert len(M.build_grid(20, 300, by=2)[0]) == 141
    with pytest.raises(M.InvalidSpec):
        M.build_grid(50, 40, by=5)
    g, _ = M.build_grid(10, 1000, by=5, log_grid=True)
    assert g[0] == 10 and g[-1] == 1000 and len(g) == 5 and g == sorted(set(g))


def test_wilson_ci():  # aggregation.rs:18-33, tests/test_aggregation.rs
    assert M.wilson_ci(0, 0) == (0.0, 1.0)
    for n in (1, 7, 100, 1600):
        prev = (-1.0, -1.0)
        for k in range(n + 1):
            lo, hi = M.wilson_ci(k, n)
            assert 0.0 <= lo <= k / n + 1e-12 and k / n - 1e-12 <= hi <= 1.0 and lo >= prev[0] and hi >= prev[1]
            prev = (lo, hi)
    lo, hi = M.wilson_ci(75, 100)
    z, p = 1.96, 0.75
    c = (p + z * z / 200) / (1 + z * z / 100)
    m = z * math.sqrt(p * (1 - p) / 100 + z * z / 40000) / (1 + z * z / 100)
    assert abs(lo - (c - m)) < 1e-15 and abs(hi - (c + m)) < 1e-15
