"""Parity at BASELINE.json's full sizes (configs[1]: 1e6 sims per cell over the 50-point sweep; the headline shape at 4 Mi sims).
The oracle cannot run a million studies in test time, so full-size runs are checked (a) against the oracle on WINDOWS of sim ids
cut out of the full-size byte tensors — a study's bits depend on its global sim id only, so any window of the big run must equal
the oracle's run of exactly those ids — and (b) through size-independent properties: COUNTS == aggregate(BYTES) over the whole
1.1 GB, shard concatenation, monotone power."""
import numpy as np
import pytest

import mcpower_b200 as M
from mcpower_b200 import spec as S
from oracle import orc

pytestmark = pytest.mark.gpu
SEED = M.splitmix64(2137)


def test_config2_at_one_million_sims_per_cell():
    spec = S.sweep_spec()
    grid = list(range(20, 1001, 20))
    n = 1_000_000
    got = M.run_batch(spec, grid, n, SEED, mode=M.MODE_COUNTS | M.MODE_BYTES)
    assert got.uncorrected.shape == (n, 50, 10) and got.n_waves > 1
    # (a) windows of the full-size tensors against the oracle on the same global sim ids (first, wave-interior, last)
    for begin in (0, 137_211, 654_321, n - 48):
        want = orc.run_batch(spec, grid, 48, SEED, sim_begin=begin, n_threads=orc.host_threads())
        for key in ("uncorrected", "corrected", "converged", "overall", "factor_excluded"):
            assert np.array_equal(getattr(got, key)[begin:begin + 48], want[key]), (begin, key)
    # (b) the counters are the fold of the byte tensors (aggregate_batch, aggregation.rs:91-242) — all 5e7 fits
    T = got.n_targets
    assert np.array_equal(got.uncorrected.sum(axis=0, dtype=np.uint64), got.counters[:, :T])
    assert np.array_equal(got.corrected.sum(axis=0, dtype=np.uint64), got.counters[:, T:2 * T])
    assert np.array_equal(got.converged.sum(axis=0, dtype=np.uint64), got.counters[:, 2 * T])
    assert np.array_equal(got.overall.sum(axis=0, dtype=np.uint64), got.counters[:, 2 * T + 1])
    assert np.array_equal((got.factor_excluded[:, :, 0] == 1).sum(axis=0, dtype=np.uint64), got.n_factor_excluded[:, 0])
    hist = np.stack([np.bincount(got.uncorrected[:, g].sum(axis=1), minlength=T + 1) for g in range(50)])
    assert np.array_equal(hist.astype(np.uint64), got.hist_uncorrected)
    # the COUNTS-only run of the same call is the same integers
    counts_only = M.run_batch(spec, grid, n, SEED)
    assert np.array_equal(counts_only.counters, got.counters)
    # power rises with N for a real effect; corrected <= uncorrected everywhere
    p = got.power()[:, 0]
    assert p[-1] > 0.99 and np.all(np.diff(p) > -0.004)
    assert np.all(got.success_corrected <= got.success_uncorrected)


def test_headline_shape_at_four_million_sims_matches_the_oracle_on_windows():
    spec = S.headline_spec()
    n = 1 << 22
    got = M.run_batch(spec, [500], n, SEED, mode=M.MODE_COUNTS | M.MODE_BYTES)
    for begin in (0, 1_234_567, n - 64):
        want = orc.run_batch(spec, [500], 64, SEED, sim_begin=begin, n_threads=orc.host_threads())
        assert np.array_equal(got.uncorrected[begin:begin + 64], want["uncorrected"]), begin
        assert np.array_equal(got.converged[begin:begin + 64], want["converged"]), begin
    a = M.run_batch(spec, [500], n // 2, SEED)
    b = M.run_batch(spec, [500], n // 2, SEED, sim_begin=n // 2)
    assert np.array_equal(a.counters + b.counters, got.counters)
    # analytic non-central-t power of one slope (beta = 0.18, n = 500, P = 9) within 4 binomial standard errors
    import _cases as CS
    p_hat = got.counters[0, 0] / n
    p_ref = CS.analytic_ols_power(0.18, 500, 9)
    assert abs(p_hat - p_ref) < 4 * np.sqrt(p_ref * (1 - p_ref) / n) + 2e-4
