"""The host path around the kernels (reference: run_batch's ProgressSink handling, batch.rs:449-450, 2149-2177; the rayon
fan-out over sims, batch.rs:637-742; result ownership, spec.rs:552-600): progress without cutting the batch into small
launches, cancellation, the pipelined BYTES read-back, and sharding one call over several devices.  Every variation must leave
every output bit unchanged."""
import numpy as np
import pytest

import mcpower_b200 as M
from mcpower_b200 import spec as S

pytestmark = pytest.mark.gpu

SEED = 0x0DDC0FFEE0DDF00D
ALL = M.MODE_COUNTS | M.MODE_BYTES | M.MODE_CAPTURE


def _same(a: M.BatchResult, b: M.BatchResult):
    assert np.array_equal(a.counters, b.counters)
    for key in ("uncorrected", "corrected", "converged", "overall", "factor_excluded", "joint", "boundary_hit", "posthoc_unc", "posthoc_cor"):
        x, y = getattr(a, key), getattr(b, key)
        assert (x is None) == (y is None), key
        if x is not None:
            assert np.array_equal(x, y), key
    if a.t_sq is not None:
        assert np.array_equal(a.t_sq.view(np.uint64), b.t_sq.view(np.uint64))


def test_progress_is_reported_without_cutting_the_batch():
    spec, n = S.headline_spec(), 200_000
    plain = M.run_batch(spec, [500], n, SEED)
    seen = []
    got = M.run_batch(spec, [500], n, SEED, sink=lambda cur, tot: seen.append((cur, tot)) or True)
    assert np.array_equal(got.counters, plain.counters)
    # one wave, two launches (generate + accumulate, solve): the sink costs no extra launch and no host synchronisation
    assert got.n_waves == 1 and got.gpu_launches == 2 == plain.gpu_launches
    assert seen and seen[-1] == (n, n)
    cur = [c for c, _ in seen]
    assert cur == sorted(cur) and len(set(cur)) == len(cur), "reports must be strictly increasing"
    assert all(t == n for _, t in seen)
    assert len(seen) <= 51  # ~50 per run (batch.rs:449-450)
    step = n // 50
    assert all(c % step == 0 for c in cur[:-1])


def test_default_size_call_with_sink_is_two_launches():
    # BASELINE config 1 through the default host path: 1600 sims, progress sink attached (engine-py always passes one)
    spec = S.add_interaction(S.base_spec(2), [1, 2], 0.2)
    seen = []
    got = M.run_batch(spec, [100], 1600, SEED, sink=lambda c, t: seen.append(c) or True)
    assert got.gpu_launches == 2 and got.n_waves == 1
    assert seen[-1] == 1600


@pytest.mark.parametrize("arm", ["ols", "glm", "lmm"])
def test_cancel_from_the_sink_discards_the_batch(arm):
    spec, grid, n = {"ols": (S.headline_spec(), [500], 1 << 20), "glm": (S.logit_spec(), [200, 400], 1 << 17),
                     "lmm": (S.lmm_spec(30, 0.2), [300, 600], 1 << 18)}[arm]
    calls = []

    def sink(cur, tot):
        calls.append(cur)
        return len(calls) < 3

    with pytest.raises(M.Cancelled):
        M.run_batch(spec, grid, n, SEED, sink=sink)
    assert len(calls) == 3 and calls[-1] < n
    # the device is left clean: the next call runs and gives the uncancelled answer
    a = M.run_batch(spec, grid, 4096, SEED)
    b = M.run_batch(spec, grid, 4096, SEED, sink=lambda c, t: True)
    assert np.array_equal(a.counters, b.counters)


def test_is_cancelled_poll_stops_the_batch():
    polls = []

    def is_cancelled():
        polls.append(1)
        return len(polls) >= 5

    with pytest.raises(M.Cancelled):
        M.run_batch(S.headline_spec(), [500], 1 << 21, SEED, is_cancelled=is_cancelled)
    assert len(polls) == 5


def test_cancel_reaches_a_running_mixed_model_search():
    """The fits with extra groupings run for milliseconds per thread: the simplex search polls the cancel flag, so a cancelled call
    returns while the launch is still in flight instead of after its slowest fit."""
    import time
    s = S.lmm_spec(12, 0.2)
    s["cluster"]["extra_groupings"] = [{"relation": {"Crossed": {"n_clusters": 6}}, "tau_squared": 0.15, "slopes": []}]
    t0 = time.perf_counter()
    full = M.run_batch(s, [144, 288], 8192, SEED)
    t_full = time.perf_counter() - t0
    polls = []

    def is_cancelled():
        polls.append(1)
        return len(polls) >= 20

    t0 = time.perf_counter()
    with pytest.raises(M.Cancelled):
        M.run_batch(s, [144, 288], 8192, SEED, is_cancelled=is_cancelled)
    t_cancel = time.perf_counter() - t0
    assert t_cancel < 0.8 * t_full + 0.05, (t_cancel, t_full)
    again = M.run_batch(s, [144, 288], 8192, SEED)
    assert np.array_equal(again.counters, full.counters)


def test_bytes_readback_is_pipelined_and_bit_identical():
    spec = S.sweep_spec()
    grid = [40, 120, 400]
    n = 20_000
    one = M.run_batch(spec, grid, n, SEED, mode=ALL)
    assert one.n_waves == 1
    waves = M.run_batch(spec, grid, n, SEED, mode=ALL, wave_sims=4096)
    assert waves.n_waves == 5
    _same(one, waves)
    agg = M.counters_from_bytes(waves.uncorrected, waves.corrected, waves.converged, waves.overall, waves.factor_excluded)
    assert np.array_equal(agg, waves.counters)
    # the byte accounting is the tensors' size, not a constant
    per_cell = 2 * one.n_targets + 1 + one.n_factors
    assert one.d2h_bytes >= n * len(grid) * (per_cell + 8 * one.n_targets)
    assert one.h2d_bytes > 0


def test_large_bytes_call_splits_into_pipelined_waves():
    spec = S.headline_spec()
    n = 1 << 26  # 64 Mi sims x (2 + 2 + 1) bytes = 320 MiB of byte tensors: two pieces of >= 128 MiB
    got = M.run_batch(spec, [60], n, SEED, mode=M.MODE_COUNTS | M.MODE_BYTES)
    assert got.n_waves >= 2  # (the statistics arena caps a wave at 4 GiB, so this one is cut further)
    agg = np.zeros_like(got.counters)
    agg[0, :2] = got.uncorrected.sum(axis=0, dtype=np.uint64)[0]
    assert np.array_equal(agg[0, :2], got.counters[0, :2])


@pytest.mark.parametrize("arm", ["ols_lean", "ols_split", "glm", "lmm"])
def test_one_call_over_several_devices_equals_one_device(arm):
    if M.device_count() < 2:
        pytest.skip("needs two GPUs")
    spec, grid, n = {"ols_lean": (S.sweep_spec(), [40, 200, 600], 30_001),
                     "ols_split": (dict(S.sweep_spec(), scenario=S.scenario("doomer")), [40, 200], 9_001),
                     "glm": (S.logit_spec(), [100, 300], 5_001),
                     "lmm": (S.lmm_spec(30, 0.2), [300, 600], 7_001)}[arm]
    one = M.run_batch(spec, grid, n, SEED, mode=ALL, device=0)
    nd = min(M.device_count(), 8)
    seen = []
    many = M.run_batch(spec, grid, n, SEED, mode=ALL, device=0, n_devices=nd, sink=lambda c, t: seen.append(c) or True)
    assert many.n_devices == nd
    _same(one, many)
    assert seen[-1] == n
    allv = M.run_batch(spec, grid, n, SEED, device=0, n_devices=-1)
    assert np.array_equal(allv.counters, one.counters)


def test_more_devices_than_sims():
    if M.device_count() < 2:
        pytest.skip("needs two GPUs")
    one = M.run_batch(S.headline_spec(), [60], 1, SEED, mode=ALL, device=0)
    many = M.run_batch(S.headline_spec(), [60], 1, SEED, mode=ALL, device=0, n_devices=-1)
    _same(one, many)
    assert many.n_devices == 1  # empty shards are dropped


def test_device_out_of_range_is_a_cuda_error():
    with pytest.raises(M.CudaError):
        M.run_batch(S.headline_spec(), [60], 8, SEED, device=M.device_count())
    with pytest.raises(M.CudaError):
        M.run_batch(S.headline_spec(), [60], 8, SEED, device=0, n_devices=M.device_count() + 1)


def test_zero_sims_returns_an_empty_shaped_result():  # batch.rs:390-421
    got = M.run_batch(S.sweep_spec(), [40, 80], 0, SEED, mode=ALL)
    assert got.n_sims == 0 and got.counters.shape[0] == 2 and not got.counters.any()
    assert got.uncorrected.shape == (0, 2, got.n_targets)


def test_concurrent_calls_from_threads_are_serialised_per_device():
    import threading
    spec = S.sweep_spec()
    want = M.run_batch(spec, [60, 200], 8192, SEED).counters
    out, errs = [None] * 4, []

    def work(i):
        try:
            out[i] = M.run_batch(spec, [60, 200], 8192, SEED, sink=lambda c, t: True).counters
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs
    assert all(np.array_equal(o, want) for o in out)


def test_contract_runners_report_one_progress_range_and_preflight_warnings():
    """find_power over two scenarios: the sink sees offset + done of n_sims x scenarios (find_power.rs:311-337); a sparse level of a
    fixed-allocation factor raises the pre-run advisory (find_power.rs:67-116, 364-368); find_sample_size shares one grid and names
    the N at which the factor clears the minimum (find_power.rs:122-170)."""
    import ctypes as C
    import json
    import msgpack
    import test_contract_lowering as TC
    from mcpower_b200 import engine
    terms = [TC.CONST, TC.direct(0), TC.dummy(1, 1)]
    base = TC.contract([TC.normal(), TC.factor(2, [0.9, 0.1])], terms, [0.0, 0.4, 0.5], [TC.marginal(1), TC.marginal(2)])
    base["scenario"]["name"] = "optimistic"
    second = __import__("copy").deepcopy(base)
    second["scenario"] = TC.scenario(name="noisy", heterogeneity=0.3)
    blob = msgpack.packb([base, second], use_bin_type=True)
    seen = []
    cb = engine._PROGRESS(lambda cur, tot, _u: seen.append((int(cur), int(tot))) or 1)
    out = C.c_void_p()
    opt = engine._options(0, M.MODE_COUNTS)
    n_sims = 50_000
    rc = engine.lib().mcp_find_power_contracts(blob, len(blob), 0, 40, n_sims, 2137, C.byref(opt), cb, None, C.byref(out))
    assert rc == 0, engine.lib().mcp_last_error()
    res = engine._take_json(out)
    assert [t for _, t in seen] == [2 * n_sims] * len(seen)
    cur = [c for c, _ in seen]
    assert cur == sorted(cur) and cur[-1] == 2 * n_sims and n_sims in cur
    for _, pr in res["scenarios"]:
        w = [m for m in pr["grid_warnings"] if m.startswith("factor 1: level 2 receives 4 of 40 observations (minimum 5)")]
        assert len(w) == 1, pr["grid_warnings"]
        assert pr["factor_exclusion_counts"] == [n_sims]
    ss = M.find_sample_size_contracts(blob, 20, 100, 2048, 2137, by=20)
    for _, r in ss["scenarios"]:
        assert [p["n"] for p in r["grid_or_trace"]] == [20, 40, 60, 80, 100]
        assert any("factor 1: levels reach the 5-observation minimum only at N >= 46" in m for m in r["grid_warnings"]), r["grid_warnings"]
