"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol include/mcpower_b200.h
declares; host logic (JSON spec parsing, validation errors, critical values) behaves like the reference."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

import mcpower_b200 as M
from mcpower_b200 import engine, spec as S
from oracle import orc

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "mcpower_b200.h").read_text()
    declared = set(re.findall(r"\b(mcp_[a-z0-9_]+)\s*\(", header))
    assert declared == set(engine.EXPORTS), declared ^ set(engine.EXPORTS)
    lib = ctypes.CDLL(str(engine._SO))
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert b"sm_100a" in engine.lib().mcp_version()


def test_splitmix64_matches_reference_kat():  # tests/rng_byte_identity.rs:36-42
    assert M.splitmix64(1) == 0x5692161D100B05E5 and M.splitmix64(0) == 0


@pytest.mark.parametrize("method", ["none", "bonferroni", "holm", "benjamini_hochberg", "tukey_hsd"])
def test_product_crit_table_bit_identical_to_oracle(method):
    s = S.sweep_spec()
    s["correction_method"] = method
    if method == "tukey_hsd":
        s["target_indices"] = [9, 10]
        grid = [30, 200]
    else:
        grid = [12, 20, 57, 1000, 250000]
    a, b = M.crit_table(s, grid), orc.crit_table(s, grid)
    for k in a:
        assert np.array_equal(a[k], b[k], equal_nan=True), (method, k)
    a, b = M.crit_table(s, [40], n_predictors=9), orc.crit_table(s, [40], n_predictors=9)
    assert np.array_equal(a["correction_t_crit_sq"], b["correction_t_crit_sq"], equal_nan=True)


def test_crit_quantile_goldens():  # critvals.rs:1836-1852, batch.rs:5035-5055
    s = S.base_spec(1)
    s["target_indices"] = [1]
    ct = M.crit_table(s, [32], n_predictors=2)   # df = 30
    assert abs(np.sqrt(ct["t_crit_sq_uncorrected"][0]) / 2.042272 - 1) < 1e-5
    ct = M.crit_table(s, [12], n_predictors=2)   # df = 10: t² = F(1, 10)
    assert abs(ct["t_crit_sq_uncorrected"][0] - 4.9646027) < 1e-5
    assert abs(ct["overall_crit"][0] - ct["t_crit_sq_uncorrected"][0]) < 1e-7


def test_validation_errors_without_a_gpu():  # batch.rs:208-330 run before any device work
    s = S.headline_spec()
    with pytest.raises(M.InvalidSpec, match="strictly increasing"):
        M.run_batch(s, [50, 50], 4, 1)
    with pytest.raises(M.InvalidSpec, match="non-empty"):
        M.run_batch(s, [], 4, 1)
    with pytest.raises(M.InvalidSpec, match="must be > 0"):
        M.run_batch(s, [0, 5], 4, 1)
    bad = dict(s, effect_sizes=[0.0, 0.1])
    with pytest.raises(M.InvalidSpec, match="effect_sizes length"):
        M.run_batch(bad, [50], 4, 1)
    bad = dict(s, target_indices=[1, 99])
    with pytest.raises(M.InvalidSpec, match="out of range"):
        M.run_batch(bad, [50], 4, 1)
    bad = dict(s, correction_method="sidak")
    with pytest.raises(M.InvalidSpec, match="unknown variant"):
        M.run_batch(bad, [50], 4, 1)
    bad = S.base_spec(2)
    bad["correlation"] = [1.0, 1.5, 1.5, 1.0]
    with pytest.raises(M.CorrelationNotPSD):
        M.run_batch(bad, [50], 4, 1)
    mle = dict(s, estimator="mle")
    with pytest.raises(M.InvalidSpec, match="requires ClusterSpec"):
        M.run_batch(mle, [50], 4, 1)


def test_uploaded_table_validation_without_a_gpu():
    """The uploaded-data tables are checked with the rest of the spec, before any device work (the reference indexes them
    unchecked and would panic; data_gen.rs:117-121, 563)."""
    frame = [[0.1 * i - 1.0, float(i % 3)] for i in range(20)]
    strict = S.with_uploaded_frame(S.add_factor(S.base_spec(1), [0.5, 0.3, 0.2], [0.1, 0.2]), frame, {0: 0}, factors={0: 1}, bootstrap=True)
    bad = dict(strict, bootstrap_frame_map=[0, 7])
    with pytest.raises(M.InvalidSpec, match="bootstrap_frame_map entry 7 out of range"):
        M.run_batch(bad, [50], 4, 1)
    bad = dict(strict, bootstrap_frame_map=[0])
    with pytest.raises(M.InvalidSpec, match="n_non_factor \\+ n_factors"):
        M.run_batch(bad, [50], 4, 1)
    bad = dict(strict, upload_data_shape=[19, 2])
    with pytest.raises(M.InvalidSpec, match="does not match `upload_data_shape`"):
        M.run_batch(bad, [50], 4, 1)
    norta = S.with_uploaded_frame(S.base_spec(2), frame, {1: 0})
    assert norta["upload_normal_shape"] == [20, 2] and norta["upload_normal"][1::2] == sorted(r[0] for r in frame)
    bad = dict(norta, upload_normal=norta["upload_normal"][:-2], upload_normal_shape=[19, 2])
    bad["upload_normal_shape"] = [19, 1]
    with pytest.raises(M.InvalidSpec):
        M.run_batch(bad, [50], 4, 1)
    menu = dict(S.base_spec(2), scenario=dict(S.scenario("realistic"), new_distributions=["uploaded_data"]))
    with pytest.raises(M.Unsupported, match="scenario menu"):
        M.run_batch(menu, [50], 4, 1)


def test_contract_runners_report_lowering_errors_without_a_gpu():
    """mcp_find_power_contracts / mcp_find_sample_size_contracts lower every contract first (find_power.rs:251-288)."""
    bad = {"generation": {"columns": [], "correlations": "Identity", "cluster": None, "uploaded_frame": None}}
    with pytest.raises(M.InvalidSpec, match=r"contract 0 .*missing field"):
        M.find_power_contracts([bad], 100, 16)
    with pytest.raises(M.InvalidSpec, match="cannot be empty"):
        M.find_sample_size_contracts([], 20, 100, 16)


def test_no_kernel_is_instantiated_in_two_translation_units():
    """Every translation unit carries its own __constant__ plan; a template kernel instantiated with the same arguments in two
    of them is merged by the linker, and a launch may then run the other module's copy against ITS (stale) plan.  So each
    kernel symbol must live in exactly one cubin of the library."""
    import shutil
    import subprocess
    from collections import Counter
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(tool).exists():
        pytest.skip("cuobjdump not available")
    out = subprocess.run([tool, "-res-usage", str(ROOT / "mcpower_b200" / "libmcpower_b200.so")], capture_output=True, text=True).stdout
    names = [ln.split("Function", 1)[1].strip().rstrip(":") for ln in out.splitlines() if ln.lstrip().startswith("Function ")]
    assert len(names) > 100, "cuobjdump listed no kernels"
    dupes = [n for n, c in Counter(names).items() if c > 1]
    assert not dupes, dupes[:5]


def test_no_cpu_fallback():
    """Without a device the product must fail loudly, never compute on the host."""
    if M.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(M.CudaError, match="no CPU fallback"):
        M.run_batch(S.headline_spec(), [500], 8, 1)
    with pytest.raises(M.CudaError):
        M.generate_sim(S.headline_spec(), 1, 0, 16)


def test_product_does_not_reference_the_oracle():
    for path in list((ROOT / "mcpower_b200").rglob("*.py")) + list((ROOT / "mcpower_b200" / "csrc").glob("*")):
        if path.is_file() and path.suffix in (".py", ".cu", ".cuh", ".hpp", ".h"):
            text = path.read_text()
            assert "mcp_oracle" not in text and "from oracle" not in text and "import oracle" not in text, path


def test_counters_from_bytes_matches_manual_aggregation():
    rng = np.random.default_rng(1)
    unc = (rng.random((50, 3, 4)) < 0.4).astype(np.uint8)
    cor = (unc & (rng.random((50, 3, 4)) < 0.7)).astype(np.uint8)
    conv = np.ones((50, 3), dtype=np.uint8)
    c = M.counters_from_bytes(unc, cor, conv, None, None)
    assert c.shape == (3, 28)  # 4T + 12 + 2F + 4C with T = 4, F = C = 0
    assert c[1, 2] == unc[:, 1, 2].sum() and c[2, 4 + 1] == cor[:, 2, 1].sum() and c[0, 8] == 50
    assert c[0, 10:15].sum() == 50 and c[0, 15:20].sum() == 50


def test_reference_arm_of_bench_prints_the_contract_line():
    """`bench.py --impl reference` times the CPU restatement alone (no GPU needed) and prints ONE JSON line with the arm's
    contract keys: same metric / unit / config as the CUDA arm, impl = reference, a cpu_baseline describing the run and an e2e
    object that repeats the value with zero copy bytes."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-400:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "ols_fits_per_sec" and d["unit"] == "fits/s"
    assert d["higher_is_better"] is True and d["steps"] == 1 and d["n_gpus"] == 1 and d["value"] > 0
    assert d["config"]["workload"].startswith("ols_n500_k8_p9") and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "fits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
