"""Host-side mirror of `engine_core::run_batch` over the C ABI of libmcpower_b200.so.

Reference interface (crates/engine-core/src/batch.rs:130-146, spec.rs:19-52,552-600):

    run_batch(spec, sample_sizes, n_sims, base_seed, sink) -> Result<BatchResult, EngineError>

Same argument meaning, same error classes (InvalidSpec / CorrelationNotPSD / RankDeficient /
Cancelled), same sim-major integer tensors.  There is no CPU execution path in this package: if the
shared library is missing or no CUDA device is present the call raises — it never falls back.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from dataclasses import dataclass, field
from pathlib import Path
from typing import Callable, Optional, Sequence

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = Path(os.environ.get("MCPOWER_B200_LIB", _HERE / "libmcpower_b200.so"))

MODE_COUNTS, MODE_BYTES, MODE_CAPTURE = 1, 2, 4


class EngineError(Exception):
    """Base of the EngineError family (spec.rs:19-29)."""
    code = -1


class InvalidSpec(EngineError, ValueError):
    code = 1


class CorrelationNotPSD(EngineError, ValueError):
    code = 2


class RankDeficient(EngineError, ValueError):
    code = 3


class Cancelled(EngineError, KeyboardInterrupt):
    """PyO3 maps EngineError::Cancelled to KeyboardInterrupt (engine-py/src/lib.rs:141-167)."""
    code = 4


class CudaError(EngineError, RuntimeError):
    code = 5


class Unsupported(EngineError, NotImplementedError):
    code = 6


_ERRORS = {c.code: c for c in (InvalidSpec, CorrelationNotPSD, RankDeficient, Cancelled, CudaError, Unsupported)}


_CANCEL = C.CFUNCTYPE(C.c_int, C.c_void_p)


class _Options(C.Structure):
    """McpOptions (include/mcpower_b200.h)."""
    _fields_ = [("device", C.c_int32), ("mode", C.c_uint32), ("sim_begin", C.c_uint64), ("wave_sims", C.c_uint32),
                ("n_devices", C.c_int32), ("is_cancelled", _CANCEL)]


def _options(device=-1, mode=1, sim_begin=0, wave_sims=0, n_devices=0, is_cancelled=None) -> "_Options":
    return _Options(device, mode, sim_begin, wave_sims, n_devices, is_cancelled if is_cancelled is not None else C.cast(None, _CANCEL))


class _Result(C.Structure):
    _fields_ = [("n_sims", C.c_uint32), ("n_sample_sizes", C.c_uint32), ("n_targets", C.c_uint32),
                ("n_factors", C.c_uint32),
                ("uncorrected", C.POINTER(C.c_uint8)), ("corrected", C.POINTER(C.c_uint8)),
                ("converged", C.POINTER(C.c_uint8)), ("overall", C.POINTER(C.c_uint8)),
                ("factor_excluded", C.POINTER(C.c_uint8)), ("joint", C.POINTER(C.c_uint8)),
                ("boundary_hit", C.POINTER(C.c_uint8)), ("t_sq", C.POINTER(C.c_double)),
                ("counter_width", C.c_uint32), ("gpu_launches", C.c_uint32),
                ("counters", C.POINTER(C.c_uint64)), ("kernel_ms", C.c_double),
                ("n_posthoc", C.c_uint32), ("reserved", C.c_uint32),
                ("posthoc_unc", C.POINTER(C.c_uint8)), ("posthoc_cor", C.POINTER(C.c_uint8)),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("n_devices", C.c_uint32), ("n_waves", C.c_uint32),
                ("lower_ms", C.c_double), ("enqueue_ms", C.c_double), ("wait_ms", C.c_double)]


_PROGRESS = C.CFUNCTYPE(C.c_int, C.c_uint64, C.c_uint64, C.c_void_p)

EXPORTS = ("mcp_run_batch", "mcp_free_result", "mcp_generate_sim", "mcp_fit_data", "mcp_crit_table", "mcp_splitmix64",
           "mcp_pipe_bench", "mcp_device_count", "mcp_last_error", "mcp_version", "mcp_find_power", "mcp_find_sample_size",
           "mcp_free_string", "mcp_fit_crossing", "mcp_build_grid", "mcp_wilson_ci", "mcp_lower_contracts",
           "mcp_find_power_contracts", "mcp_find_sample_size_contracts", "mcp_test_force_ols_variant",
           "mcp_fit_provided_data")


class _FitInfo(C.Structure):
    """McpFitInfo (include/mcpower_b200.h)."""
    _fields_ = [("converged", C.c_int32), ("boundary_hit", C.c_int32), ("iterations", C.c_int32), ("n_targets", C.c_int32),
                ("crit_sq_uncorrected", C.c_double), ("df_resid", C.c_double), ("sigma_sq_hat", C.c_double),
                ("tau_sq_hat", C.c_double), ("deviance", C.c_double), ("deviance_null", C.c_double),
                ("slope_var_hat", C.c_double * 2), ("slope_corr_hat", C.c_double * 3), ("extra_var_hat", C.c_double * 2),
                ("pinned_components", C.c_int32)]


class _CrossingFit(C.Structure):
    _fields_ = [("status", C.c_int32), ("has_ci_lo", C.c_int32), ("has_ci_hi", C.c_int32), ("has_n_approx", C.c_int32),
                ("n_star", C.c_double), ("ci_lo", C.c_double), ("ci_hi", C.c_double), ("max_violation", C.c_double),
                ("n_achievable", C.c_uint64), ("n_min", C.c_uint64), ("n_approx", C.c_uint64)]

_lib = None


def lib():
    """Load libmcpower_b200.so (built in-tree by mcpower_b200/build.py). Raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not _SO.exists():
        raise CudaError(f"{_SO} is missing — build it with `python -m mcpower_b200.build` "
                        "(the CUDA backend has no CPU fallback)")
    L = C.CDLL(str(_SO))
    L.mcp_run_batch.restype = C.c_int
    L.mcp_run_batch.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_uint32), C.c_uint32, C.c_uint32, C.c_uint64,
                                C.POINTER(_Options), _PROGRESS, C.c_void_p, C.POINTER(C.POINTER(_Result))]
    L.mcp_free_result.argtypes = [C.POINTER(_Result)]
    L.mcp_free_result.restype = None
    L.mcp_generate_sim.restype = C.c_int
    L.mcp_generate_sim.argtypes = [C.c_char_p, C.c_size_t, C.c_uint32, C.c_uint64, C.c_uint64, C.c_int32,
                                   C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.mcp_fit_data.restype = C.c_int
    L.mcp_fit_data.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_uint32, C.c_uint32, C.c_int32,
                               C.POINTER(C.c_double)]
    L.mcp_fit_provided_data.restype = C.c_int
    L.mcp_fit_provided_data.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_double), C.c_uint32, C.c_uint32,
                                        C.POINTER(C.c_double), C.POINTER(C.c_uint32), C.c_int32, C.POINTER(C.c_double),
                                        C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_FitInfo)]
    L.mcp_crit_table.restype = C.c_int
    L.mcp_crit_table.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_uint32), C.c_uint32, C.c_uint32,
                                 C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                 C.POINTER(C.c_double)]
    L.mcp_splitmix64.restype = C.c_uint64
    L.mcp_splitmix64.argtypes = [C.c_uint64]
    L.mcp_pipe_bench.restype = C.c_int
    L.mcp_pipe_bench.argtypes = [C.c_int32, C.c_int32, C.POINTER(C.c_double)]
    L.mcp_find_power.restype = C.c_int
    L.mcp_find_power.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_uint32), C.c_uint32, C.c_uint32, C.c_uint64,
                                 C.POINTER(_Options), _PROGRESS, C.c_void_p, C.POINTER(C.c_void_p)]
    L.mcp_find_sample_size.restype = C.c_int
    L.mcp_find_sample_size.argtypes = [C.c_char_p, C.c_size_t, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int32, C.c_int32,
                                       C.c_double, C.c_uint32, C.c_uint64, C.POINTER(_Options), _PROGRESS, C.c_void_p,
                                       C.POINTER(C.c_void_p)]
    L.mcp_lower_contracts.restype = C.c_int
    L.mcp_lower_contracts.argtypes = [C.c_char_p, C.c_size_t, C.c_int32, C.POINTER(C.c_void_p)]
    L.mcp_find_power_contracts.restype = C.c_int
    L.mcp_find_power_contracts.argtypes = [C.c_char_p, C.c_size_t, C.c_int32, C.c_uint64, C.c_uint32, C.c_uint64,
                                           C.POINTER(_Options), _PROGRESS, C.c_void_p, C.POINTER(C.c_void_p)]
    L.mcp_find_sample_size_contracts.restype = C.c_int
    L.mcp_find_sample_size_contracts.argtypes = [C.c_char_p, C.c_size_t, C.c_int32, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int32,
                                                 C.c_int32, C.c_double, C.c_uint32, C.c_uint64, C.POINTER(_Options), _PROGRESS,
                                                 C.c_void_p, C.POINTER(C.c_void_p)]
    L.mcp_free_string.restype = None
    L.mcp_free_string.argtypes = [C.c_void_p]
    L.mcp_fit_crossing.restype = C.c_int
    L.mcp_fit_crossing.argtypes = [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_uint32, C.c_uint64, C.c_double,
                                   C.c_uint64, C.POINTER(_CrossingFit)]
    L.mcp_build_grid.restype = C.c_int
    L.mcp_build_grid.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int32, C.c_int32, C.c_uint64, C.c_uint64,
                                 C.POINTER(C.c_uint64), C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.mcp_wilson_ci.restype = None
    L.mcp_wilson_ci.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.mcp_device_count.restype = C.c_int
    L.mcp_test_force_ols_variant.restype = C.c_int
    L.mcp_test_force_ols_variant.argtypes = [C.c_int]
    L.mcp_last_error.restype = C.c_char_p
    L.mcp_version.restype = C.c_char_p
    _lib = L
    return L


def _raise(code: int):
    msg = lib().mcp_last_error().decode("utf-8", "replace")
    raise _ERRORS.get(code, EngineError)(msg)


def splitmix64(z: int) -> int:
    """rng::splitmix64_finalize — hosts pass base_seed = splitmix64(user_seed) (find_power.rs:259)."""
    return int(lib().mcp_splitmix64(z & 0xFFFFFFFFFFFFFFFF))


@dataclass
class BatchResult:
    """BatchResult (spec.rs:552-600). Byte tensors are None unless MODE_BYTES was requested."""
    n_sims: int
    n_sample_sizes: int
    n_targets: int
    n_factors: int
    counters: np.ndarray                       # [n_ss, 4T+7+2F+4C] uint64 (layout: include/mcpower_b200.h)
    uncorrected: Optional[np.ndarray] = None   # [n_sims, n_ss, T] uint8
    corrected: Optional[np.ndarray] = None
    converged: Optional[np.ndarray] = None     # [n_sims, n_ss]
    overall: Optional[np.ndarray] = None
    factor_excluded: Optional[np.ndarray] = None
    joint: Optional[np.ndarray] = None         # [n_sims, n_ss] joint Wald decision (estimator mle)
    boundary_hit: Optional[np.ndarray] = None  # [n_sims, n_ss] (estimator mle) bits 0-1: 0 none / 1 a variance component at 0 /
    #                                            2 optimiser at its bound; bits 2-6: pinned_components (see the properties below)
    t_sq: Optional[np.ndarray] = None          # [n_sims, n_ss, T] float64 (MODE_CAPTURE)
    n_posthoc: int = 0                         # C = pairwise posthoc contrasts over all blocks
    posthoc_unc: Optional[np.ndarray] = None   # [n_sims, n_ss, C] uint8 (BatchResult::posthoc_unc, spec.rs:557-559)
    posthoc_cor: Optional[np.ndarray] = None
    kernel_ms: float = 0.0
    gpu_launches: int = 0
    h2d_bytes: int = 0                         # bytes the call copied host → device (plan + tables, every device)
    d2h_bytes: int = 0                         # bytes the call copied device → host (counters, byte tensors, capture)
    n_devices: int = 1
    n_waves: int = 1
    lower_ms: float = 0.0                      # host: validation + lowering
    enqueue_ms: float = 0.0                    # host: allocation + enqueueing uploads / launches / copies on every device
    wait_ms: float = 0.0                       # host: waiting for the devices (kernels + copies)
    extras: dict = field(default_factory=dict)

    # ---- views of the counter rows (what aggregate_batch derives, aggregation.rs:91-242)
    @property
    def success_uncorrected(self):
        return self.counters[:, :self.n_targets]

    @property
    def success_corrected(self):
        return self.counters[:, self.n_targets:2 * self.n_targets]

    @property
    def n_converged(self):
        return self.counters[:, 2 * self.n_targets]

    @property
    def boundary_code(self):
        """BatchResult::boundary_hit (spec.rs:562-565): 0 none, 1 a variance component at 0, 2 optimiser at its bound."""
        return None if self.boundary_hit is None else self.boundary_hit & 3

    @property
    def pinned_components(self):
        """BatchResult::pinned_components (spec.rs:566-574): bit k set iff variance component k of the fitted model sat at 0."""
        return None if self.boundary_hit is None else (self.boundary_hit >> 2).astype(np.uint64)

    @property
    def pinned_component_counts(self):
        """[n_ss, 5]: converged fits with component k (intercept, slopes, extra groupings) pinned — EstimatorExtras::Mle."""
        return self.counters[:, -5:]

    @property
    def success_overall(self):
        return self.counters[:, 2 * self.n_targets + 1]

    @property
    def _j(self):
        return self.n_targets + self.n_posthoc  # tests pooled in the joint histograms (aggregation.rs:186-217)

    @property
    def hist_uncorrected(self):
        T = self.n_targets
        return self.counters[:, 2 * T + 2:2 * T + 3 + self._j]

    @property
    def hist_corrected(self):
        o = 2 * self.n_targets + 3 + self._j
        return self.counters[:, o:o + self._j + 1]

    @property
    def _off_fx(self):
        return 2 * self.n_targets + 4 + 2 * self._j

    @property
    def n_factor_excluded(self):
        return self.counters[:, self._off_fx:self._off_fx + self.n_factors]

    @property
    def success_joint(self):
        return self.counters[:, self._off_fx + self.n_factors]

    @property
    def n_boundary(self):
        return self.counters[:, self._off_fx + self.n_factors + 1]

    @property
    def success_posthoc_uncorrected(self):
        o = self._off_fx + self.n_factors + 2
        return self.counters[:, o:o + self.n_posthoc]

    @property
    def success_posthoc_corrected(self):
        o = self._off_fx + self.n_factors + 2 + self.n_posthoc
        return self.counters[:, o:o + self.n_posthoc]

    @property
    def n_singular(self):
        return self.counters[:, self._off_fx + self.n_factors + 2 + 2 * self.n_posthoc]

    @property
    def n_factor_separation(self):
        o = self._off_fx + self.n_factors + 3 + 2 * self.n_posthoc
        return self.counters[:, o:o + self.n_factors]

    def power(self, corrected: bool = False) -> np.ndarray:
        """successes / n_sims per (sample size, target) — PowerResult.power (aggregation.rs)."""
        s = self.success_corrected if corrected else self.success_uncorrected
        return s.astype(np.float64) / max(self.n_sims, 1)


def counters_from_bytes(uncorrected, corrected, converged, overall=None, factor_excluded=None, joint=None,
                        boundary_hit=None, posthoc_unc=None, posthoc_cor=None) -> np.ndarray:
    """aggregate_batch's integer reductions (aggregation.rs:91-242) over the byte tensors —
    used to prove COUNTS mode equals BYTES mode + host aggregation."""
    n_sims, n_ss, T = uncorrected.shape
    F = 0 if factor_excluded is None else factor_excluded.shape[2]
    Cph = 0 if posthoc_unc is None else posthoc_unc.shape[2]
    J = T + Cph
    out = np.zeros((n_ss, 4 * T + 12 + 2 * F + 4 * Cph), dtype=np.uint64)
    out[:, :T] = uncorrected.sum(axis=0)
    out[:, T:2 * T] = corrected.sum(axis=0)
    out[:, 2 * T] = converged.sum(axis=0)
    if overall is not None:
        out[:, 2 * T + 1] = overall.sum(axis=0)
    nu = uncorrected.sum(axis=2, dtype=np.int64)
    nc = corrected.sum(axis=2, dtype=np.int64)
    if Cph:
        nu = nu + posthoc_unc.sum(axis=2, dtype=np.int64)
        nc = nc + posthoc_cor.sum(axis=2, dtype=np.int64)
    o_hc = 2 * T + 3 + J
    o_fx = o_hc + J + 1
    for g in range(n_ss):
        out[g, 2 * T + 2:2 * T + 3 + J] = np.bincount(nu[:, g], minlength=J + 1)
        out[g, o_hc:o_hc + J + 1] = np.bincount(nc[:, g], minlength=J + 1)
    if F:
        out[:, o_fx:o_fx + F] = (factor_excluded == 1).sum(axis=0)
        out[:, o_fx + F + 3 + 2 * Cph:o_fx + 2 * F + 3 + 2 * Cph] = (factor_excluded == 2).sum(axis=0)
    if joint is not None:
        out[:, o_fx + F] = joint.sum(axis=0)
    if boundary_hit is not None:
        code, pins = boundary_hit & 3, boundary_hit >> 2  # BatchResult::boundary_hit | pinned_components << 2
        out[:, o_fx + F + 1] = (code != 0).sum(axis=0)
        out[:, o_fx + F + 2 + 2 * Cph] = ((code == 1) & (converged != 0)).sum(axis=0)
        for k in range(5):
            out[:, o_fx + 2 * F + 3 + 2 * Cph + k] = ((((pins >> k) & 1) != 0) & (converged != 0)).sum(axis=0)
    if Cph:
        out[:, o_fx + F + 2:o_fx + F + 2 + Cph] = posthoc_unc.sum(axis=0)
        out[:, o_fx + F + 2 + Cph:o_fx + F + 2 + 2 * Cph] = posthoc_cor.sum(axis=0)
    return out


ProgressSink = Callable[[int, int], bool]


class _ResultOwner:
    """Owns one McpBatchResult; mcp_free_result runs when the last view of its buffers is collected."""

    def __init__(self, ptr):
        self._ptr = ptr

    def __del__(self):
        p, self._ptr = self._ptr, None
        if p is not None and _lib is not None:
            _lib.mcp_free_result(p)


def force_ols_variant(variant: int) -> int:
    """Test hook: pin the OLS power kernel (1 thread per study, 4 lane quad solving by call, 5 lean pair, 6 split pair;
    0 = the library's own choice).  Process-wide; returns the previous setting.  The cross-check kernels compute the
    same bits — tests/test_gpu_parity.py runs every OLS case under each."""
    return int(lib().mcp_test_force_ols_variant(int(variant)))


def run_batch(spec: dict, sample_sizes: Sequence[int], n_sims: int, base_seed: int,
              sink: Optional[ProgressSink] = None, *, mode: int = MODE_COUNTS, device: int = -1, sim_begin: int = 0,
              wave_sims: int = 0, n_devices: int = 0, is_cancelled: Optional[Callable[[], bool]] = None) -> BatchResult:
    """engine_core::run_batch on the GPU. `sink(current, total) -> bool` mirrors ProgressSink::report
    (return False to cancel → raises Cancelled, partial results discarded, batch.rs:128-129); `is_cancelled() -> bool`
    mirrors ProgressSink::is_cancelled.  n_devices > 1 (or -1 = all) shards the sims over several GPUs inside the call."""
    L = lib()
    blob = json.dumps(spec).encode()
    ss = np.ascontiguousarray(np.asarray(list(sample_sizes), dtype=np.int64))
    if ss.size and (ss.min() < 0 or ss.max() > 0xFFFFFFFF):
        raise InvalidSpec("invalid spec: sample sizes must fit u32")
    ss32 = ss.astype(np.uint32)
    ccb = _CANCEL(lambda _u: 1 if is_cancelled() else 0) if is_cancelled else None
    opt = _options(device, mode, sim_begin, wave_sims, n_devices, ccb)
    out = C.POINTER(_Result)()
    cb = _PROGRESS(lambda cur, tot, _u: 1 if sink(int(cur), int(tot)) else 0) if sink else C.cast(None, _PROGRESS)
    rc = L.mcp_run_batch(blob, len(blob), ss32.ctypes.data_as(C.POINTER(C.c_uint32)), ss32.size, int(n_sims),
                         base_seed & 0xFFFFFFFFFFFFFFFF, C.byref(opt), cb, None, C.byref(out))
    if rc != 0:
        _raise(rc)
    # The tensors stay where the library put them (page-locked host memory the device copied into): the arrays below are
    # views, not copies, and the C result is released when the last of them — and the BatchResult — is gone.
    owner = _ResultOwner(out)
    r = out.contents
    n_ss, T, F = r.n_sample_sizes, r.n_targets, r.n_factors

    def take(ptr, shape, dtype):
        if not ptr:
            return None
        n = int(np.prod(shape))
        if n == 0:
            return np.zeros(shape, dtype=dtype)
        nbytes = n * np.dtype(dtype).itemsize
        buf = (C.c_uint8 * nbytes).from_address(C.addressof(ptr.contents))
        buf._owner = owner  # keeps the C result alive for as long as any view of it exists
        return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)

    res = BatchResult(
        n_sims=r.n_sims, n_sample_sizes=n_ss, n_targets=T, n_factors=F,
        counters=take(r.counters, (n_ss, r.counter_width), np.uint64),
        uncorrected=take(r.uncorrected, (r.n_sims, n_ss, T), np.uint8),
        corrected=take(r.corrected, (r.n_sims, n_ss, T), np.uint8),
        converged=take(r.converged, (r.n_sims, n_ss), np.uint8),
        overall=take(r.overall, (r.n_sims, n_ss), np.uint8),
        factor_excluded=take(r.factor_excluded, (r.n_sims, n_ss, F), np.uint8),
        n_posthoc=r.n_posthoc,
        posthoc_unc=take(r.posthoc_unc, (r.n_sims, n_ss, r.n_posthoc), np.uint8),
        posthoc_cor=take(r.posthoc_cor, (r.n_sims, n_ss, r.n_posthoc), np.uint8),
        joint=take(r.joint, (r.n_sims, n_ss), np.uint8),
        boundary_hit=take(r.boundary_hit, (r.n_sims, n_ss), np.uint8),
        t_sq=take(r.t_sq, (r.n_sims, n_ss, T), np.float64),
        kernel_ms=float(r.kernel_ms), gpu_launches=int(r.gpu_launches), h2d_bytes=int(r.h2d_bytes),
        d2h_bytes=int(r.d2h_bytes), n_devices=int(r.n_devices), n_waves=int(r.n_waves), lower_ms=float(r.lower_ms),
        enqueue_ms=float(r.enqueue_ms), wait_ms=float(r.wait_ms))
    res.extras["_owner"] = owner
    return res


def generate_sim(spec: dict, base_seed: int, sim_id: int, max_n: int, device: int = -1):
    """capture_sim0_data twin (introspect.rs:104-155) for any sim id → (X [max_n, P] f32, y [max_n] f32)."""
    from .spec import n_predictors
    blob = json.dumps(spec).encode()
    P = n_predictors(spec)
    x = np.zeros((P, max_n), dtype=np.float32)
    y = np.zeros(max_n, dtype=np.float32)
    rc = lib().mcp_generate_sim(blob, len(blob), max_n, base_seed & 0xFFFFFFFFFFFFFFFF, sim_id, device,
                                x.ctypes.data_as(C.POINTER(C.c_float)), y.ctypes.data_as(C.POINTER(C.c_float)))
    if rc != 0:
        _raise(rc)
    return x.T, y


def fit_data(x, y, device: int = -1) -> dict:
    """fit_provided_data twin (introspect.rs:174-293), OLS: β, SE, t², σ², SST, RSS on supplied bytes."""
    x = np.asarray(x, dtype=np.float64)
    n, p = x.shape
    xc = np.ascontiguousarray(x.T)
    yv = np.ascontiguousarray(y, dtype=np.float64)
    out = np.zeros(3 * p + 4)
    rc = lib().mcp_fit_data(xc.ctypes.data_as(C.POINTER(C.c_double)), yv.ctypes.data_as(C.POINTER(C.c_double)), n, p,
                            device, out.ctypes.data_as(C.POINTER(C.c_double)))
    if rc != 0:
        _raise(rc)
    return {"betas": out[:p].copy(), "se": out[p:2 * p].copy(), "t_sq": out[2 * p:3 * p].copy(),
            "sigma_sq": out[3 * p], "sst": out[3 * p + 1], "rss": out[3 * p + 2], "converged": bool(out[3 * p + 3])}


def fit_provided_data(spec: dict, x, y, cluster_ids=None, device: int = -1) -> dict:
    """introspect::fit_provided_data (introspect.rs:174-562): fit the supplied (design [n, P], outcome[, cluster ids]) with the
    shipped fit code of `spec["estimator"]` — OLS, unclustered GLM (IRLS from the truth start), random-intercept MLE (profiled
    REML) — and return IntrospectResults as a dict: betas per column, se / statistic per target, crit_sq_uncorrected, df_resid."""
    x = np.asarray(x, dtype=np.float64)
    n, p = x.shape
    xc = np.ascontiguousarray(x.T)
    yv = np.ascontiguousarray(y, dtype=np.float64)
    blob = json.dumps(spec).encode()
    nt = len(spec["target_indices"])
    betas, se, stat = np.zeros(p), np.zeros(max(nt, 1)), np.zeros(max(nt, 1))
    info = _FitInfo()
    ids = None if cluster_ids is None else np.ascontiguousarray(cluster_ids, dtype=np.uint32)
    dp = C.POINTER(C.c_double)
    rc = lib().mcp_fit_provided_data(blob, len(blob), xc.ctypes.data_as(dp), n, p, yv.ctypes.data_as(dp),
                                     None if ids is None else ids.ctypes.data_as(C.POINTER(C.c_uint32)), device,
                                     betas.ctypes.data_as(dp), se.ctypes.data_as(dp), stat.ctypes.data_as(dp), C.byref(info))
    if rc != 0:
        _raise(rc)
    return {"betas": betas, "se": se[:nt], "statistic": stat[:nt], "target_cols": list(spec["target_indices"]),
            "converged": bool(info.converged), "crit_sq_uncorrected": info.crit_sq_uncorrected, "df_resid": info.df_resid,
            "boundary_hit": int(info.boundary_hit), "iterations": int(info.iterations), "sigma_sq_hat": info.sigma_sq_hat,
            "tau_sq_hat": info.tau_sq_hat, "deviance": info.deviance, "deviance_null": info.deviance_null,
            "slope_var_hat": [float(v) for v in info.slope_var_hat], "slope_corr_hat": [float(v) for v in info.slope_corr_hat], "extra_var_hat": [float(v) for v in info.extra_var_hat], "pinned_components": int(info.pinned_components)}


def crit_table(spec: dict, sample_sizes: Sequence[int], n_predictors: Optional[int] = None) -> dict:
    """CritValueTable::build_with_tukey_k twin (host f64; runs without a GPU)."""
    from .spec import n_predictors as npred
    blob = json.dumps(spec).encode()
    ss = np.asarray(list(sample_sizes), dtype=np.uint32)
    T = len(spec["target_indices"]) + len(spec.get("contrast_pairs", []) or [])
    P = npred(spec) if n_predictors is None else n_predictors
    t_unc, corr, post, ov = np.zeros(len(ss)), np.zeros((len(ss), max(T, 1))), np.zeros(len(ss)), np.zeros(len(ss))
    dp = C.POINTER(C.c_double)
    rc = lib().mcp_crit_table(blob, len(blob), ss.ctypes.data_as(C.POINTER(C.c_uint32)), len(ss), P,
                              t_unc.ctypes.data_as(dp), corr.ctypes.data_as(dp), post.ctypes.data_as(dp),
                              ov.ctypes.data_as(dp))
    if rc != 0:
        _raise(rc)
    return {"t_crit_sq_uncorrected": t_unc, "correction_t_crit_sq": corr[:, :T], "posthoc_t_crit_sq": post,
            "overall_crit": ov}


def pipe_bench(kind: int, device: int = -1) -> float:
    """Measured thread-operations/s of one instruction class (roofline denominators; see the C header)."""
    v = C.c_double()
    rc = lib().mcp_pipe_bench(kind, device, C.byref(v))
    if rc != 0:
        _raise(rc)
    return v.value


def device_count() -> int:
    return int(lib().mcp_device_count())


# ---------------------------------------------------------------------------------------------------------------------
# Orchestrator entry points for one lowered scenario (engine-orchestrator: find_power / find_sample_size)
def _take_json(ptr) -> object:
    try:
        return json.loads(C.string_at(ptr).decode("utf-8"))
    finally:
        lib().mcp_free_string(ptr)


def find_power(spec: dict, sample_sizes: Sequence[int], n_sims: int, base_seed: int, *, mode: int = MODE_COUNTS,
               device: int = -1, sim_begin: int = 0) -> list:
    """find_power for one scenario: run_batch + aggregate_batch → one PowerResult dict per sample size
    (field names of engine-orchestrator/src/result.rs:54-144)."""
    blob = json.dumps(spec).encode()
    ss32 = np.ascontiguousarray(np.asarray(list(sample_sizes), dtype=np.uint32))
    opt = _options(device, mode, sim_begin)
    out = C.c_void_p()
    rc = lib().mcp_find_power(blob, len(blob), ss32.ctypes.data_as(C.POINTER(C.c_uint32)), ss32.size, int(n_sims),
                              base_seed & 0xFFFFFFFFFFFFFFFF, C.byref(opt), C.cast(None, _PROGRESS), None, C.byref(out))
    if rc != 0:
        _raise(rc)
    return _take_json(out)


def find_sample_size(spec: dict, from_size: int, to_size: int, n_sims: int, base_seed: int, *, by: Optional[int] = None,
                     count: int = 12, log_grid: bool = False, target_power: float = 80.0, device: int = -1) -> dict:
    """find_sample_size for one scenario (find_sample_size.rs:83-257): `by` = fixed step (linear) / point count (log);
    by=None → ~`count` auto-placed points.  Returns the SampleSizeResult dict."""
    blob = json.dumps(spec).encode()
    opt = _options(device, MODE_COUNTS)
    out = C.c_void_p()
    rc = lib().mcp_find_sample_size(blob, len(blob), int(from_size), int(to_size), int(count if by is None else by),
                                    1 if by is None else 0, 1 if log_grid else 0, float(target_power), int(n_sims),
                                    base_seed & 0xFFFFFFFFFFFFFFFF, C.byref(opt), C.cast(None, _PROGRESS), None, C.byref(out))
    if rc != 0:
        _raise(rc)
    return _take_json(out)


def _contract_payload(contracts) -> tuple:
    """bytes → the msgpack `Vec<SimulationContract>` blob the reference hosts forward (format 0); a dict / list of dicts →
    the same tree as serde JSON (format 1)."""
    if isinstance(contracts, (bytes, bytearray, memoryview)):
        return bytes(contracts), 0
    return json.dumps(contracts if isinstance(contracts, list) else [contracts]).encode(), 1


def lower_contracts(contracts) -> list:
    """`lower_contracts` (find_power.rs:251-288) over `contract_to_simulation_spec` (contract_adapter.rs:74-230): one
    {"label", "spec", "report_target_indices", "report_contrast_pairs"} dict per contract."""
    blob, fmt = _contract_payload(contracts)
    out = C.c_void_p()
    rc = lib().mcp_lower_contracts(blob, len(blob), fmt, C.byref(out))
    if rc != 0:
        _raise(rc)
    return _take_json(out)


def find_power_contracts(contracts, sample_size: int, n_sims: int, user_seed: int = 2137, *, device: int = -1) -> dict:
    """find_power (find_power.rs:301-392) on the hosts' contract payload → {"scenarios": [[label, PowerResult], ...]}."""
    blob, fmt = _contract_payload(contracts)
    opt = _options(device, MODE_COUNTS)
    out = C.c_void_p()
    rc = lib().mcp_find_power_contracts(blob, len(blob), fmt, int(sample_size), int(n_sims), user_seed & 0xFFFFFFFFFFFFFFFF,
                                        C.byref(opt), C.cast(None, _PROGRESS), None, C.byref(out))
    if rc != 0:
        _raise(rc)
    return _take_json(out)


def find_sample_size_contracts(contracts, from_size: int, to_size: int, n_sims: int, user_seed: int = 2137, *,
                               by: Optional[int] = None, count: int = 12, log_grid: bool = False, target_power: float = 80.0,
                               device: int = -1) -> dict:
    """find_sample_size (find_sample_size.rs:83-257) on the hosts' contract payload →
    {"scenarios": [[label, SampleSizeResult], ...]}."""
    blob, fmt = _contract_payload(contracts)
    opt = _options(device, MODE_COUNTS)
    out = C.c_void_p()
    rc = lib().mcp_find_sample_size_contracts(blob, len(blob), fmt, int(from_size), int(to_size), int(count if by is None else by),
                                              1 if by is None else 0, 1 if log_grid else 0, float(target_power), int(n_sims),
                                              user_seed & 0xFFFFFFFFFFFFFFFF, C.byref(opt), C.cast(None, _PROGRESS), None,
                                              C.byref(out))
    if rc != 0:
        _raise(rc)
    return _take_json(out)


_FIT_STATUS = ("fitted", "at_or_below_min", "not_reached", "non_monotone")


def fit_crossing(grid: Sequence[int], counts: Sequence[int], n_sims: int, target_power: float, atom: int = 1) -> dict:
    """fit_crossing (fit.rs:45-120) → CrossingFit as a dict tagged by `status`."""
    g = np.ascontiguousarray(np.asarray(list(grid), dtype=np.uint64))
    k = np.ascontiguousarray(np.asarray(list(counts), dtype=np.uint64))
    f = _CrossingFit()
    rc = lib().mcp_fit_crossing(g.ctypes.data_as(C.POINTER(C.c_uint64)), k.ctypes.data_as(C.POINTER(C.c_uint64)), g.size,
                                int(n_sims), float(target_power), int(atom), C.byref(f))
    if rc != 0:
        _raise(rc)
    st = _FIT_STATUS[f.status]
    if st == "fitted":
        return {"status": st, "n_star": f.n_star, "n_achievable": int(f.n_achievable),
                "ci_lo": f.ci_lo if f.has_ci_lo else None, "ci_hi": f.ci_hi if f.has_ci_hi else None}
    if st == "at_or_below_min":
        return {"status": st, "n_min": int(f.n_min)}
    if st == "not_reached":
        return {"status": st, "n_approx": int(f.n_approx) if f.has_n_approx else None}
    return {"status": st, "max_violation": f.max_violation}


def build_grid(from_size: int, to_size: int, *, by: Optional[int] = None, count: int = 12, log_grid: bool = False,
               atom: int = 1, hard_min: int = 1):
    """build_grid (grid.rs:37-147) → (grid list, number of advisory warnings)."""
    buf = np.zeros(1 << 16, dtype=np.uint64)
    n, w = C.c_uint32(), C.c_uint32()
    rc = lib().mcp_build_grid(int(from_size), int(to_size), int(count if by is None else by), 1 if by is None else 0,
                              1 if log_grid else 0, int(atom), int(hard_min), buf.ctypes.data_as(C.POINTER(C.c_uint64)),
                              buf.size, C.byref(n), C.byref(w))
    if rc != 0:
        _raise(rc)
    return [int(v) for v in buf[:n.value]], int(w.value)


def wilson_ci(k: int, n: int):
    lo, hi = C.c_double(), C.c_double()
    lib().mcp_wilson_ci(int(k), int(n), C.byref(lo), C.byref(hi))
    return lo.value, hi.value
