"""ctypes loader for the CPU ORACLE (test infrastructure, NOT product code).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  It converts a `SimulationSpec` dict (the serde JSON
shape of crates/engine-core/src/spec.rs:250-417) into the oracle's C struct and
exposes the oracle entry points as numpy-friendly functions.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_ref" / "libmcp_oracle.so"

DIST = {"normal": 0, "binary": 1, "right_skewed": 2, "left_skewed": 3, "high_kurtosis": 4, "uniform": 5,
        "uploaded_factor": 6, "uploaded_binary": 7, "uploaded_data": 8}
RES = {"normal": 0, "right_skewed": 1, "left_skewed": 2, "high_kurtosis": 3, "uniform": 4}
OUT = {"continuous": 0, "binary": 1, "count": 2}
EST = {"ols": 0, "glm": 1, "mle": 2}
CORR = {"none": 0, "bonferroni": 1, "holm": 2, "benjamini_hochberg": 3, "tukey_hsd": 4}


class OrcSpec(C.Structure):
    _fields_ = [
        ("n_non_factor", C.c_int32), ("n_factor_dummies", C.c_int32), ("n_factors", C.c_int32),
        ("n_interactions", C.c_int32),
        ("correlation", C.POINTER(C.c_double)), ("var_types", C.POINTER(C.c_int32)),
        ("var_pinned", C.POINTER(C.c_uint8)), ("var_params", C.POINTER(C.c_double)),
        ("factor_n_levels", C.POINTER(C.c_int32)), ("factor_proportions", C.POINTER(C.c_double)),
        ("factor_sampled", C.POINTER(C.c_int32)), ("inter_ptr", C.POINTER(C.c_int32)),
        ("inter_cols", C.POINTER(C.c_int32)), ("effect_sizes", C.POINTER(C.c_double)),
        ("n_marginal_targets", C.c_int32), ("target_indices", C.POINTER(C.c_int32)),
        ("n_contrasts", C.c_int32), ("contrast_pairs", C.POINTER(C.c_int32)),
        ("n_fit_columns", C.c_int32), ("fit_columns", C.POINTER(C.c_int32)),
        ("correction_method", C.c_int32), ("alpha", C.c_double), ("posthoc_alpha", C.c_double),
        ("hsk_driver", C.c_int32),
        ("residual_dist", C.c_int32), ("residual_pinned", C.c_int32), ("outcome_kind", C.c_int32),
        ("link_probit", C.c_int32), ("estimator", C.c_int32),
        ("report_overall", C.c_int32), ("factor_min_level_count", C.c_int32),
        ("has_cluster", C.c_int32), ("sizing_fixed_size", C.c_int32), ("sizing_value", C.c_int32),
        ("tau_squared", C.c_double),
        ("heterogeneity", C.c_double), ("hsk_ratio", C.c_double), ("corr_noise_sd", C.c_double),
        ("dist_change_prob", C.c_double),
        ("n_new_dists", C.c_int32), ("new_dists", C.POINTER(C.c_int32)),
        ("resid_change_prob", C.c_double),
        ("n_resid_dists", C.c_int32), ("resid_dists", C.POINTER(C.c_int32)),
        ("residual_df", C.c_double),
        ("sampled_factor_proportions", C.c_int32),
        ("has_lme", C.c_int32), ("re_dist", C.c_int32),
        ("re_df", C.c_double), ("icc_noise_sd", C.c_double),
        ("n_posthoc", C.c_int32), ("posthoc_pairs", C.POINTER(C.c_int32)),
        ("n_posthoc_blocks", C.c_int32), ("posthoc_block_levels", C.POINTER(C.c_int32)),
        ("posthoc_block_ncontrasts", C.POINTER(C.c_int32)),
        ("upload_normal", C.POINTER(C.c_double)), ("un_rows", C.c_int32), ("un_cols", C.c_int32),
        ("upload_data", C.POINTER(C.c_double)), ("ud_rows", C.c_int32), ("ud_cols", C.c_int32),
        ("n_frame_map", C.c_int32), ("frame_map", C.POINTER(C.c_int32)),
        ("n_var_params", C.c_int32),
        ("n_between", C.c_int32), ("between_cols", C.POINTER(C.c_int32)),
        ("n_slopes", C.c_int32), ("slope_col", C.c_int32 * 2), ("slope_variance", C.c_double * 2), ("slope_corr", C.c_double * 2),
        ("slope_corr12", C.c_double),
        ("n_extra", C.c_int32), ("extra_nested", C.c_int32 * 2), ("extra_count", C.c_int32 * 2), ("extra_tau_squared", C.c_double * 2),
    ]


_lib = None


def build(force: bool = False) -> Path:
    """Compile oracle/mcp_oracle.c → oracle/_ref/libmcp_oracle.so (gcc; see Makefile)."""
    src_m = max((_HERE / f).stat().st_mtime for f in ("mcp_oracle.c", "mcp_oracle.h", "Makefile"))
    if force or not _SO.exists() or _SO.stat().st_mtime < src_m:
        subprocess.run(["make", "-C", str(_HERE), "-B"], check=True, capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not _SO.exists():
        build()
    L = C.CDLL(str(_SO))
    u64, u32, i32, dbl = C.c_uint64, C.c_uint32, C.c_int32, C.c_double
    P = C.POINTER
    L.orc_splitmix64.restype = u64
    L.orc_splitmix64.argtypes = [u64]
    L.orc_pcg_mix64.restype = u64
    L.orc_pcg_mix64.argtypes = [u64, u64]
    L.orc_philox4x32_10.argtypes = [P(u32), P(u32), P(u32)]
    L.orc_u32_to_unit_f32.restype = C.c_float
    L.orc_u32_to_unit_f32.argtypes = [u32]
    L.orc_norm_inv_cdf_f32.restype = C.c_float
    L.orc_norm_inv_cdf_f32.argtypes = [C.c_float]
    L.orc_seq_stream.argtypes = [u64, u64, C.c_int, C.c_int, P(u32)]
    L.orc_seq_categorical.argtypes = [u64, u64, P(dbl), C.c_int, C.c_int, P(i32)]
    L.orc_planar_normals.argtypes = [u64, u64, u32, u32, C.c_int, P(u32)]
    L.orc_sample_t.argtypes = [u64, u64, dbl, C.c_int, P(dbl)]
    L.orc_draw_residual.argtypes = [u64, u64, C.c_int, dbl, C.c_int, P(dbl)]
    for name, n in (("orc_exp", 1), ("orc_log", 1), ("orc_phi", 1), ("orc_t3_lookup", 1), ("orc_norm_ppf", 1),
                    ("orc_t_ppf", 2), ("orc_chi2_ppf", 2), ("orc_f_ppf", 3), ("orc_q_tukey_ppf", 3),
                    ("orc_poisson_quantile", 2)):
        f = getattr(L, name)
        f.restype = dbl
        f.argtypes = [dbl] * n
    L.orc_apply_marginal.restype = dbl
    L.orc_apply_marginal.argtypes = [dbl, C.c_int, dbl]
    L.orc_apply_marginal_upload.restype = dbl
    L.orc_apply_marginal_upload.argtypes = [dbl, C.c_int, dbl, P(dbl), C.c_int, C.c_int, C.c_int]
    L.orc_build_crit_table.argtypes = [P(OrcSpec), P(u32), C.c_int, C.c_int, P(dbl), P(dbl), P(dbl), P(dbl)]
    L.orc_apply_correction.argtypes = [C.c_int, P(dbl), C.c_int, dbl, P(dbl), P(C.c_uint8)]
    L.orc_generate_sim.argtypes = [P(OrcSpec), u64, u64, C.c_int, P(C.c_float), P(C.c_float), P(i32)]
    L.orc_design_factor.argtypes = [P(OrcSpec), u64, u64, P(dbl)]
    L.orc_fit_lmm_slope.argtypes = [P(dbl), P(dbl), P(i32), C.c_int, C.c_int, C.c_int, P(dbl), P(dbl), P(dbl), P(C.c_int), P(C.c_int), P(C.c_int)]
    L.orc_extra_level_of_row.argtypes = [P(OrcSpec), C.c_int, C.c_int]
    L.orc_extra_levels_at.argtypes = [P(OrcSpec), C.c_int, C.c_int]
    L.orc_fit_lmm_general.argtypes = [P(dbl), P(dbl), P(i32), C.c_int, C.c_int, C.c_int, P(i32), C.c_int, P(dbl), P(dbl), P(dbl),
                                      P(C.c_int), P(C.c_int), P(C.c_int)]
    L.orc_fit_lmm_slopes.argtypes = [P(dbl), P(dbl), P(i32), C.c_int, C.c_int, C.c_int, P(i32), P(dbl), P(dbl), P(dbl), P(C.c_int), P(C.c_int),
                                     P(C.c_int)]
    L.orc_fit_ols.argtypes = [P(dbl), P(dbl), C.c_int, C.c_int, P(dbl), P(dbl), P(dbl), P(dbl), P(dbl), P(dbl),
                              P(C.c_int)]
    L.orc_fit_glm.argtypes = [P(dbl), P(dbl), C.c_int, C.c_int, C.c_int, P(dbl), P(dbl), P(dbl), P(dbl), P(dbl),
                              P(C.c_int), P(C.c_int)]
    L.orc_fit_lmm.argtypes = [P(dbl), P(dbl), P(i32), C.c_int, C.c_int, P(dbl), P(dbl), P(dbl), P(dbl), P(C.c_int),
                              P(C.c_int)]
    L.orc_run_batch.argtypes = [P(OrcSpec), P(u32), C.c_int, u32, u64, u64, C.c_int, P(C.c_uint8), P(C.c_uint8),
                                P(C.c_uint8), P(C.c_uint8), P(C.c_uint8), P(dbl), P(C.c_uint8), P(C.c_uint8)]
    L.orc_last_error.restype = C.c_char_p
    _lib = L
    return L


class OracleError(ValueError):
    def __init__(self, code, msg):
        super().__init__(f"oracle error {code}: {msg}")
        self.code = code


def _ptr(a, ty):
    return a.ctypes.data_as(C.POINTER(ty))


class PackedSpec:
    """Owns the numpy buffers an OrcSpec points into."""

    def __init__(self, spec: dict):
        k = int(spec["n_non_factor"])
        self.keep = []

        def arr(vals, dtype):
            a = np.ascontiguousarray(np.asarray(list(vals), dtype=dtype))
            if a.size == 0:
                a = np.zeros(1, dtype=dtype)
            self.keep.append(a)
            return a

        s = OrcSpec()
        s.n_non_factor = k
        s.n_factor_dummies = int(spec["n_factor_dummies"])
        levels = list(spec.get("factor_n_levels", []))
        s.n_factors = len(levels)
        inter = spec.get("interactions", []) or []
        s.n_interactions = len(inter)
        s.correlation = _ptr(arr(spec["correlation"], np.float64), C.c_double)
        s.var_types = _ptr(arr([DIST[v] for v in spec["var_types"]], np.int32), C.c_int32)
        pinned = list(spec.get("var_pinned", []) or [])
        pinned = (pinned + [False] * k)[:k]
        s.var_pinned = _ptr(arr([1 if v else 0 for v in pinned], np.uint8), C.c_uint8)
        vp = list(spec.get("var_params", []) or [])
        s.n_var_params = min(len(vp), k)
        vp = (vp + [0.5] * k)[:k]  # data_gen.rs:347 unwrap_or(0.5)
        # uploaded-data arms (spec.rs upload_* fields): NORTA table, raw frame, strict-bootstrap column map
        un = list(spec.get("upload_normal", []) or [])
        uns = list(spec.get("upload_normal_shape", [0, 0]) or [0, 0])
        ud = list(spec.get("upload_data", []) or [])
        uds = list(spec.get("upload_data_shape", [0, 0]) or [0, 0])
        fm = list(spec.get("bootstrap_frame_map", []) or [])
        if len(un) != int(uns[0]) * int(uns[1]) or len(ud) != int(uds[0]) * int(uds[1]):
            raise ValueError("invalid spec: upload table length does not match its shape")
        s.upload_normal = _ptr(arr(un, np.float64), C.c_double)
        s.un_rows, s.un_cols = int(uns[0]), int(uns[1])
        s.upload_data = _ptr(arr(ud, np.float64), C.c_double)
        s.ud_rows, s.ud_cols = int(uds[0]), int(uds[1])
        s.n_frame_map = len(fm)
        s.frame_map = _ptr(arr([-1 if v is None else int(v) for v in fm], np.int32), C.c_int32)
        s.var_params = _ptr(arr(vp, np.float64), C.c_double)
        bv = [int(v) for v in (spec.get("between_var_indices", []) or [])]
        s.n_between = len(bv)
        s.between_cols = _ptr(arr(bv, np.int32), C.c_int32)
        s.factor_n_levels = _ptr(arr(levels, np.int32), C.c_int32)
        s.factor_proportions = _ptr(arr(spec.get("factor_proportions", []), np.float64), C.c_double)
        fs = list(spec.get("factor_sampled", []) or [])
        fs = (fs + [None] * len(levels))[:len(levels)]
        s.factor_sampled = _ptr(arr([-1 if v is None else int(bool(v)) for v in fs], np.int32), C.c_int32)
        ptr, cols = [0], []
        for comp in inter:
            cols += [int(c) for c in comp]
            ptr.append(len(cols))
        s.inter_ptr = _ptr(arr(ptr, np.int32), C.c_int32)
        s.inter_cols = _ptr(arr(cols, np.int32), C.c_int32)
        s.effect_sizes = _ptr(arr(spec["effect_sizes"], np.float64), C.c_double)
        ti = list(spec.get("target_indices", []))
        s.n_marginal_targets = len(ti)
        s.target_indices = _ptr(arr(ti, np.int32), C.c_int32)
        cp = list(spec.get("contrast_pairs", []) or [])
        s.n_contrasts = len(cp)
        s.contrast_pairs = _ptr(arr([v for pr in cp for v in pr], np.int32), C.c_int32)
        # posthoc blocks → flat pairwise contrasts in canonical order (build_posthoc_contrasts, batch.rs:2183-2227)
        pairs, levels, ncs = [], [], []
        for blk in list(spec.get("posthoc", []) or []):
            d = [int(v) for v in blk["target_indices"]]
            kl = len(d) + 1
            for a in range(kl):
                for b in range(a + 1, kl):
                    pairs += [d[b - 1], d[a - 1] if a > 0 else -1]
            levels.append(kl)
            ncs.append(kl * (kl - 1) // 2)
        s.n_posthoc = len(pairs) // 2
        s.posthoc_pairs = _ptr(arr(pairs, np.int32), C.c_int32)
        s.n_posthoc_blocks = len(levels)
        s.posthoc_block_levels = _ptr(arr(levels, np.int32), C.c_int32)
        s.posthoc_block_ncontrasts = _ptr(arr(ncs, np.int32), C.c_int32)
        fc = list(spec.get("fit_columns", []) or [])
        s.n_fit_columns = len(fc)
        s.fit_columns = _ptr(arr(fc, np.int32), C.c_int32)
        s.correction_method = CORR[spec.get("correction_method", "none")]
        cv = spec["crit_values"]
        s.alpha = float(cv["alpha"])
        pa = cv.get("posthoc_alpha")
        s.posthoc_alpha = float("nan") if pa is None else float(pa)
        drv = spec.get("heteroskedasticity_driver")
        s.hsk_driver = -1 if drv is None else int(drv)
        s.residual_dist = RES[spec.get("residual_dist", "normal")]
        s.residual_pinned = int(bool(spec.get("residual_pinned", False)))
        s.outcome_kind = OUT[spec.get("outcome_kind", "continuous")]
        s.link_probit = 1 if spec.get("link") == "probit" else 0
        s.estimator = EST[spec.get("estimator", "ols")]
        s.report_overall = int(bool(spec.get("report_overall", False)))
        s.factor_min_level_count = int(spec.get("factor_min_level_count", 0))
        cl = spec.get("cluster")
        if cl:
            s.has_cluster = 1
            sizing = cl["sizing"]
            if "FixedClusters" in sizing:
                s.sizing_fixed_size, s.sizing_value = 0, int(sizing["FixedClusters"]["n_clusters"])
            else:
                s.sizing_fixed_size, s.sizing_value = 1, int(sizing["FixedSize"]["cluster_size"])
            s.tau_squared = float(cl["tau_squared"])
            slopes = list(cl.get("slopes", []) or [])
            extras = list(cl.get("extra_groupings", []) or [])
            if len(slopes) > 2 or len(extras) > 2 or any(e.get("slopes") for e in extras):
                raise ValueError("oracle: two random slopes on the primary grouping, two intercept-only extra groupings at most")
            s.n_extra = len(extras)
            for g, e in enumerate(extras):
                rel = e["relation"]
                s.extra_nested[g] = 1 if "NestedWithin" in rel else 0
                s.extra_count[g] = int(rel["NestedWithin"]["n_per_parent"] if "NestedWithin" in rel else rel["Crossed"]["n_clusters"])
                s.extra_tau_squared[g] = float(e["tau_squared"])
            if slopes:
                cols = list(spec.get("cluster_slope_design_cols", []) or [])
                if len(cols) != len(slopes):
                    raise ValueError("invalid spec: cluster_slope_design_cols must name one design column per slope")
                s.n_slopes = len(slopes)
                for a, sl in enumerate(slopes):
                    s.slope_col[a] = int(cols[a])
                    s.slope_variance[a] = float(sl["variance"])
                    s.slope_corr[a] = float(sl["corr_with_intercept"])
                if len(slopes) == 2:
                    cw = list(slopes[1].get("corr_with", []) or [])
                    s.slope_corr12 = float(cw[0]) if cw else 0.0
        sc = spec["scenario"]
        s.heterogeneity = float(sc.get("heterogeneity", 0.0))
        s.hsk_ratio = float(sc.get("heteroskedasticity_ratio", 1.0))
        s.corr_noise_sd = float(sc.get("correlation_noise_sd", 0.0))
        s.dist_change_prob = float(sc.get("distribution_change_prob", 0.0))
        nd = [DIST[v] for v in sc.get("new_distributions", [])]
        s.n_new_dists = len(nd)
        s.new_dists = _ptr(arr(nd, np.int32), C.c_int32)
        s.resid_change_prob = float(sc.get("residual_change_prob", 0.0))
        rd = [RES[v] for v in sc.get("residual_dists", [])]
        s.n_resid_dists = len(rd)
        s.resid_dists = _ptr(arr(rd, np.int32), C.c_int32)
        s.residual_df = float(sc.get("residual_df", 0.0))
        s.sampled_factor_proportions = int(bool(sc.get("sampled_factor_proportions", False)))
        lme = sc.get("lme")
        if lme:
            s.has_lme = 1
            s.re_dist = RES[lme["random_effect_dist"]]
            s.re_df = float(lme["random_effect_df"])
            s.icc_noise_sd = float(lme["icc_noise_sd"])
        self.c = s
        self.P = 1 + k + s.n_factor_dummies + s.n_interactions
        self.T = s.n_marginal_targets + s.n_contrasts
        self.n_factors = s.n_factors


def _check(rc):
    if rc != 0:
        raise OracleError(rc, lib().orc_last_error().decode())


def splitmix64(z: int) -> int:
    return int(lib().orc_splitmix64(z & 0xFFFFFFFFFFFFFFFF))


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return [int(v) for v in o]


def seq_stream(base_seed, sim_id, kind, n):
    out = np.zeros(n, dtype=np.uint32)
    lib().orc_seq_stream(base_seed, sim_id, {"normal": 0, "uniform": 1}[kind], n, _ptr(out, C.c_uint32))
    return out


def seq_categorical(base_seed, sim_id, probs, n):
    p = np.asarray(probs, dtype=np.float64)
    out = np.zeros(n, dtype=np.int32)
    lib().orc_seq_categorical(base_seed, sim_id, _ptr(p, C.c_double), len(p), n, _ptr(out, C.c_int32))
    return out


def planar_normals(base_seed, sim_id, cls, col, n):
    out = np.zeros(n, dtype=np.uint32)
    lib().orc_planar_normals(base_seed, sim_id, cls, col, n, _ptr(out, C.c_uint32))
    return out.view(np.float32)


def sample_t(base_seed, sim_id, df, n):
    out = np.zeros(n, dtype=np.float64)
    lib().orc_sample_t(base_seed, sim_id, df, n, _ptr(out, C.c_double))
    return out


def draw_residual(base_seed, sim_id, dist: str, df, n):
    out = np.zeros(n, dtype=np.float64)
    lib().orc_draw_residual(base_seed, sim_id, RES[dist], df, n, _ptr(out, C.c_double))
    return out


def crit_table(spec: dict, sample_sizes, n_predictors=None):
    ps = PackedSpec(spec)
    ss = np.asarray(sample_sizes, dtype=np.uint32)
    n = len(ss)
    P = ps.P if n_predictors is None else n_predictors
    t_unc = np.zeros(n)
    corr = np.zeros((n, max(ps.T, 1)))
    post = np.zeros(n)
    ov = np.zeros(n)
    _check(lib().orc_build_crit_table(C.byref(ps.c), _ptr(ss, C.c_uint32), n, P, _ptr(t_unc, C.c_double),
                                      _ptr(corr, C.c_double), _ptr(post, C.c_double), _ptr(ov, C.c_double)))
    return {"t_crit_sq_uncorrected": t_unc, "correction_t_crit_sq": corr[:, :ps.T], "posthoc_t_crit_sq": post,
            "overall_crit": ov}


def apply_correction(method: str, t_sq, crit_unc, crit_sq):
    t = np.ascontiguousarray(t_sq, dtype=np.float64)
    c = np.ascontiguousarray(crit_sq, dtype=np.float64)
    out = np.zeros(len(t), dtype=np.uint8)
    lib().orc_apply_correction(CORR[method], _ptr(t, C.c_double), len(t), float(crit_unc), _ptr(c, C.c_double),
                               _ptr(out, C.c_uint8))
    return out


def generate_sim(spec: dict, base_seed: int, sim_id: int, max_n: int):
    """→ (X [max_n, P] float32 (column-major storage), y [max_n] float32, cluster_ids)."""
    ps = PackedSpec(spec)
    x = np.zeros((ps.P, max_n), dtype=np.float32)
    y = np.zeros(max_n, dtype=np.float32)
    ids = np.zeros(max_n, dtype=np.int32)
    _check(lib().orc_generate_sim(C.byref(ps.c), base_seed, sim_id, max_n, _ptr(x, C.c_float), _ptr(y, C.c_float),
                                  _ptr(ids, C.c_int32)))
    return x.T, y, ids


def design_factor(spec: dict, base_seed: int, sim_id: int) -> np.ndarray:
    """Lower Cholesky factor L [k, k] of the study's (perturbed, PSD-repaired) correlation (populate_design)."""
    ps = PackedSpec(spec)
    k = int(spec["n_non_factor"])
    out = np.zeros((k, k))
    _check(lib().orc_design_factor(C.byref(ps.c), base_seed & 0xFFFFFFFFFFFFFFFF, sim_id, _ptr(out, C.c_double)))
    return out.T.copy()  # column-major in C


def fit_ols(x, y):
    """x [n, p] float64 → dict(betas, se, t_sq, sigma_sq, sst, rss, converged)."""
    x = np.asarray(x, dtype=np.float64)
    n, p = x.shape
    xc = np.ascontiguousarray(x.T)
    yv = np.ascontiguousarray(y, dtype=np.float64)
    b, se, t2 = np.zeros(p), np.zeros(p), np.zeros(p)
    s2, sst, rss = C.c_double(), C.c_double(), C.c_double()
    conv = C.c_int()
    _check(lib().orc_fit_ols(_ptr(xc, C.c_double), _ptr(yv, C.c_double), n, p, _ptr(b, C.c_double),
                             _ptr(se, C.c_double), _ptr(t2, C.c_double), C.byref(s2), C.byref(sst), C.byref(rss),
                             C.byref(conv)))
    return {"betas": b, "se": se, "t_sq": t2, "sigma_sq": s2.value, "sst": sst.value, "rss": rss.value,
            "converged": bool(conv.value)}


def fit_glm(x, y, family: str, beta_start=None):
    x = np.asarray(x, dtype=np.float64)
    n, p = x.shape
    xc = np.ascontiguousarray(x.T)
    yv = np.ascontiguousarray(y, dtype=np.float64)
    b, se = np.zeros(p), np.zeros(p)
    dev, dn = C.c_double(), C.c_double()
    conv, it = C.c_int(), C.c_int()
    bs = None if beta_start is None else np.ascontiguousarray(beta_start, dtype=np.float64)
    _check(lib().orc_fit_glm(_ptr(xc, C.c_double), _ptr(yv, C.c_double), n, p,
                             {"logit": 0, "probit": 1, "poisson": 2}[family],
                             None if bs is None else _ptr(bs, C.c_double), _ptr(b, C.c_double), _ptr(se, C.c_double),
                             C.byref(dev), C.byref(dn), C.byref(conv), C.byref(it)))
    return {"betas": b, "se": se, "deviance": dev.value, "deviance_null": dn.value, "converged": bool(conv.value),
            "iters": it.value}


def fit_lmm(x, y, ids):
    """Random-intercept profiled-REML fit on supplied data → dict(betas, se, theta, sigma_sq, converged, boundary)."""
    x = np.asarray(x, dtype=np.float64)
    n, p = x.shape
    xc = np.ascontiguousarray(x.T)
    yv = np.ascontiguousarray(y, dtype=np.float64)
    idv = np.ascontiguousarray(ids, dtype=np.int32)
    b, se = np.zeros(p), np.zeros(p)
    th, s2 = C.c_double(), C.c_double()
    conv, bd = C.c_int(), C.c_int()
    _check(lib().orc_fit_lmm(_ptr(xc, C.c_double), _ptr(yv, C.c_double), _ptr(idv, C.c_int32), n, p, _ptr(b, C.c_double),
                             _ptr(se, C.c_double), C.byref(th), C.byref(s2), C.byref(conv), C.byref(bd)))
    return {"betas": b, "se": se, "theta": th.value, "sigma_sq": s2.value, "converged": bool(conv.value),
            "boundary": bd.value}


def fit_lmm_slopes(x, y, ids, slope_cols):
    """Random intercept + one or two random slopes (design columns `slope_cols`) by profiled REML on supplied data →
    dict(betas, se, variances [intercept, slopes...], corr [(1,0), (2,0), (2,1)], sigma_sq, converged, boundary, n_eval)."""
    x = np.asarray(x, dtype=np.float64)
    n, p = x.shape
    q = 1 + len(slope_cols)
    xc = np.ascontiguousarray(x.T)
    yv = np.ascontiguousarray(y, dtype=np.float64)
    idv = np.ascontiguousarray(ids, dtype=np.int32)
    cols = np.ascontiguousarray(slope_cols, dtype=np.int32)
    b, se, vc = np.zeros(p), np.zeros(p), np.zeros(q + q * (q - 1) // 2 + 1)
    conv, bd, ne = C.c_int(), C.c_int(), C.c_int()
    _check(lib().orc_fit_lmm_slopes(_ptr(xc, C.c_double), _ptr(yv, C.c_double), _ptr(idv, C.c_int32), n, p, len(slope_cols),
                                    _ptr(cols, C.c_int32), _ptr(b, C.c_double), _ptr(se, C.c_double), _ptr(vc, C.c_double),
                                    C.byref(conv), C.byref(bd), C.byref(ne)))
    return {"betas": b, "se": se, "variances": vc[:q].copy(), "corr": vc[q:-1].copy(), "sigma_sq": vc[-1],
            "converged": bool(conv.value), "boundary": bd.value, "n_eval": ne.value}


def extra_level_of_row(spec: dict, g: int, i: int) -> int:
    """ClusterSpec::extra_level_of_row (generation.rs:386-430)."""
    s = PackedSpec(spec)
    return int(lib().orc_extra_level_of_row(C.byref(s.c), g, i))


def extra_levels_at(spec: dict, g: int, n: int) -> int:
    s = PackedSpec(spec)
    return int(lib().orc_extra_levels_at(C.byref(s.c), g, n))


def fit_lmm_general(x, y, ids, slope_cols=(), extra_ids=()):
    """Primary grouping `ids` (intercept + slopes on design columns `slope_cols`) plus intercept-only extra groupings (`extra_ids`:
    one 0-based level vector each) by profiled REML on supplied data → dict(betas, se, variances [primary intercept, slopes...],
    corr, extra_variances, sigma_sq, converged, boundary, n_eval)."""
    x = np.asarray(x, dtype=np.float64)
    n, p = x.shape
    q = 1 + len(slope_cols)
    ne = len(extra_ids)
    xc = np.ascontiguousarray(x.T)
    yv = np.ascontiguousarray(y, dtype=np.float64)
    idv = np.ascontiguousarray(np.concatenate([np.asarray(ids, dtype=np.int32)] + [np.asarray(e, dtype=np.int32) for e in extra_ids]))
    cols = np.ascontiguousarray(list(slope_cols) + [0], dtype=np.int32)
    nc = q * (q - 1) // 2
    b, se, vc = np.zeros(p), np.zeros(p), np.zeros(q + nc + ne + 1)
    conv, bd, nev = C.c_int(), C.c_int(), C.c_int()
    _check(lib().orc_fit_lmm_general(_ptr(xc, C.c_double), _ptr(yv, C.c_double), _ptr(idv, C.c_int32), n, p, len(slope_cols),
                                     _ptr(cols, C.c_int32), ne, _ptr(b, C.c_double), _ptr(se, C.c_double), _ptr(vc, C.c_double),
                                     C.byref(conv), C.byref(bd), C.byref(nev)))
    return {"betas": b, "se": se, "variances": vc[:q].copy(), "corr": vc[q:q + nc].copy(), "extra_variances": vc[q + nc:-1].copy(),
            "sigma_sq": vc[-1], "converged": bool(conv.value), "boundary": bd.value, "n_eval": nev.value}


def fit_lmm_slope(x, y, ids, slope_col: int):
    """Random intercept + one random slope (on design column `slope_col`) by profiled REML on supplied data →
    dict(betas, se, var_intercept, var_slope, corr, sigma_sq, converged, boundary, n_eval)."""
    x = np.asarray(x, dtype=np.float64)
    n, p = x.shape
    xc = np.ascontiguousarray(x.T)
    yv = np.ascontiguousarray(y, dtype=np.float64)
    idv = np.ascontiguousarray(ids, dtype=np.int32)
    b, se, vc = np.zeros(p), np.zeros(p), np.zeros(4)
    conv, bd, ne = C.c_int(), C.c_int(), C.c_int()
    _check(lib().orc_fit_lmm_slope(_ptr(xc, C.c_double), _ptr(yv, C.c_double), _ptr(idv, C.c_int32), n, p, int(slope_col),
                                   _ptr(b, C.c_double), _ptr(se, C.c_double), _ptr(vc, C.c_double), C.byref(conv), C.byref(bd),
                                   C.byref(ne)))
    return {"betas": b, "se": se, "var_intercept": vc[0], "var_slope": vc[1], "corr": vc[2], "sigma_sq": vc[3],
            "converged": bool(conv.value), "boundary": bd.value, "n_eval": ne.value}


def run_batch(spec: dict, sample_sizes, n_sims: int, base_seed: int, sim_begin: int = 0, n_threads: int = 1,
              capture: bool = False):
    """CPU restatement of engine_core::run_batch (OLS arm). Returns the BatchResult tensors as numpy arrays."""
    ps = PackedSpec(spec)
    ss = np.asarray(sample_sizes, dtype=np.uint32)
    n_ss, T, F = len(ss), ps.T, ps.n_factors
    unc = np.zeros((n_sims, n_ss, T), dtype=np.uint8)
    cor = np.zeros((n_sims, n_ss, T), dtype=np.uint8)
    conv = np.zeros((n_sims, n_ss), dtype=np.uint8)
    ov = np.zeros((n_sims, n_ss), dtype=np.uint8) if ps.c.report_overall else None
    fx = np.zeros((n_sims, n_ss, F), dtype=np.uint8) if F > 0 else None
    cap = np.full((n_sims, n_ss, T), np.nan) if capture else None
    mle = ps.c.estimator == EST["mle"]
    joint = np.zeros((n_sims, n_ss), dtype=np.uint8) if mle else None
    bh = np.zeros((n_sims, n_ss), dtype=np.uint8) if mle else None
    nph = int(ps.c.n_posthoc)
    phu = np.zeros((n_sims, n_ss, nph), dtype=np.uint8) if nph else None
    phc = np.zeros((n_sims, n_ss, nph), dtype=np.uint8) if nph else None
    u8 = C.c_uint8
    _check(lib().orc_run_batch(C.byref(ps.c), _ptr(ss, C.c_uint32), n_ss, n_sims, sim_begin, base_seed, n_threads,
                               _ptr(unc, u8), _ptr(cor, u8), _ptr(conv, u8),
                               None if ov is None else _ptr(ov, u8), None if fx is None else _ptr(fx, u8),
                               None if cap is None else _ptr(cap, C.c_double),
                               None if joint is None else _ptr(joint, u8), None if bh is None else _ptr(bh, u8),
                               None if phu is None else _ptr(phu, u8), None if phc is None else _ptr(phc, u8)))
    return {"uncorrected": unc, "corrected": cor, "converged": conv, "overall": ov, "factor_excluded": fx,
            "t_sq": cap, "joint": joint, "boundary_hit": bh, "posthoc_unc": phu, "posthoc_cor": phc}


def host_threads() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
