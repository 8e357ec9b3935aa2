//! `engine_core::run_batch` on B200 GPUs: a thin FFI shim over `libmcpower_b200.so` (C ABI: `include/mcpower_b200.h`).
//!
//! Seam: `crates/engine-core/src/batch.rs:130-146`.  `engine_core::run_batch` selects this backend behind a feature flag
//! (`#[cfg(feature = "cuda")] return engine_cuda::run_batch(spec, sample_sizes, n_sims, base_seed, sink);`) or after a
//! runtime probe of [`device_count`]; nothing else in the workspace changes.
//!
//! * the spec crosses the boundary as the serde JSON of `SimulationSpec` (the derive the wasm port already round-trips,
//!   `crates/engine-wasm/src/lib.rs:203-230`);
//! * `ProgressSink::report` / `is_cancelled` are served from the calling thread while the kernels run — the library never
//!   splits the batch for them;
//! * `n_devices = -1` shards the sims of one call over every visible GPU (same `base_seed`, contiguous global sim ids:
//!   the result is bit-identical to one device), so rayon's role (batch.rs:637-742) is played inside the library;
//! * errors map one to one onto `EngineError` (spec.rs:19-29).
#![allow(clippy::missing_safety_doc)]

use engine_core::spec::{BatchResult, EngineError, PosthocBlockShape, ProgressSink, ResultShape, SimulationSpec};
use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};

pub const MODE_COUNTS: u32 = 1;
pub const MODE_BYTES: u32 = 2;
pub const MODE_CAPTURE: u32 = 4;

type ProgressFn = extern "C" fn(u64, u64, *mut c_void) -> c_int;
type CancelFn = extern "C" fn(*mut c_void) -> c_int;

/// `McpOptions` (include/mcpower_b200.h).
#[repr(C)]
pub struct McpOptions {
    pub device: i32,
    pub mode: u32,
    pub sim_begin: u64,
    pub wave_sims: u32,
    pub n_devices: i32,
    pub is_cancelled: Option<CancelFn>,
}

/// `McpBatchResult` (include/mcpower_b200.h).  Buffers are owned by the library until `mcp_free_result`.
#[repr(C)]
pub struct McpBatchResult {
    pub n_sims: u32,
    pub n_sample_sizes: u32,
    pub n_targets: u32,
    pub n_factors: u32,
    pub uncorrected: *mut u8,
    pub corrected: *mut u8,
    pub converged: *mut u8,
    pub overall: *mut u8,
    pub factor_excluded: *mut u8,
    pub joint: *mut u8,
    pub boundary_hit: *mut u8,
    pub t_sq: *mut f64,
    pub counter_width: u32,
    pub gpu_launches: u32,
    pub counters: *mut u64,
    pub kernel_ms: f64,
    pub n_posthoc: u32,
    pub reserved: u32,
    pub posthoc_unc: *mut u8,
    pub posthoc_cor: *mut u8,
    pub h2d_bytes: u64,
    pub d2h_bytes: u64,
    pub n_devices: u32,
    pub n_waves: u32,
    pub lower_ms: f64,
    pub enqueue_ms: f64,
    pub wait_ms: f64,
}

extern "C" {
    fn mcp_run_batch(
        spec_json: *const c_char,
        spec_len: usize,
        sample_sizes: *const u32,
        n_sample_sizes: u32,
        n_sims: u32,
        base_seed: u64,
        options: *const McpOptions,
        progress: Option<ProgressFn>,
        user: *mut c_void,
        out: *mut *mut McpBatchResult,
    ) -> c_int;
    fn mcp_free_result(result: *mut McpBatchResult);
    fn mcp_last_error() -> *const c_char;
    fn mcp_device_count() -> c_int;
    fn mcp_find_power_contracts(
        contracts: *const u8,
        len: usize,
        format: i32,
        sample_size: u64,
        n_sims: u32,
        user_seed: u64,
        options: *const McpOptions,
        progress: Option<ProgressFn>,
        user: *mut c_void,
        json_out: *mut *mut c_char,
    ) -> c_int;
    fn mcp_find_sample_size_contracts(
        contracts: *const u8,
        len: usize,
        format: i32,
        from_size: u64,
        to_size: u64,
        by: u64,
        by_auto: i32,
        log_grid: i32,
        target_power: f64,
        n_sims: u32,
        user_seed: u64,
        options: *const McpOptions,
        progress: Option<ProgressFn>,
        user: *mut c_void,
        json_out: *mut *mut c_char,
    ) -> c_int;
    fn mcp_free_string(s: *mut c_char);
    fn mcp_fit_provided_data(
        spec_json: *const c_char,
        spec_len: usize,
        design_colmajor: *const f64,
        nrow: u32,
        ncol: u32,
        outcome: *const f64,
        cluster_ids: *const u32,
        device: i32,
        betas: *mut f64,
        se: *mut f64,
        statistic: *mut f64,
        info: *mut McpFitInfo,
    ) -> c_int;
}

/// `McpFitInfo` (include/mcpower_b200.h): what `introspect::fit_provided_data` reports next to beta / SE / statistic.
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct McpFitInfo {
    pub converged: i32,
    pub boundary_hit: i32,
    pub iterations: i32,
    pub n_targets: i32,
    pub crit_sq_uncorrected: f64,
    pub df_resid: f64,
    pub sigma_sq_hat: f64,
    pub tau_sq_hat: f64,
    pub deviance: f64,
    pub deviance_null: f64,
    pub slope_var_hat: [f64; 2],
    pub slope_corr_hat: [f64; 3],
    pub extra_var_hat: [f64; 2],
    pub pinned_components: i32,
}

/// The shipped fit code of `spec.estimator` on supplied bytes (`introspect.rs:174-562`): OLS solve, IRLS from the truth start, or
/// profiled REML (random intercept, slopes on the primary grouping, crossed / nested extra groupings) on the CUDA kernels.
/// `design` is column-major `nrow x ncol` with the intercept column first; returns (betas[ncol], se[n_targets],
/// statistic[n_targets], info).
pub fn fit_provided_data(
    spec: &SimulationSpec,
    design: &[f64],
    nrow: usize,
    ncol: usize,
    outcome: &[f64],
    cluster_ids: Option<&[u32]>,
) -> Result<(Vec<f64>, Vec<f64>, Vec<f64>, McpFitInfo), EngineError> {
    let json = serde_json::to_vec(spec).map_err(|e| EngineError::InvalidSpec(e.to_string()))?;
    if design.len() != nrow * ncol || outcome.len() != nrow {
        return Err(EngineError::InvalidSpec("design / outcome shape".into()));
    }
    let nt = spec.target_indices.len();
    let (mut betas, mut se, mut stat) = (vec![0.0; ncol], vec![0.0; nt], vec![0.0; nt]);
    let mut info = McpFitInfo::default();
    let rc = unsafe {
        mcp_fit_provided_data(
            json.as_ptr() as *const c_char,
            json.len(),
            design.as_ptr(),
            nrow as u32,
            ncol as u32,
            outcome.as_ptr(),
            cluster_ids.map_or(std::ptr::null(), |c| c.as_ptr()),
            0,
            betas.as_mut_ptr(),
            se.as_mut_ptr(),
            stat.as_mut_ptr(),
            &mut info,
        )
    };
    if rc != 0 {
        return Err(status_to_error(rc));
    }
    Ok((betas, se, stat, info))
}

/// Visible CUDA devices; 0 = fall back to the CPU engine.
pub fn device_count() -> usize {
    unsafe { mcp_device_count().max(0) as usize }
}

fn last_error() -> String {
    unsafe { CStr::from_ptr(mcp_last_error()).to_string_lossy().into_owned() }
}

fn status_to_error(rc: c_int) -> EngineError {
    match rc {
        1 => EngineError::InvalidSpec(last_error()),
        2 => EngineError::CorrelationNotPSD,
        3 => EngineError::RankDeficient(0),
        4 => EngineError::Cancelled,
        // 5 = CUDA error, 6 = outside the CUDA path's scope: surfaced, never papered over by a CPU run
        _ => EngineError::InvalidSpec(format!("cuda backend: {}", last_error())),
    }
}

/// `user` points at a `&dyn ProgressSink` living on the caller's stack for the duration of the call.
extern "C" fn report_trampoline(current: u64, total: u64, user: *mut c_void) -> c_int {
    let sink: &&dyn ProgressSink = unsafe { &*(user as *const &dyn ProgressSink) };
    sink.report(current, total) as c_int
}

extern "C" fn cancel_trampoline(user: *mut c_void) -> c_int {
    let sink: &&dyn ProgressSink = unsafe { &*(user as *const &dyn ProgressSink) };
    sink.is_cancelled() as c_int
}

unsafe fn take<T: Copy>(p: *const T, n: usize) -> Vec<T> {
    if p.is_null() || n == 0 {
        Vec::new()
    } else {
        std::slice::from_raw_parts(p, n).to_vec()
    }
}

/// Posthoc block shapes as `run_batch` derives them (batch.rs:2314-2326): one block per `spec.posthoc` entry, k levels =
/// dummies + 1, C(k, 2) contrasts.
fn posthoc_block_shapes(spec: &SimulationSpec) -> Vec<PosthocBlockShape> {
    spec.posthoc
        .iter()
        .map(|b| {
            let k = b.target_indices.len() as u32 + 1;
            PosthocBlockShape { n_levels: k, n_contrasts: k * (k - 1) / 2 }
        })
        .collect()
}

/// Options of the drop-in: every visible device, the sim-major byte tensors `BatchResult` carries (+ counters).
pub fn default_options() -> McpOptions {
    McpOptions { device: 0, mode: MODE_BYTES | MODE_COUNTS, sim_begin: 0, wave_sims: 0, n_devices: -1, is_cancelled: None }
}

/// Drop-in for `engine_core::run_batch` (batch.rs:130-146).
pub fn run_batch(
    spec: &SimulationSpec,
    sample_sizes: &[u32],
    n_sims: u32,
    base_seed: u64,
    sink: Option<&dyn ProgressSink>,
) -> Result<BatchResult, EngineError> {
    let json = serde_json::to_vec(spec).map_err(|e| EngineError::InvalidSpec(e.to_string()))?;
    let mut opt = default_options();
    let mut out: *mut McpBatchResult = std::ptr::null_mut();
    let (progress, user): (Option<ProgressFn>, *mut c_void) = match sink.as_ref() {
        Some(s) => {
            opt.is_cancelled = Some(cancel_trampoline);
            (Some(report_trampoline), s as *const &dyn ProgressSink as *mut c_void)
        }
        None => (None, std::ptr::null_mut()),
    };
    let rc = unsafe {
        mcp_run_batch(
            json.as_ptr() as *const c_char,
            json.len(),
            sample_sizes.as_ptr(),
            sample_sizes.len() as u32,
            n_sims,
            base_seed,
            &opt,
            progress,
            user,
            &mut out,
        )
    };
    if rc != 0 {
        return Err(status_to_error(rc));
    }
    let r = unsafe { &*out };
    let cells = r.n_sims as usize * r.n_sample_sizes as usize;
    let t = r.n_targets as usize;
    let c = r.n_posthoc as usize;
    let boundary_hit = unsafe { take(r.boundary_hit, cells) };
    let joint = unsafe { take(r.joint, cells) };
    let zeros = || vec![0u8; cells];
    let res = BatchResult {
        uncorrected: unsafe { take(r.uncorrected, cells * t) },
        corrected: unsafe { take(r.corrected, cells * t) },
        posthoc_unc: unsafe { take(r.posthoc_unc, cells * c) },
        posthoc_cor: unsafe { take(r.posthoc_cor, cells * c) },
        converged: unsafe { take(r.converged, cells) },
        // the library packs BatchResult::boundary_hit (bits 0-1) and ::pinned_components (bits 2-6) into one byte per fit
        pinned_components: if boundary_hit.is_empty() {
            vec![0u64; cells]
        } else {
            boundary_hit.iter().map(|&b| (b >> 2) as u64).collect()
        },
        boundary_hit: if boundary_hit.is_empty() { zeros() } else { boundary_hit.iter().map(|&b| b & 3).collect() },
        joint_unc: if joint.is_empty() { zeros() } else { joint.clone() },
        joint_cor: if joint.is_empty() { zeros() } else { joint },
        overall: unsafe { take(r.overall, cells) },
        factor_excluded: unsafe { take(r.factor_excluded, cells * r.n_factors as usize) },
        tau_squared_hat: vec![f64::NAN; cells],
        shape: ResultShape {
            n_sims: r.n_sims,
            n_sample_sizes: r.n_sample_sizes,
            n_targets: r.n_targets,
            posthoc_blocks: posthoc_block_shapes(spec),
            n_factors: r.n_factors,
            n_variance_components: spec.cluster.is_some() as u32,
        },
    };
    unsafe { mcp_free_result(out) };
    Ok(res)
}

/// The hosts' msgpack `Vec<SimulationContract>` blob (engine-py/src/orchestrator_bridge.rs:23-30) straight through the
/// library's orchestrator layer: JSON of `ScenarioResult<PowerResult>` (result.rs:581-584), counters folded on the device.
pub fn find_power_contracts_json(
    contracts_msgpack: &[u8],
    sample_size: u64,
    n_sims: u32,
    user_seed: u64,
    sink: Option<&dyn ProgressSink>,
) -> Result<String, EngineError> {
    let mut opt = default_options();
    opt.mode = MODE_COUNTS;
    let (progress, user): (Option<ProgressFn>, *mut c_void) = match sink.as_ref() {
        Some(s) => {
            opt.is_cancelled = Some(cancel_trampoline);
            (Some(report_trampoline), s as *const &dyn ProgressSink as *mut c_void)
        }
        None => (None, std::ptr::null_mut()),
    };
    let mut json: *mut c_char = std::ptr::null_mut();
    let rc = unsafe {
        mcp_find_power_contracts(
            contracts_msgpack.as_ptr(),
            contracts_msgpack.len(),
            0,
            sample_size,
            n_sims,
            user_seed,
            &opt,
            progress,
            user,
            &mut json,
        )
    };
    if rc != 0 {
        return Err(status_to_error(rc));
    }
    let s = unsafe { CStr::from_ptr(json).to_string_lossy().into_owned() };
    unsafe { mcp_free_string(json) };
    Ok(s)
}

/// `find_sample_size` twin of [`find_power_contracts_json`]: JSON of `ScenarioResult<SampleSizeResult>`.
#[allow(clippy::too_many_arguments)]
pub fn find_sample_size_contracts_json(
    contracts_msgpack: &[u8],
    from_size: u64,
    to_size: u64,
    by: Option<u64>,
    log_grid: bool,
    target_power: f64,
    n_sims: u32,
    user_seed: u64,
    sink: Option<&dyn ProgressSink>,
) -> Result<String, EngineError> {
    let mut opt = default_options();
    opt.mode = MODE_COUNTS;
    let (progress, user): (Option<ProgressFn>, *mut c_void) = match sink.as_ref() {
        Some(s) => {
            opt.is_cancelled = Some(cancel_trampoline);
            (Some(report_trampoline), s as *const &dyn ProgressSink as *mut c_void)
        }
        None => (None, std::ptr::null_mut()),
    };
    let mut json: *mut c_char = std::ptr::null_mut();
    let rc = unsafe {
        mcp_find_sample_size_contracts(
            contracts_msgpack.as_ptr(),
            contracts_msgpack.len(),
            0,
            from_size,
            to_size,
            by.unwrap_or(12),
            by.is_none() as i32,
            log_grid as i32,
            target_power,
            n_sims,
            user_seed,
            &opt,
            progress,
            user,
            &mut json,
        )
    };
    if rc != 0 {
        return Err(status_to_error(rc));
    }
    let s = unsafe { CStr::from_ptr(json).to_string_lossy().into_owned() };
    unsafe { mcp_free_string(json) };
    Ok(s)
}
