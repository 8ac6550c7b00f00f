//! `ivp-batch`: the `ivp` crate's API surface for batched solves on the GPU.
//!
//! Same names, defaults and error behaviour as `ivp` v0.5.1 (`Method`, `Options::builder()`, `Tolerance`,
//! `EventConfig`, `Status`, `Solution`), plus `solve_ivp_batch`.  The `IVP` trait cannot cross to the device
//! as a Rust closure, so a problem is either one of libivpb's built-ins or CUDA C source defining
//! `ivp_ode` / `ivp_events` / `ivp_jac` (compiled with NVRTC together with the solver kernels).
//! NOTE: written without a Rust toolchain (none in the build image) -- see rust/README.md.
use ivp_batch_sys as sys;
use std::ffi::{CStr, CString};
use std::ptr;

pub type Float = f64;

#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum Method { RK23, DOPRI5, DOP853, RK4, RADAU, BDF }          // ivp: src/solve/options.rs:14-27
impl From<&str> for Method {
    fn from(s: &str) -> Self {
        match s.to_uppercase().as_str() {
            "RK23" => Method::RK23, "DOPRI5" | "RK45" => Method::DOPRI5, "DOP853" => Method::DOP853,
            "RK4" => Method::RK4, "RADAU" | "RADAU5" => Method::RADAU, "BDF" | "BDF15" => Method::BDF,
            _ => Method::DOPRI5,
        }
    }
}

#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum Status { Success, UserInterrupt, NeedLargerNMax, StepSizeTooSmall, ProbablyStiff, SingularMatrix, PoorConvergence }
impl Status {
    fn from_code(c: i32) -> Status {                                   // ivp: src/status.rs:4-19, declaration order
        [Status::Success, Status::UserInterrupt, Status::NeedLargerNMax, Status::StepSizeTooSmall,
         Status::ProbablyStiff, Status::SingularMatrix, Status::PoorConvergence][c as usize]
    }
}

#[derive(Clone, Debug)]
pub enum Tolerance { Scalar(Float), Vector(Vec<Float>) }             // ivp: src/methods/mod.rs:104-107
impl From<Float> for Tolerance { fn from(v: Float) -> Self { Tolerance::Scalar(v) } }
impl From<Vec<Float>> for Tolerance { fn from(v: Vec<Float>) -> Self { Tolerance::Vector(v) } }
impl<const N: usize> From<[Float; N]> for Tolerance { fn from(v: [Float; N]) -> Self { Tolerance::Vector(v.to_vec()) } }
impl Tolerance {
    fn as_slice(&self) -> &[Float] { match self { Tolerance::Scalar(v) => std::slice::from_ref(v), Tolerance::Vector(v) => v } }
}

#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum Direction { All, Positive, Negative }                       // ivp: src/solve/event.rs:56-77
#[derive(Clone, Copy, Debug)]
pub struct EventConfig { pub direction: Direction, pub terminal_count: Option<usize> }
impl Default for EventConfig { fn default() -> Self { Self { direction: Direction::All, terminal_count: None } } }
impl EventConfig {
    pub fn new() -> Self { Self::default() }
    pub fn terminal_count(&mut self, n: usize) { self.terminal_count = Some(n); }
    pub fn terminal(&mut self) { self.terminal_count = Some(1); }
    pub fn all(&mut self) { self.direction = Direction::All; }
    pub fn positive(&mut self) { self.direction = Direction::Positive; }
    pub fn negative(&mut self) { self.direction = Direction::Negative; }
}

#[derive(Debug)]
pub enum Error { Config(String), Device(String) }                    // ivp: src/error.rs (Config); Device is new

/// `Options` of ivp (src/solve/options.rs:75-123) plus the batch-only knobs at the end.
#[derive(Clone, Debug)]
pub struct Options {
    pub method: Method, pub rtol: Tolerance, pub atol: Tolerance, pub max_steps: Option<usize>,
    pub t_eval: Option<Vec<Float>>, pub first_step: Option<Float>, pub max_step: Option<Float>,
    pub min_step: Option<Float>, pub dense_output: bool,
    pub event_config: Option<Vec<EventConfig>>, pub max_events: usize, pub max_out: usize,
    pub analytic_jac: bool, pub strict_fp: bool, pub max_segments: usize,
    /// ivp `Options.mass_storage` (Identity | Full) and `nind1..3` (src/solve/options.rs:105-122); RADAU only.
    /// the problem's own SolOut hook replaces DefaultSolOut (ivp: `Method::solve(.., Some(&mut solout))`, src/solout.rs:55-63)
    pub user_solout: bool,
    pub mass_full: bool, pub nind1: Option<usize>, pub nind2: Option<usize>, pub nind3: Option<usize>,
    /// `jac_sparsity` of ivp's Python front end (src/python/sparsity.rs): structural non-zeros as (row, col) pairs.
    pub jac_sparsity: Option<Vec<(usize, usize)>>,
}
impl Options { pub fn builder() -> OptionsBuilder { OptionsBuilder(Options::default()) } }
impl Default for Options {
    fn default() -> Self {
        Options { method: Method::DOPRI5, rtol: 1e-3.into(), atol: 1e-6.into(), max_steps: None, t_eval: None,
                  first_step: None, max_step: None, min_step: None, dense_output: false, event_config: None,
                  max_events: 8, max_out: 4096, analytic_jac: false, strict_fp: false, max_segments: 4096,
                  user_solout: false, mass_full: false, nind1: None, nind2: None, nind3: None, jac_sparsity: None }
    }
}
pub struct OptionsBuilder(Options);
impl OptionsBuilder {
    pub fn method(mut self, m: impl Into<Method>) -> Self { self.0.method = m.into(); self }
    pub fn rtol(mut self, t: impl Into<Tolerance>) -> Self { self.0.rtol = t.into(); self }
    pub fn atol(mut self, t: impl Into<Tolerance>) -> Self { self.0.atol = t.into(); self }
    pub fn max_steps(mut self, n: usize) -> Self { self.0.max_steps = Some(n); self }
    pub fn t_eval(mut self, t: Vec<Float>) -> Self { self.0.t_eval = Some(t); self }
    pub fn first_step(mut self, h: Float) -> Self { self.0.first_step = Some(h); self }
    pub fn max_step(mut self, h: Float) -> Self { self.0.max_step = Some(h); self }
    pub fn min_step(mut self, h: Float) -> Self { self.0.min_step = Some(h); self }
    pub fn dense_output(mut self, b: bool) -> Self { self.0.dense_output = b; self }
    pub fn event_config(mut self, c: Vec<EventConfig>) -> Self { self.0.event_config = Some(c); self }
    pub fn max_events(mut self, n: usize) -> Self { self.0.max_events = n; self }
    pub fn max_out(mut self, n: usize) -> Self { self.0.max_out = n; self }
    pub fn analytic_jac(mut self, b: bool) -> Self { self.0.analytic_jac = b; self }
    pub fn strict_fp(mut self, b: bool) -> Self { self.0.strict_fp = b; self }
    pub fn max_segments(mut self, n: usize) -> Self { self.0.max_segments = n; self }
    /// ivp `Options.mass_storage(MatrixStorage::Full)`: RADAU integrates `M y' = f` with the problem's `ivp_mass`.
    pub fn mass_full(mut self, b: bool) -> Self { self.0.mass_full = b; self }
    pub fn nind1(mut self, k: usize) -> Self { self.0.nind1 = Some(k); self }
    pub fn nind2(mut self, k: usize) -> Self { self.0.nind2 = Some(k); self }
    pub fn nind3(mut self, k: usize) -> Self { self.0.nind3 = Some(k); self }
    /// Sparse finite-difference Jacobian for RADAU / BDF: one RHS call per group of structurally orthogonal columns.
    pub fn jac_sparsity(mut self, nz: Vec<(usize, usize)>) -> Self { self.0.jac_sparsity = Some(nz); self }
    /// The problem's own `SolOut` (`ivp_solout`) instead of `DefaultSolOut` (ivp: `Method::solve(.., Some(&mut solout))`).
    pub fn user_solout(mut self, b: bool) -> Self { self.0.user_solout = b; self }
    pub fn build(self) -> Options { self.0 }
}

/// ivp: src/solve/solution.rs:7-20 (continuous_sol: see DESIGN.md, "next")
#[derive(Clone, Debug)]
pub struct Solution {
    pub t: Vec<Float>, pub y: Vec<Vec<Float>>, pub t_events: Vec<Vec<Float>>, pub y_events: Vec<Vec<Vec<Float>>>,
    pub nfev: usize, pub njev: usize, pub nlu: usize, pub nstep: usize, pub naccpt: usize, pub nrejct: usize,
    pub status: Status, pub h_next: Float, pub truncated: bool,
}

pub struct Context { raw: *mut sys::ivpb_ctx }
impl Context {
    pub fn new(devices: &[i32]) -> Result<Self, Error> {
        let mut raw = ptr::null_mut();
        let rc = unsafe { sys::ivpb_create(&mut raw, if devices.is_empty() { ptr::null() } else { devices.as_ptr() }, devices.len() as i32) };
        if rc != sys::IVPB_OK { return Err(Error::Device(last_error(ptr::null()))); }
        Ok(Context { raw })
    }
}
impl Drop for Context { fn drop(&mut self) { unsafe { sys::ivpb_destroy(self.raw) } } }
fn last_error(ctx: *const sys::ivpb_ctx) -> String {
    unsafe { CStr::from_ptr(sys::ivpb_last_error(ctx)).to_string_lossy().into_owned() }
}

/// Device form of `trait IVP` (ivp: src/ivp.rs:27-121).
pub struct Problem { handle: i32, pub n: usize, pub p: usize, pub n_events: usize }
impl Problem {
    pub fn builtin(ctx: &Context, id: i32) -> Result<Self, Error> {
        let (mut n, mut p, mut ne) = (0, 0, 0);
        if unsafe { sys::ivpb_builtin_problem(ctx.raw, id, &mut n, &mut p, &mut ne) } != sys::IVPB_OK { return Err(Error::Config(last_error(ctx.raw))); }
        Ok(Problem { handle: id, n: n as usize, p: p as usize, n_events: ne as usize })
    }
    pub fn from_cuda_source(ctx: &Context, src: &str, n: usize, p: usize, n_events: usize, has_jac: bool, has_mass: bool,
                            has_solout: bool) -> Result<Self, Error> {
        let c = CString::new(src).map_err(|e| Error::Config(e.to_string()))?;
        let mut h = -1;
        let rc = unsafe { sys::ivpb_nvrtc_problem(ctx.raw, c.as_ptr(), n as i32, p as i32, n_events as i32, (has_jac as i32) | ((has_mass as i32) << 1) | ((has_solout as i32) << 2), &mut h) };
        if rc != sys::IVPB_OK { return Err(Error::Config(last_error(ctx.raw))); }
        Ok(Problem { handle: h, n, p, n_events })
    }
}

/// `Solution::sol_many` for trajectory `index` of the last dense_output solve (evaluated on the device).
pub fn sol_many(ctx: &Context, index: usize, n: usize, ts: &[Float]) -> Result<Vec<Vec<Float>>, Error> {
    let traj = vec![index as i64; ts.len()];
    let mut y = vec![0.0; ts.len() * n]; let mut ok = vec![0i32; ts.len()];
    let rc = unsafe { sys::ivpb_dense_eval(ctx.raw, 0, n as i32, ts.len() as i64, traj.as_ptr(), ts.as_ptr(), y.as_mut_ptr(), ok.as_mut_ptr()) };
    if rc != sys::IVPB_OK || ok.iter().any(|&k| k == 0) { return Err(Error::Config("t outside the dense output span (InterpolationError)".into())); }
    Ok(y.chunks(n).map(|c| c.to_vec()).collect())
}

/// `ContinuousOutput::evaluate_extrapolate` (reference src/solve/cont.rs:91-150) for trajectory `index`.
pub fn sol_extrapolate(ctx: &Context, index: usize, n: usize, t: Float) -> Option<Vec<Float>> {
    let (traj, mut y, mut ok) = (index as i64, vec![0.0; n], 0i32);
    let rc = unsafe { sys::ivpb_dense_eval_extrapolate(ctx.raw, 0, n as i32, 1, &traj, &t, y.as_mut_ptr(), &mut ok) };
    if rc != sys::IVPB_OK || ok == 0 { None } else { Some(y) }
}

/// `y0` is `[N x n]` row-major, `params` `[N x p]` row-major.  One `Solution` per trajectory.
pub fn solve_ivp_batch(ctx: &Context, f: &Problem, t0: Float, tf: Float, y0: &[Float], params: &[Float], options: Options)
                       -> Result<Vec<Solution>, Error> {
    let (n, ne) = (f.n, f.n_events);
    if n == 0 || y0.len() % n != 0 { return Err(Error::Config("y0 must hold N x n values".into())); }
    let big_n = y0.len() / n;
    if params.len() != big_n * f.p { return Err(Error::Config("params must hold N x p values".into())); }
    for tol in [&options.rtol, &options.atol] {
        if let Tolerance::Vector(v) = tol { if v.len() != n { return Err(Error::Config("tolerance vector length != n".into())); } }
    }
    let (dirs, terms): (Vec<i32>, Vec<i64>) = options.event_config.as_ref().map(|c| c.iter().map(|e| (
        match e.direction { Direction::All => 0, Direction::Positive => 1, Direction::Negative => -1 },
        e.terminal_count.map(|k| k as i64).unwrap_or(-1))).unzip()).unwrap_or_default();
    let te = options.t_eval.as_deref();
    // (row, col) pairs -> compressed columns (rows ascending, duplicates dropped)
    let (sp_colptr, sp_rows): (Vec<i32>, Vec<i32>) = match &options.jac_sparsity {
        None => (Vec::new(), Vec::new()),
        Some(nz) => {
            let mut cols: Vec<Vec<i32>> = vec![Vec::new(); n];
            for &(r, c) in nz { assert!(r < n && c < n, "jac_sparsity entry out of range"); cols[c].push(r as i32); }
            let mut colptr = vec![0i32]; let mut rows = Vec::new();
            for c in cols.iter_mut() { c.sort_unstable(); c.dedup(); rows.extend_from_slice(c); colptr.push(rows.len() as i32); }
            if rows.is_empty() { rows.push(0); }
            (colptr, rows)
        }
    };
    let o = sys::ivpb_options {
        method: options.method as i32,
        n_rtol: options.rtol.as_slice().len() as i32, n_atol: options.atol.as_slice().len() as i32,
        rtol: options.rtol.as_slice().as_ptr(), atol: options.atol.as_slice().as_ptr(),
        has_first_step: options.first_step.is_some() as i32, has_max_step: options.max_step.is_some() as i32,
        has_min_step: options.min_step.is_some() as i32, has_max_steps: options.max_steps.is_some() as i32,
        first_step: options.first_step.unwrap_or(0.0), max_step: options.max_step.unwrap_or(0.0),
        min_step: options.min_step.unwrap_or(0.0), max_steps: options.max_steps.unwrap_or(0) as u64,
        has_t_eval: te.is_some() as i32, n_t_eval: te.map_or(0, |t| t.len()) as i32,
        t_eval: te.map_or(ptr::null(), |t| t.as_ptr()), dense_output: options.dense_output as i32,
        n_event_cfg: dirs.len() as i32, ev_direction: if dirs.is_empty() { ptr::null() } else { dirs.as_ptr() },
        ev_terminal_count: if terms.is_empty() { ptr::null() } else { terms.as_ptr() },
        max_events: if ne > 0 { options.max_events as i32 } else { 0 }, max_out: options.max_out as i32,
        jac_mode: options.analytic_jac as i32, flags: if options.strict_fp { sys::IVPB_FLAG_STRICT_FP } else { 0 },
        max_segments: if options.dense_output { options.max_segments as i32 } else { 0 },
        mass_storage: options.mass_full as i32, user_solout: options.user_solout as i32,
        nind1: options.nind1.map_or(-1, |k| k as i32), nind2: options.nind2.map_or(-1, |k| k as i32),
        nind3: options.nind3.map_or(-1, |k| k as i32),
        has_jac_sparsity: options.jac_sparsity.is_some() as i32,
        jac_sparsity_colptr: if sp_colptr.is_empty() { ptr::null() } else { sp_colptr.as_ptr() },
        jac_sparsity_rows: if sp_rows.is_empty() { ptr::null() } else { sp_rows.as_ptr() },
    };
    let cap = te.map_or(options.max_out, |t| t.len() + 1);
    let me = o.max_events as usize;
    let mut status = vec![0i32; big_n]; let mut n_out = vec![0i32; big_n]; let mut ev_count = vec![0i32; big_n * ne];
    let mut counters = vec![0u32; big_n * 6];
    let mut t_final = vec![0.0; big_n]; let mut y_final = vec![0.0; big_n * n]; let mut h_next = vec![0.0; big_n];
    let mut t_out = vec![0.0; big_n * cap]; let mut y_out = vec![0.0; big_n * cap * n];
    let mut ev_t = vec![0.0; big_n * ne * me]; let mut ev_y = vec![0.0; big_n * ne * me * n];
    let out = sys::ivpb_outputs {
        status: status.as_mut_ptr(), counters: counters.as_mut_ptr(), t_final: t_final.as_mut_ptr(),
        y_final: y_final.as_mut_ptr(), h_next: h_next.as_mut_ptr(), n_out: n_out.as_mut_ptr(),
        t_out: if cap > 0 { t_out.as_mut_ptr() } else { ptr::null_mut() }, y_out: if cap > 0 { y_out.as_mut_ptr() } else { ptr::null_mut() },
        ev_count: if ne > 0 { ev_count.as_mut_ptr() } else { ptr::null_mut() },
        ev_t: if ne > 0 { ev_t.as_mut_ptr() } else { ptr::null_mut() }, ev_y: if ne > 0 { ev_y.as_mut_ptr() } else { ptr::null_mut() },
        n_seg: ptr::null_mut(), seg_x: ptr::null_mut(), seg_cont: ptr::null_mut(),   // the dense log stays on the device
    };
    let rc = unsafe { sys::ivpb_solve_batch(ctx.raw, f.handle, &o, big_n as i64, t0, tf, y0.as_ptr(),
                                            if f.p > 0 { params.as_ptr() } else { ptr::null() }, &out) };
    if rc == sys::IVPB_ERR_CONFIG { return Err(Error::Config(last_error(ctx.raw))); }
    if rc != sys::IVPB_OK { return Err(Error::Device(last_error(ctx.raw))); }
    Ok((0..big_n).map(|i| {
        let m = (n_out[i].max(0) as usize).min(cap);
        let c = &counters[6 * i..6 * i + 6];
        let mut truncated = n_out[i].max(0) as usize > cap;
        let (mut t, mut y): (Vec<Float>, Vec<Vec<Float>>) = (t_out[i * cap..i * cap + m].to_vec(),
            (0..m).map(|k| y_out[(i * cap + k) * n..(i * cap + k + 1) * n].to_vec()).collect());
        if cap == 0 { t = vec![t_final[i]]; y = vec![y_final[i * n..(i + 1) * n].to_vec()]; }
        let mut t_events = Vec::with_capacity(ne); let mut y_events = Vec::with_capacity(ne);
        for e in 0..ne {
            let hits = ev_count[i * ne + e] as usize; let keep = hits.min(me); truncated |= hits > keep;
            let base = (i * ne + e) * me;
            t_events.push(ev_t[base..base + keep].to_vec());
            y_events.push((0..keep).map(|k| ev_y[(base + k) * n..(base + k + 1) * n].to_vec()).collect());
        }
        Solution { t, y, t_events, y_events, nfev: c[0] as usize, njev: c[1] as usize, nlu: c[2] as usize,
                   nstep: c[3] as usize, naccpt: c[4] as usize, nrejct: c[5] as usize,
                   status: Status::from_code(status[i]), h_next: h_next[i], truncated }
    }).collect())
}
