//! Raw bindings to `include/ivpb.h` (hand-written, bindgen-free).  Field order and types must match the
//! header exactly; `tests/test_abi_and_host.py::test_struct_layout_matches_header` pins the same layout for
//! the ctypes mirror.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct ivpb_ctx {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct ivpb_options {
    pub method: i32,
    pub n_rtol: i32,
    pub n_atol: i32,
    pub rtol: *const f64,
    pub atol: *const f64,
    pub has_first_step: i32,
    pub has_max_step: i32,
    pub has_min_step: i32,
    pub has_max_steps: i32,
    pub first_step: f64,
    pub max_step: f64,
    pub min_step: f64,
    pub max_steps: u64,
    pub has_t_eval: i32,
    pub n_t_eval: i32,
    pub t_eval: *const f64,
    pub dense_output: i32,
    pub n_event_cfg: i32,
    pub ev_direction: *const i32,
    pub ev_terminal_count: *const i64,
    pub max_events: i32,
    pub max_out: i32,
    pub jac_mode: i32,
    pub flags: i32,
    pub max_segments: i32,
    pub mass_storage: i32,
    pub user_solout: i32,
    pub nind1: i32, pub nind2: i32, pub nind3: i32,
    pub has_jac_sparsity: i32,
    pub jac_sparsity_colptr: *const i32,
    pub jac_sparsity_rows: *const i32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct ivpb_outputs {
    pub status: *mut i32,
    pub counters: *mut u32,
    pub t_final: *mut f64,
    pub y_final: *mut f64,
    pub h_next: *mut f64,
    pub n_out: *mut i32,
    pub t_out: *mut f64,
    pub y_out: *mut f64,
    pub ev_count: *mut i32,
    pub ev_t: *mut f64,
    pub ev_y: *mut f64,
    pub n_seg: *mut i32,
    pub seg_x: *mut f64,
    pub seg_cont: *mut f64,
}

pub const IVPB_OK: c_int = 0;
pub const IVPB_ERR_CONFIG: c_int = 1;
pub const IVPB_ERR_CUDA: c_int = 2;
pub const IVPB_ERR_NVRTC: c_int = 3;
pub const IVPB_FLAG_STRICT_FP: i32 = 1;
pub const IVPB_FLAG_NO_REFILL: i32 = 2;
pub const IVPB_FLAG_NO_ZEROCOPY: i32 = 4;
pub const IVPB_FLAG_FAST_FP: i32 = 8;
pub const IVPB_FLAG_NO_SORT: i32 = 16;
pub const IVPB_FLAG_SORT: i32 = 32;

extern "C" {
    pub fn ivpb_create(out: *mut *mut ivpb_ctx, device_ids: *const c_int, n_devices: c_int) -> c_int;
    pub fn ivpb_destroy(ctx: *mut ivpb_ctx);
    pub fn ivpb_last_error(ctx: *const ivpb_ctx) -> *const c_char;
    pub fn ivpb_device_count(ctx: *const ivpb_ctx) -> c_int;
    pub fn ivpb_builtin_problem(ctx: *mut ivpb_ctx, builtin_id: c_int, n: *mut c_int, p: *mut c_int, n_events: *mut c_int) -> c_int;
    pub fn ivpb_nvrtc_problem(ctx: *mut ivpb_ctx, cuda_src: *const c_char, n: c_int, p: c_int, n_events: c_int, has_jac: c_int, handle: *mut c_int) -> c_int;
    pub fn ivpb_solve_batch(ctx: *mut ivpb_ctx, problem: c_int, opt: *const ivpb_options, n: i64, t0: f64, tf: f64,
                            y0: *const f64, params: *const f64, out: *const ivpb_outputs) -> c_int;
    pub fn ivpb_solve_batch_device(ctx: *mut ivpb_ctx, problem: c_int, opt: *const ivpb_options, n: i64, t0: f64, tf: f64,
                                   d_y0: *const f64, d_params: *const f64, d_out: *const ivpb_outputs, stream: *mut c_void) -> c_int;
    pub fn ivpb_dense_eval(ctx: *mut ivpb_ctx, generation: u64, n: c_int, n_query: i64, traj: *const i64, ts: *const f64, y: *mut f64, ok: *mut i32) -> c_int;
    pub fn ivpb_dense_eval_extrapolate(ctx: *mut ivpb_ctx, generation: u64, n: c_int, n_query: i64, traj: *const i64, ts: *const f64, y: *mut f64, ok: *mut i32) -> c_int;
    pub fn ivpb_dense_span(ctx: *mut ivpb_ctx, generation: u64, first: i64, count: i64, t_start: *mut f64, t_end: *mut f64, n_seg: *mut i32) -> c_int;
    pub fn ivpb_dense_generation(ctx: *const ivpb_ctx) -> u64;
    pub fn ivpb_host_alloc(bytes: usize) -> *mut c_void;
    pub fn ivpb_host_free(p: *mut c_void);
    pub fn ivpb_launch_count(ctx: *const ivpb_ctx) -> u64;
    pub fn ivpb_measure_fp64_peak(ctx: *mut ivpb_ctx, tflops: *mut f64) -> c_int;
    pub fn ivpb_version() -> *const c_char;
}
