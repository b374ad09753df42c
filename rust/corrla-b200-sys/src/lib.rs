//! Raw FFI surface of `libcorrla_b200.so`; mirrors `include/corrla_b200.h` field for field.
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_void};

pub const CORRLA_OK: c_int = 0;
pub const CORRLA_ERR_INVALID: c_int = -1;
pub const CORRLA_ERR_RANK: c_int = -2;
pub const CORRLA_ERR_CUDA: c_int = -3;
pub const CORRLA_ERR_UNSUPPORTED: c_int = -4;
pub const CORRLA_ERR_ALLOC: c_int = -5;
pub const CORRLA_ERR_COMM: c_int = -6;
pub const CORRLA_ERR_NO_DEVICE: c_int = -7;

#[repr(C)]
pub struct corrla_ctx { _private: [u8; 0] }
#[repr(C)]
pub struct corrla_comm { _private: [u8; 0] }

#[repr(C)]
#[derive(Clone, Copy)]
pub struct corrla_rsvd_opts {
    pub seed: u64,
    pub omega: *const f64,
    pub omega_rs: i64,
    pub omega_cs: i64,
    pub omega_on_device: c_int,
    pub schedule: c_int,
    pub a_on_device: c_int,
    pub out_on_device: c_int,
    pub device: c_int,
    pub stream: *mut c_void,
    pub ctx: *mut corrla_ctx,
    pub comm: *mut corrla_comm,
    pub global_rows: i64,
    pub center: c_int,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct corrla_timings {
    pub total_ms: f64,
    pub h2d_ms: f64,
    pub device_ms: f64,
    pub d2h_ms: f64,
    pub gpu_launches: c_int,
    pub passes_over_a: c_int,
    pub qr_third_passes: c_int,
    pub qr_refills: c_int,
    pub jacobi_sweeps: c_int,
    pub live_columns: c_int,
    pub pass_launches: c_int,
    pub pass_ms: f64,
    pub pass_flops: f64,
    pub p2p_exchanges: c_int,
    pub streamed_chunks: c_int,
    pub jacobi_converged: c_int,
    pub fused_small: c_int,
}

extern "C" {
    pub fn corrla_rsvd_opts_default(opts: *mut corrla_rsvd_opts);
    pub fn corrla_rsvd_f64(a: *const f64, nrows: i64, ncols: i64, row_stride: i64, col_stride: i64,
                           n_rank: usize, n_iter: usize, n_oversamples: usize, opts: *const corrla_rsvd_opts,
                           u: *mut f64, s: *mut f64, vt: *mut f64, timings: *mut corrla_timings) -> c_int;
    pub fn corrla_power_iter_f64(a: *const f64, nrows: i64, ncols: i64, row_stride: i64, col_stride: i64,
                                 omega_rank: usize, n_iter: usize, opts: *const corrla_rsvd_opts, q: *mut f64,
                                 timings: *mut corrla_timings) -> c_int;
    pub fn corrla_rsvd_f32(a: *const f32, nrows: i64, ncols: i64, row_stride: i64, col_stride: i64,
                           n_rank: usize, n_iter: usize, n_oversamples: usize, opts: *const corrla_rsvd_opts,
                           u: *mut f32, s: *mut f32, vt: *mut f32, timings: *mut corrla_timings) -> c_int;
    pub fn corrla_power_iter_f32(a: *const f32, nrows: i64, ncols: i64, row_stride: i64, col_stride: i64,
                                 omega_rank: usize, n_iter: usize, opts: *const corrla_rsvd_opts, q: *mut f32,
                                 timings: *mut corrla_timings) -> c_int;
    pub fn corrla_rpca_f64(a: *const f64, nrows: i64, ncols: i64, row_stride: i64, col_stride: i64, n_rank: usize,
                           opts: *const corrla_rsvd_opts, s: *mut f64, components: *mut f64, means: *mut f64,
                           timings: *mut corrla_timings) -> c_int;
    pub fn corrla_par_matmul_f64(res: *mut f64, res_rs: i64, res_cs: i64, lhs: *const f64, lhs_rows: i64,
                                 lhs_cols: i64, lhs_rs: i64, lhs_cs: i64, rhs: *const f64, rhs_cols: i64,
                                 rhs_rs: i64, rhs_cs: i64, beta: f64, on_device: c_int,
                                 opts: *const corrla_rsvd_opts) -> c_int;
    pub fn corrla_random_mat_normal_f64(seed: u64, n_rows: i64, n_cols: i64, out: *mut f64, out_on_device: c_int,
                                        opts: *const corrla_rsvd_opts) -> c_int;
    pub fn corrla_dmdc_f64(x: *const f64, n_x: i64, n_snap: i64, x_rs: i64, x_cs: i64, u: *const f64, n_u: i64,
                           u_rs: i64, u_cs: i64, n_modes: usize, n_iters: usize, opts: *const corrla_rsvd_opts,
                           omega_y: *const f64, a_til: *mut f64, b: *mut f64, modes_scale: *mut f64, s_til: *mut f64,
                           u_hat: *mut f64, timings: *mut corrla_timings) -> c_int;
    pub fn corrla_pod_f64(x: *const f64, n_snap: i64, n_points: i64, row_stride: i64, col_stride: i64, n_modes: usize,
                          opts: *const corrla_rsvd_opts, modes: *mut f64, weights: *mut f64, s: *mut f64,
                          timings: *mut corrla_timings) -> c_int;
    pub fn corrla_cov_f64(x: *const f64, nrows: i64, ncols: i64, row_stride: i64, col_stride: i64, kind: c_int,
                          scale: f64, opts: *const corrla_rsvd_opts, out: *mut f64, means: *mut f64, evals: *mut f64,
                          evecs: *mut f64) -> c_int;
    pub fn corrla_active_ss_f64(x: *const f64, n_samples: i64, n_features: i64, x_rs: i64, x_cs: i64, y: *const f64,
                                y_stride: i64, order: c_int, n_nbr: c_int, opts: *const corrla_rsvd_opts,
                                evals: *mut f64, evecs: *mut f64, grad_mat: *mut f64, n_deficient: *mut c_int) -> c_int;
    pub fn corrla_poly_grad_at_f64(x: *const f64, n_samples: i64, n_features: i64, x_rs: i64, x_cs: i64, y: *const f64,
                                   y_stride: i64, order: c_int, n_nbr: c_int, xq: *const f64, n_query: i64, q_rs: i64,
                                   q_cs: i64, opts: *const corrla_rsvd_opts, grad_out: *mut f64,
                                   n_deficient: *mut c_int) -> c_int;
    pub fn corrla_thin_q_f64(a: *const f64, nrows: i64, ncols: i64, row_stride: i64, col_stride: i64,
                             on_device: c_int, opts: *const corrla_rsvd_opts, q: *mut f64, rank_out: *mut c_int) -> c_int;
    pub fn corrla_host_alloc(bytes: usize) -> *mut c_void;
    pub fn corrla_host_free(p: *mut c_void, bytes: usize);
    pub fn corrla_ctx_create(device: c_int, out: *mut *mut corrla_ctx) -> c_int;
    pub fn corrla_ctx_destroy(ctx: *mut corrla_ctx);
    pub fn corrla_ctx_trim(ctx: *mut corrla_ctx) -> usize;
    pub fn corrla_comm_unique_id(id: *mut u8) -> c_int;
    pub fn corrla_comm_init(id: *const u8, rank: c_int, nranks: c_int, device: c_int, out: *mut *mut corrla_comm) -> c_int;
    pub fn corrla_comm_destroy(comm: *mut corrla_comm);
    pub fn corrla_comm_rank(comm: *const corrla_comm) -> c_int;
    pub fn corrla_comm_size(comm: *const corrla_comm) -> c_int;
    pub fn corrla_status_str(status: c_int) -> *const c_char;
    pub fn corrla_last_error() -> *const c_char;
    pub fn corrla_version() -> *const c_char;
}
