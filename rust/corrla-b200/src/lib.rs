//! Drop-in bodies for `corrla_rs::lib_math_utils::random_svd::{random_svd, power_iter}` and
//! `mat_utils::par_matmul_helper` (f64 only -- every caller in the reference instantiates f64).
//! Same signatures, same shapes (U m x k, S k x 1, Vt k x n), and the same failure mode: the reference
//! panics (faer asserts / out-of-range `get`), so a non-zero status panics here too.
use corrla_b200_sys as sys;
use faer::{Mat, MatMut, MatRef};
use std::ffi::CStr;

fn check(status: i32) {
    if status != sys::CORRLA_OK {
        let (what, detail) = unsafe {
            (CStr::from_ptr(sys::corrla_status_str(status)).to_string_lossy().into_owned(),
             CStr::from_ptr(sys::corrla_last_error()).to_string_lossy().into_owned())
        };
        panic!("corrla_b200: {what}: {detail}");
    }
}

fn default_opts() -> sys::corrla_rsvd_opts {
    let mut o = std::mem::MaybeUninit::<sys::corrla_rsvd_opts>::zeroed();
    unsafe { sys::corrla_rsvd_opts_default(o.as_mut_ptr()); o.assume_init() }
}

/// Replaces random_svd.rs:63-110.
pub fn random_svd(a_mat: MatRef<f64>, omega_rank: usize, n_iter: usize, n_oversamples: usize)
    -> (Mat<f64>, Mat<f64>, Mat<f64>)
{
    let (m, n) = (a_mat.nrows(), a_mat.ncols());
    let mut u = Mat::<f64>::zeros(m, omega_rank);
    let mut s = Mat::<f64>::zeros(omega_rank, 1);
    let mut vt = Mat::<f64>::zeros(omega_rank, n);
    assert_eq!(u.col_stride() as usize, m.max(1));        // outputs are written column-major, contiguous
    let mut opts = default_opts();
    opts.seed = rand_seed();
    let st = unsafe {
        sys::corrla_rsvd_f64(a_mat.as_ptr(), m as i64, n as i64, a_mat.row_stride() as i64, a_mat.col_stride() as i64,
                             omega_rank, n_iter, n_oversamples, &opts, u.as_mut().as_ptr_mut(),
                             s.as_mut().as_ptr_mut(), vt.as_mut().as_ptr_mut(), std::ptr::null_mut())
    };
    check(st);
    (u, s, vt)
}

/// `random_svd::<f32>`: the reference function is generic over `T: RealField + Float` (random_svd.rs:63-66).
pub fn random_svd_f32(a_mat: MatRef<f32>, omega_rank: usize, n_iter: usize, n_oversamples: usize)
    -> (Mat<f32>, Mat<f32>, Mat<f32>)
{
    let (m, n) = (a_mat.nrows(), a_mat.ncols());
    let mut u = Mat::<f32>::zeros(m, omega_rank);
    let mut s = Mat::<f32>::zeros(omega_rank, 1);
    let mut vt = Mat::<f32>::zeros(omega_rank, n);
    let mut opts = default_opts();
    opts.seed = rand_seed();
    let st = unsafe {
        sys::corrla_rsvd_f32(a_mat.as_ptr(), m as i64, n as i64, a_mat.row_stride() as i64, a_mat.col_stride() as i64,
                             omega_rank, n_iter, n_oversamples, &opts, u.as_mut().as_ptr_mut(),
                             s.as_mut().as_ptr_mut(), vt.as_mut().as_ptr_mut(), std::ptr::null_mut())
    };
    check(st);
    (u, s, vt)
}

/// Replaces random_svd.rs:15-59.
pub fn power_iter(a_mat: MatRef<f64>, omega_rank: usize, n_iter: usize) -> Mat<f64> {
    let m = a_mat.nrows();
    let mut q = Mat::<f64>::zeros(m, omega_rank);
    let mut opts = default_opts();
    opts.seed = rand_seed();
    let st = unsafe {
        sys::corrla_power_iter_f64(a_mat.as_ptr(), m as i64, a_mat.ncols() as i64, a_mat.row_stride() as i64,
                                   a_mat.col_stride() as i64, omega_rank, n_iter, &opts,
                                   q.as_mut().as_ptr_mut(), std::ptr::null_mut())
    };
    check(st);
    q
}

/// Replaces mat_utils.rs:20-33 (alpha = None: overwrite).  `n_threads` is ignored, as in the reference.
pub fn par_matmul_helper(mut res: MatMut<f64>, lhs: MatRef<f64>, rhs: MatRef<f64>, beta: f64, _n_threads: usize) {
    assert_eq!(lhs.ncols(), rhs.nrows());
    assert_eq!((res.nrows(), res.ncols()), (lhs.nrows(), rhs.ncols()));
    let st = unsafe {
        sys::corrla_par_matmul_f64(res.as_ptr_mut(), res.row_stride() as i64, res.col_stride() as i64,
                                   lhs.as_ptr(), lhs.nrows() as i64, lhs.ncols() as i64, lhs.row_stride() as i64,
                                   lhs.col_stride() as i64, rhs.as_ptr(), rhs.ncols() as i64, rhs.row_stride() as i64,
                                   rhs.col_stride() as i64, beta, 0, std::ptr::null())
    };
    check(st);
}

/// The device-side part of `DMDc::_calc_dmdc_modes` / `_calc_modes` (dmd_rom.rs:64-109, :128-139): returns
/// (`_A` r x r, `_B` n_x x n_u, `tmp_modes_scale` n_x x r, `s_til` r x 1, `u_hat` n_x x r).  `_calc_eigs` stays with the
/// caller: `modes_re = tmp_modes_scale * w_re`, `modes_im = tmp_modes_scale * w_im` (two `par_matmul_helper` calls).
pub fn dmdc_operators(x_data: MatRef<f64>, u_data: MatRef<f64>, n_modes: usize, n_iters: usize)
    -> (Mat<f64>, Mat<f64>, Mat<f64>, Mat<f64>, Mat<f64>)
{
    assert_eq!(x_data.ncols(), u_data.ncols());
    let (n_x, n_snap, n_u) = (x_data.nrows(), x_data.ncols(), u_data.nrows());
    let mut a_til = Mat::<f64>::zeros(n_modes, n_modes);
    let mut b = Mat::<f64>::zeros(n_x, n_u);
    let mut modes_scale = Mat::<f64>::zeros(n_x, n_modes);
    let mut s_til = Mat::<f64>::zeros(n_modes, 1);
    let mut u_hat = Mat::<f64>::zeros(n_x, n_modes);
    let mut opts = default_opts();
    opts.seed = rand_seed();
    let st = unsafe {
        sys::corrla_dmdc_f64(x_data.as_ptr(), n_x as i64, n_snap as i64, x_data.row_stride() as i64,
                             x_data.col_stride() as i64, u_data.as_ptr(), n_u as i64, u_data.row_stride() as i64,
                             u_data.col_stride() as i64, n_modes, n_iters, &opts, std::ptr::null(),
                             a_til.as_mut().as_ptr_mut(), b.as_mut().as_ptr_mut(), modes_scale.as_mut().as_ptr_mut(),
                             s_til.as_mut().as_ptr_mut(), u_hat.as_mut().as_ptr_mut(), std::ptr::null_mut())
    };
    check(st);
    (a_til, b, modes_scale, s_til, u_hat)
}

/// `PodI::_modes` + `PodI::_weights` (pod_rom.rs:53-75): (modes n_points x n_modes, mode_weights n_snap x n_modes).
pub fn pod_modes_weights(x_data: MatRef<f64>, n_modes: usize) -> (Mat<f64>, Mat<f64>) {
    let (n_snap, n_points) = (x_data.nrows(), x_data.ncols());
    let mut modes = Mat::<f64>::zeros(n_points, n_modes);
    let mut weights = Mat::<f64>::zeros(n_snap, n_modes);
    let mut opts = default_opts();
    opts.seed = rand_seed();
    let st = unsafe {
        sys::corrla_pod_f64(x_data.as_ptr(), n_snap as i64, n_points as i64, x_data.row_stride() as i64,
                            x_data.col_stride() as i64, n_modes, &opts, modes.as_mut().as_ptr_mut(),
                            weights.as_mut().as_ptr_mut(), std::ptr::null_mut(), std::ptr::null_mut())
    };
    check(st);
    (modes, weights)
}

fn rand_seed() -> u64 {
    // the reference draws Omega from thread_rng() (mat_utils.rs:166-173): unseeded, different every call
    use std::time::{SystemTime, UNIX_EPOCH};
    let t = SystemTime::now().duration_since(UNIX_EPOCH).map(|d| d.as_nanos() as u64).unwrap_or(0);
    t ^ (&t as *const u64 as u64).rotate_left(32)
}
