/* libcorrla_b200.so -- C ABI of the B200-native randomized-SVD engine.
 *
 * Drop-in boundary for the one hot path of wgurecky/CORRLA_RS.  The reference has no FFI of its own; the
 * entry points below are what a Rust `-sys` crate / the pyo3 module would bind in place of the pure-Rust
 * bodies (paths relative to the reference checkout):
 *
 *   corrla_rsvd_f64               <- random_svd<T>()        src/lib_math_utils/random_svd.rs:63-110
 *                                    and the pyo3 rsvd()    src/lib_math_utils_py.rs:21-36
 *   corrla_power_iter_f64         <- power_iter<T>()        src/lib_math_utils/random_svd.rs:15-59
 *   corrla_par_matmul_f64         <- par_matmul_helper<T>() src/lib_math_utils/mat_utils.rs:20-33
 *   corrla_random_mat_normal_f64  <- random_mat_normal<T>() src/lib_math_utils/mat_utils.rs:161-175
 *
 * Conventions: plain pointers and sizes only; every function returns CORRLA_OK (0) or a negative
 * corrla_status and never unwinds.  Matrices are described like a faer MatRef: (ptr, nrows, ncols,
 * row_stride, col_stride) with strides in ELEMENTS; inputs are borrowed and never modified.  Outputs are
 * written column-major (faer `Mat` layout) into caller-owned buffers.  All arithmetic is IEEE f64 on the GPU
 * (DMMA tensor pipe); there is no CPU fallback: without a usable CUDA device the calls fail with
 * CORRLA_ERR_NO_DEVICE.
 */
#ifndef CORRLA_B200_H
#define CORRLA_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define CORRLA_API __attribute__((visibility("default")))
#else
#define CORRLA_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef enum corrla_status {
  CORRLA_OK = 0,
  CORRLA_ERR_INVALID = -1,      /* bad argument (null pointer, zero-sized matrix, bad strides) */
  CORRLA_ERR_RANK = -2,         /* n_rank > min(n_rank + n_oversamples, ncols(thin a)): the reference panics here
                                   (out-of-range get at random_svd.rs:98-107) */
  CORRLA_ERR_CUDA = -3,         /* a CUDA call failed; see corrla_last_error() */
  CORRLA_ERR_UNSUPPORTED = -4,  /* n_rank + n_oversamples > 2048, or an option combination that is not provided */
  CORRLA_ERR_ALLOC = -5,        /* device or host allocation failed */
  CORRLA_ERR_COMM = -6,         /* NCCL failure or libnccl not loadable */
  CORRLA_ERR_NO_DEVICE = -7     /* no CUDA device / driver */
} corrla_status;

/* Opaque per-GPU context: stream, cached device buffers (grow-only), split-K workspace. Thread-safe per ctx. */
typedef struct corrla_ctx corrla_ctx;
/* Opaque communicator over the GPUs that hold the row shards of one tall matrix (one process per GPU). */
typedef struct corrla_comm corrla_comm;

typedef struct corrla_rsvd_opts {
  uint64_t seed;           /* Philox key for Omega when `omega` is NULL */
  const double* omega;     /* NULL => Omega = Philox4x32-10 N(0,1); else the injected test matrix,
                              ncols(thin a) x l with element strides (omega_rs, omega_cs) */
  int64_t omega_rs, omega_cs;
  int omega_on_device;     /* 0: host pointer, 1: device pointer */
  int schedule;            /* 0 = reference (QR only when i > 2, Frobenius scaling each trip; random_svd.rs:35-56)
                              1 = stabilised (re-orthonormalise before every A^T*Y) */
  int a_on_device;         /* 0: `a` is a host pointer (copied in, counted in timings.h2d_ms); 1: device pointer */
  int out_on_device;       /* 0: outputs are host pointers; 1: device pointers */
  int device;              /* CUDA ordinal, or -1 for the current device */
  void* stream;            /* cudaStream_t to enqueue on, or NULL for the context's own (non-blocking) stream; the
                              legacy default stream must be named as cudaStreamLegacy, not as NULL */
  corrla_ctx* ctx;         /* reuse buffers across calls; NULL => a temporary context per call */
  corrla_comm* comm;       /* NULL => single GPU.  Else `a` is this rank's block of rows of the thin matrix */
  int64_t global_rows;     /* with comm: total rows over all ranks (0 => computed with an all-reduce) */
  int center;              /* 1 => decompose a - 1*mean_cols(a)^T (PCA centring, center_mat_col: mat_utils.rs:482-502)
                              without forming the centred copy when the matrix is tall (rank-1 corrections inside the
                              passes); fat inputs get an explicit centred copy */
} corrla_rsvd_opts;

typedef struct corrla_timings {
  double total_ms;         /* whole call, host clock */
  double h2d_ms;           /* host->device copy of A (0 when a_on_device) */
  double device_ms;        /* CUDA-event time of everything enqueued after A is resident */
  double d2h_ms;           /* device->host copy of the results */
  int gpu_launches;        /* kernels of this library launched by the call */
  int passes_over_a;       /* 2 + 2*n_iter */
  int qr_third_passes;     /* how many QR calls left plain CholeskyQR2 for the sketch-preconditioned (robust) stage */
  int qr_refills;          /* how many CholeskyQR calls replaced numerically dependent columns by random vectors */
  int jacobi_sweeps;
  int live_columns;        /* numerical rank kept by the last CholeskyQR (<= l) */
  int pass_launches;       /* launches of the DMMA GEMM kernel that stream A (== passes_over_a) */
  double pass_ms;          /* summed CUDA-event duration of those launches (the dominant kernel) */
  double pass_flops;       /* algorithmic flops of one such launch on this GPU: 2 * local_rows * ncols * l */
  int p2p_exchanges;       /* cross-rank sums done inside the reduction kernel over NVLink peer memory (0 => NCCL only) */
  int streamed_chunks;     /* host input: row chunks whose first product(s) ran behind the host->device copy (0 = copied first) */
  int jacobi_converged;    /* 1: the Jacobi SVD of the l x l core met its tolerance; 0: it stopped at the sweep limit (a
                              warning -- the factors are still orthogonal to working accuracy, sigma to ~1e-10) */
  int fused_small;         /* 1: the whole call ran as ONE kernel with the matrix resident in shared memory (tiny inputs) */
} corrla_timings;

CORRLA_API void corrla_rsvd_opts_default(corrla_rsvd_opts* opts);

/* (U, S, Vt) = random_svd(a, n_rank, n_iter, n_oversamples).
 *   u  : nrows x n_rank, column-major      s : n_rank      vt : n_rank x ncols, column-major
 * For nrows < ncols the routine works on the transposed view exactly like random_svd.rs:69-74,96-102.
 * With opts->comm the caller passes its local rows of the THIN matrix (nrows = local rows >= 1, all ranks
 * the same ncols); u receives the matching local rows, s and vt are replicated. */
CORRLA_API int corrla_rsvd_f64(const double* a, int64_t nrows, int64_t ncols, int64_t row_stride, int64_t col_stride,
                    size_t n_rank, size_t n_iter, size_t n_oversamples, const corrla_rsvd_opts* opts,
                    double* u, double* s, double* vt, corrla_timings* timings);

/* q (nrows x omega_rank, column-major) = power_iter(a, omega_rank, n_iter); `a` is used as given (no fat/thin
 * swap, as in the reference). */
CORRLA_API int corrla_power_iter_f64(const double* a, int64_t nrows, int64_t ncols, int64_t row_stride, int64_t col_stride,
                          size_t omega_rank, size_t n_iter, const corrla_rsvd_opts* opts, double* q,
                          corrla_timings* timings);

/* T = f32 instantiation of random_svd<T> / power_iter<T> (random_svd.rs:15-18, :63-66 are generic over
 * T: faer::RealField + Float).  Same arguments with float matrices; opts->omega, when given, is still a double matrix.
 * The data are widened to f64 once on the device, the f64 engine runs, and the factors are rounded to f32 on the way
 * out: at least as accurate as an all-f32 evaluation.  No communicator. */
CORRLA_API int corrla_rsvd_f32(const float* a, int64_t nrows, int64_t ncols, int64_t row_stride, int64_t col_stride,
                    size_t n_rank, size_t n_iter, size_t n_oversamples, const corrla_rsvd_opts* opts,
                    float* u, float* s, float* vt, corrla_timings* timings);
CORRLA_API int corrla_power_iter_f32(const float* a, int64_t nrows, int64_t ncols, int64_t row_stride, int64_t col_stride,
                          size_t omega_rank, size_t n_iter, const corrla_rsvd_opts* opts, float* q,
                          corrla_timings* timings);

/* res = beta * lhs * rhs (alpha = None: the destination is overwritten); rhs is skinny (column panels of <= 128).
 * lhs is m x kk, rhs is kk x rhs_cols, res is m x rhs_cols; all strided, all on the host or all on the device
 * (on_device).  opts may be NULL (defaults) -- only device/stream/ctx are read. */
CORRLA_API int corrla_par_matmul_f64(double* res, int64_t res_rs, int64_t res_cs,
                          const double* lhs, int64_t lhs_rows, int64_t lhs_cols, int64_t lhs_rs, int64_t lhs_cs,
                          const double* rhs, int64_t rhs_cols, int64_t rhs_rs, int64_t rhs_cs,
                          double beta, int on_device, const corrla_rsvd_opts* opts);

/* out (n_rows x n_cols, column-major like faer Mat::from_fn) = i.i.d. N(0,1) from Philox4x32-10 keyed by seed.
 * Element (i, j) is draw number i*n_cols + j. */
CORRLA_API int corrla_random_mat_normal_f64(uint64_t seed, int64_t n_rows, int64_t n_cols, double* out, int out_on_device,
                                 const corrla_rsvd_opts* opts);

/* PCA by RSVD, PcaRsvd::new (src/lib_math_utils/pca_rsvd.rs:56-82) behind the pyo3 rpca (lib_math_utils_py.rs:38-55):
 * column means, random_svd(centred a, n_rank, 20, min(ncols, 10)).  s: n_rank singular values; components: n_rank x ncols
 * column-major (= Vt); means: ncols (optional, may be NULL).  U is never formed. */
CORRLA_API int corrla_rpca_f64(const double* a, int64_t nrows, int64_t ncols, int64_t row_stride, int64_t col_stride,
                    size_t n_rank, const corrla_rsvd_opts* opts, double* s, double* components, double* means,
                    corrla_timings* timings);

/* DMD with control, DMDc::new / _calc_dmdc_modes / _calc_modes (src/lib_math_utils/dmd_rom.rs:46-146) behind the pyo3
 * PyDMDc (lib_math_utils_py.rs:255-283): every step that touches the tall snapshot matrix.
 *   x : n_x x n_snap snapshots, u : n_u x n_snap control inputs (element strides; both host or both device,
 *       opts->a_on_device).  n_u may be 0 (plain DMD; the reference requires a control matrix).
 *   The two RSVDs (:72, :82; n_oversamples = 12) run on the views [x; u][:, :n_snap-1] and x[:, 1:] of ONE device copy
 *   of the stacked snapshots -- mat_vstack (:66) and the shifted copies are never materialised twice.
 *   opts->omega / omega_y (optional, same stride/residency fields) inject the two sketch matrices, shaped
 *   n_thin x l of the respective view (as corrla_rsvd_f64 takes them); otherwise Philox(seed), Philox(seed + 1).
 * Outputs (host, or device with opts->out_on_device; all column-major; any may be NULL):
 *   a_til       r x r      = self._A (:95-97, eq. 29)          r = n_modes
 *   b           n_x x n_u  = self._B = u_hat * b_til (:100-106, eq. 30)
 *   modes_scale n_x x r    = tmp_modes_scale (:133-139, eq. 36): modes_re/im = modes_scale * Re/Im(W), W = eigenvectors
 *                            of a_til.  The r x r eigendecomposition (:116) stays with the caller.
 *   s_til       r          singular values of the input space;  u_hat  n_x x r  basis of the output space
 * With opts->comm every rank passes its block of state rows of x (n_x = local rows, opts->global_rows = all of them)
 * and the whole u; a_til and s_til are replicated, b / modes_scale / u_hat hold the local rows. */
CORRLA_API int corrla_dmdc_f64(const double* x, int64_t n_x, int64_t n_snap, int64_t x_rs, int64_t x_cs,
                    const double* u, int64_t n_u, int64_t u_rs, int64_t u_cs,
                    size_t n_modes, size_t n_iters, const corrla_rsvd_opts* opts, const double* omega_y,
                    double* a_til, double* b, double* modes_scale, double* s_til, double* u_hat,
                    corrla_timings* timings);

/* POD modes and weights, PodI::_modes + PodI::_weights (src/lib_math_utils/pod_rom.rs:53-75) behind the pyo3 PyPodI
 * (lib_math_utils_py.rs:223-250).  x : n_snap x n_points, one snapshot per row (fat in practice: the RSVD runs on the
 * transposed view).  modes : n_points x n_modes column-major = V of random_svd(x, n_modes, 10, 10);
 * weights : n_snap x n_modes column-major = x * modes  (the reference's pinv(modes) * x_row^T per snapshot: the
 * modes are orthonormal, so the pseudo-inverse is the transpose).  s (optional): n_modes singular values.
 * With opts->comm every rank passes its block of points (columns of x; n_points = local count, opts->global_rows = all
 * of them): modes holds the local points, weights and s are replicated. */
CORRLA_API int corrla_pod_f64(const double* x, int64_t n_snap, int64_t n_points, int64_t row_stride, int64_t col_stride,
                   size_t n_modes, const corrla_rsvd_opts* opts, double* modes, double* weights, double* s,
                   corrla_timings* timings);

/* Second-moment matrices of a tall sample matrix on the Gram kernel (the part that is a GEMM):
 *   kind 0  out = scale * x^T x                       ActiveSsRsvd::fit's grad_mat * grad_mat^T / N
 *                                                     (src/lib_math_utils/active_subspaces.rs:253; pass grad_mat^T as x)
 *   kind 1  out = (x - mean)^T (x - mean) / (N - 1)   mat_cov_centered (src/lib_math_utils/stats_corr.rs:32-43)
 *   kind 2  the same on z-scored columns              pearson_corr (stats_corr.rs:14-28; std with N - 1, mat_utils.rs:122-160)
 * x : nrows x ncols samples-by-features (element strides; host, or device with opts->a_on_device), ncols <= 128.
 * out : ncols x ncols (symmetric).  means (optional) : ncols column means (kinds 1, 2).
 * evals / evecs (optional, both or neither): eigenvalues in descending order and eigenvectors (ncols x ncols,
 * column-major, one per column) of `out` by one-sided Jacobi -- the sorted decomposition ActiveSsRsvd::fit takes from
 * faer (active_subspaces.rs:259-271).  With opts->comm the rows are sharded and every output is replicated. */
#define CORRLA_COV_GRAM 0
#define CORRLA_COV_CENTERED 1
#define CORRLA_COV_PEARSON 2
CORRLA_API int corrla_cov_f64(const double* x, int64_t nrows, int64_t ncols, int64_t row_stride, int64_t col_stride,
                   int kind, double scale, const corrla_rsvd_opts* opts, double* out, double* means,
                   double* evals, double* evecs);

/* Active subspace from samples: the gradient matrix of PolyGradientEstimator + ActiveSsRsvd::create_grad_mat
 * (src/lib_math_utils/active_subspaces.rs:66-141, :226-238) and the sorted eigendecomposition of grad grad^T / N of
 * ActiveSsRsvd::fit (:248-278), behind the pyo3 active_ss (lib_math_utils_py.rs:56-86).
 *   x : n_samples x n_features (element strides), y : n_samples values (element stride y_stride); host, or device with
 *       opts->a_on_device.  order 1 (linear fit) or 2 (quadratic fit); n_nbr nearest samples per fit (<= 128).
 *   The kd-tree walk becomes one exact brute-force nearest-neighbour kernel over all samples, the per-sample
 *   pseudo-inverse a batched Householder least-squares kernel; order 2 differentiates the fitted polynomial analytically
 *   where the reference takes a forward difference with eps = 1e-10.
 *   evals : n_features eigenvalues, descending; evecs : n_features x n_features column-major (components_);
 *   grad_mat (optional) : n_features x n_samples column-major; n_deficient (optional) : samples whose neighbourhood
 *   gave a numerically rank-deficient design matrix.  n_features <= 128 (order 1) / 14 (order 2).
 * CORRLA_ERR_INVALID where the reference asserts (too few samples or neighbours for the fit). */
CORRLA_API int corrla_active_ss_f64(const double* x, int64_t n_samples, int64_t n_features, int64_t x_rs, int64_t x_cs,
                         const double* y, int64_t y_stride, int order, int n_nbr, const corrla_rsvd_opts* opts,
                         double* evals, double* evecs, double* grad_mat, int* n_deficient);

/* PolyGradientEstimator::grad_at (src/lib_math_utils/active_subspaces.rs:66-141) for a batch of evaluation points that
 * need not be samples: for every row of xq (n_query x n_features, element strides) the n_nbr nearest samples of x are
 * found (exact search, same distance and tie rule as above), the order-1 or order-2 polynomial is fitted through them
 * and its gradient is taken at the query point.  grad_out : n_query x n_features row-major (one grad_at row per query;
 * equivalently n_features x n_query column-major, the layout of create_grad_mat, :226-238).  Same limits and error
 * behaviour as corrla_active_ss_f64; x, y and xq share opts->a_on_device, grad_out follows opts->out_on_device. */
CORRLA_API int corrla_poly_grad_at_f64(const double* x, int64_t n_samples, int64_t n_features, int64_t x_rs, int64_t x_cs,
                         const double* y, int64_t y_stride, int order, int n_nbr, const double* xq, int64_t n_query,
                         int64_t q_rs, int64_t q_cs, const corrla_rsvd_opts* opts, double* grad_out, int* n_deficient);

/* thin Q (nrows x ncols, column-major) of a tall matrix by adaptive CholeskyQR2/3 (the engine's replacement for
 * faer qr().compute_thin_q(), random_svd.rs:38,:57).  ncols <= 2048 (column panels above 128).  rank_out (optional) =
 * live columns found before the orthonormal completion. */
CORRLA_API int corrla_thin_q_f64(const double* a, int64_t nrows, int64_t ncols, int64_t row_stride, int64_t col_stride,
                      int on_device, const corrla_rsvd_opts* opts, double* q, int* rank_out);

/* Host buffers for large outputs: page-aligned, advised to use transparent huge pages (2 MiB), so that the threaded
 * device->host copy is not dominated by 4 KiB first-touch page faults.  Any host memory works as an output; this is
 * only faster.  Free with corrla_host_free. */
CORRLA_API void* corrla_host_alloc(size_t bytes);
CORRLA_API void corrla_host_free(void* p, size_t bytes);

/* contexts.  A context keeps its device buffers between calls (grow-only): after one host-path call on a large
 * matrix the device copy of A, Y and the outputs stay allocated.  corrla_ctx_trim gives them back (returns the bytes
 * released); the next call allocates again.  Failures reported by a kernel (peer-exchange or cluster-exchange
 * timeouts) are returned by the call that hit them whether or not a timings struct was passed. */
CORRLA_API int corrla_ctx_create(int device, corrla_ctx** out);
CORRLA_API void corrla_ctx_destroy(corrla_ctx* ctx);
CORRLA_API size_t corrla_ctx_trim(corrla_ctx* ctx);

/* communicator (NCCL, loaded with dlopen("libnccl.so.2")); id is the 128-byte ncclUniqueId from rank 0 */
CORRLA_API int corrla_comm_unique_id(unsigned char id[128]);
CORRLA_API int corrla_comm_init(const unsigned char id[128], int rank, int nranks, int device, corrla_comm** out);
CORRLA_API void corrla_comm_destroy(corrla_comm* comm);
CORRLA_API int corrla_comm_rank(const corrla_comm* comm);
CORRLA_API int corrla_comm_size(const corrla_comm* comm);

CORRLA_API const char* corrla_status_str(int status);
CORRLA_API const char* corrla_last_error(void);   /* thread-local text of the last failure */
CORRLA_API const char* corrla_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CORRLA_B200_H */
