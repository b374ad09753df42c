// Small single-CTA / elementwise kernels around the skinny GEMM: Philox test matrix, Cholesky + triangular
// inverse for CholeskyQR, one-sided Jacobi SVD of the l x l core, strided repack.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace corrla {

// Omega ~ N(0,1), rows x cols, row-major with pitch ld (pad untouched).  Replaces random_mat_normal
// (reference src/lib_math_utils/mat_utils.rs:161-175) with a counter-based generator so every GPU
// regenerates the same matrix without a broadcast.  oracle/ref_rsvd.py:philox_normal restates it.
cudaError_t philox_normal_launch(double* out, int64_t rows, int cols, int64_t ld, uint64_t seed, cudaStream_t s);

// kCholCheck: plain, and *flag3 = (min pivot ratio < 1e-3).
// kCholProbe: plain, and flag3[0] = (min pivot ratio < 1e-8, i.e. cond(Y) beyond ~1e4, or a deflated column): the
//             robust sketch-preconditioned stage is needed; flag3[1] = !flag3[0]: plain CholeskyQR2 may proceed.
enum CholMode { kCholAuto = 0, kCholPlain = 1, kCholCheck = 2, kCholProbe = 3 };

// Upper Cholesky G = R^T R of the l x l Gram matrix (row-major, pitch ldg) in shared memory, then the
// "deflated" inverse T = R^-1 (columns whose pivot vanished are zero, so Q = Y*T has exact zero columns there).
// T is written as Lrows x ldt (zero outside the upper triangle).  kCholAuto: if the smallest pivot ratio
// falls under 1e-10 the factorisation is redone on G + shift*I (shifted CholeskyQR3, Fukaya et al. 2020)
// and *flag3 = 1 (a third pass is needed), else *flag3 = 0.
// info[0] = live columns, info[1] = shifted?, dinfo[0] = smallest pivot ratio.
// Together with the Gram/apply GEMMs this replaces faer's qr().compute_thin_q() at random_svd.rs:38 and :57.
// deadmask (optional, l ints) receives 1 for deflated columns and *flag_dead = (any deflated column).
cudaError_t chol_inv_launch(const double* G, int ldg, int l, double* T, int Lrows, int ldt, int mode,
                            double global_rows, int* flag3, int* info, double* dinfo, int* deadmask, int* flag_dead,
                            const int* cond_flag, cudaStream_t s);

// X[i][j] = N(0,1) for every column j with deadmask[j] != 0 (rows x l, pitch ld); draw number
// ((stream_id << 40) + i) * 128 + j of Philox4x32-10 keyed by seed.  Runs only if *cond_flag != 0.
cudaError_t refill_dead_launch(double* X, int64_t rows, int l, int64_t ld, const int* deadmask, uint64_t seed,
                               uint64_t stream_id, const int* cond_flag, cudaStream_t s);

// Columns whose cosine is below this in EVERY pair of a sweep are orthogonal to ~cos^2 <= 1e-15 after that sweep (the
// cyclic Jacobi method converges quadratically), so the sweep that would only confirm convergence is not run.
constexpr double kJacobiNearCos2 = 9e-16;      // (3e-8)^2

// One-sided (Hestenes) Jacobi SVD of the l x l matrix W (row-major, pitch ldw): W = Ur * diag(sigma) * Vr^T,
// sigma sorted descending.  Ur, Vr are written as Lrows x ldo row-major, zero padded (ready to be a GEMM B
// operand).  scratch: 2*l*(l|1) doubles of global memory, used when the matrices do not fit in shared memory.
// Replaces faer's svd() at random_svd.rs:89 (after the QR preconditioning done by the caller).
cudaError_t jacobi_svd_launch(const double* W, int ldw, int l, double* sigma, double* Vr, double* Ur, int Lrows,
                              int ldo, double* scratch, int* info, cudaStream_t s, int transpose = 1);

// T = 3/2 I - 1/2 G on the leading l x l block (G: upper triangle valid, pitch ldg), zero elsewhere in Lrows x ldt.
// For a Gram matrix G = I + E of nearly orthonormal columns X, X*T has orthogonality error (3/8) E^2: the symmetric
// (Loewdin) orthogonalisation to first order, which replaces the second Cholesky of CholeskyQR2 when |E| <~ 1e-8.
// *redo (device int) = 1 when ||E||_F > 2e-8: the caller then runs the real Cholesky behind that flag.
cudaError_t lowdin_launch(const double* G, int ldg, int l, double* T, int Lrows, int ldt, int* redo, cudaStream_t s);

// Cluster variant (jacobi_cluster.cu): row slabs of X and V in the shared memory of 4 or 8 CTAs, partial dot products
// exchanged through DSMEM.  cudaErrorNotSupported when it does not apply (l < 32, slabs too large, or
// CORRLA_B200_JACOBI_CLUSTER=0); jacobi_svd_launch tries it first.
// Ring variant (jacobi_ring.cu, l <= 256, first choice): column pairs in the registers of the warps of a cluster, one
// column per warp handed to a neighbour between steps (odd-even ordering); CORRLA_B200_JACOBI_RING=0 disables it.
cudaError_t jacobi_svd_ring_launch(const double* W, int ldw, int l, double* sigma, double* Vr, double* Ur, int Lrows,
                                   int ldo, int* info, cudaStream_t s, int transpose);
cudaError_t jacobi_svd_cluster_launch(const double* W, int ldw, int l, double* sigma, double* Vr, double* Ur, int Lrows,
                                      int ldo, int* info, cudaStream_t s, int transpose);

// dst[i*ldd + j] = scale * src[i*rs + j*cs] for i < rows, j < cols (any strides, device pointers).
cudaError_t repack_launch(const double* src, int64_t rows, int64_t cols, int64_t rs, int64_t cs, double* dst,
                          int64_t ldd, cudaStream_t s, double scale = 1.0);

// dst[i*drs + j*dcs] = src[i*ld + j]  (scatter a padded row-major matrix to arbitrary output strides)
cudaError_t scatter_launch(const double* src, int64_t rows, int64_t cols, int64_t ld, double* dst, int64_t drs,
                           int64_t dcs, cudaStream_t s);

// ---- sketch-preconditioned CholeskyQR (randomized Householder-Cholesky, Higgins et al. 2023) ------------------------
// Sparse sign embedding of the tall matrix Y (rows x Lc, pitch ld): each row is added, with a random sign, to
// kSketchZeta of the s buckets.  Rows are taken in tiles of 64; inside a tile hash t sends row j to bucket
// (a_t * j + offset(tile, t)) mod s with a_t coprime to s, an injection, so one (tile, t) step updates 64 distinct
// bucket rows and needs neither atomics nor a fixed summation tree: the result is bit-reproducible.
// Every CTA writes its partial sketch into `partials` in the split-K workspace layout [cta][sketch_tiles*128][Lc], so the
// ordinary split-K reduction (and its fused cross-GPU variant) finishes the job.  Returns the grid size used.
constexpr int kSketchZeta = 8;
int sketch_rows(int Lc);                     // s = 2 * Lc
size_t sketch_ws_bytes(int Lc, int num_sms);
cudaError_t sketch_launch(const double* Y, int64_t rows, int Lc, int64_t ld, uint64_t seed, uint64_t stream_id,
                          double* partials, int num_sms, int* grid_out, const int* cond_flag, cudaStream_t s);

// Householder QR of the s x l sketch SK (row-major, pitch ldsk) in shared memory, column-relative rank test, then the
// deflated inverse T = R^-1 (l x l upper triangular on the live columns, zero columns where dead), written as
// Lrows x ldt like chol_inv.  Y * T is then well conditioned (cond ~ 5) whatever cond(Y) was, and one plain
// CholeskyQR pass finishes the orthonormalisation.  deadmask / flag_dead / info as in chol_inv_launch.
cudaError_t hqr_inv_launch(const double* SK, int ldsk, int s_rows, int l, double* T, int Lrows, int ldt, int* info,
                           int* deadmask, int* flag_dead, const int* cond_flag, cudaStream_t s);

// ---- column statistics and rank-1 corrections for on-the-fly centring (PCA: pca_rsvd.rs:60-66) ----------------
// out[i] = sum over o < outer of p[o*ld + i], i < inner  (column sums of a row-major matrix).  Deterministic:
// per-block partial sums in `partials` (>= sum_blocks(outer) * inner doubles) are added in a fixed order.
int sum_blocks(int64_t outer);
cudaError_t sum_over_outer_launch(const double* p, int64_t inner, int64_t outer, int64_t ld, double* partials,
                                  double* out, cudaStream_t s);
// out[o] = sum over i < inner of p[o*ld + i], o < outer  (column sums of a column-major matrix).
cudaError_t sum_over_inner_launch(const double* p, int64_t inner, int64_t outer, int64_t ld, double* out,
                                  cudaStream_t s);
// v[i] *= scale
cudaError_t scale_vec_launch(double* v, int64_t n, double scale, cudaStream_t s);
// b[c] = sum_r X[r*ld + c] * mu[r], c < cols (cols <= 128)
cudaError_t gemv_t_launch(const double* X, int64_t rows, int cols, int64_t ld, const double* mu, double* b,
                          cudaStream_t s);
// Z[r*ld + c] -= u[r] * v[c]
cudaError_t rank1_sub_launch(double* Z, int64_t rows, int cols, int64_t ld, const double* u, const double* v,
                             cudaStream_t s);
// dst[o*ldd + i] = src[o*lds + i] - (mean_along_inner ? mu[i] : mu[o])   (explicit centred copy, fallback path)
cudaError_t center_copy_launch(const double* src, int64_t inner, int64_t outer, int64_t lds, double* dst, int64_t ldd,
                               const double* mu, int mu_indexed_by_inner, cudaStream_t s);

}  // namespace corrla
