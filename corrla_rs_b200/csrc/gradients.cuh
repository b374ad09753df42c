// Gradient-matrix build for active subspaces: the reference walks a kd-tree and takes one
// pseudo-inverse per sample, serially (active_subspaces.rs:66-141, :226-238).  Here: one exact brute-force
// k-nearest-neighbour kernel over all samples, and one batched local least-squares kernel (a CTA per sample).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace corrla {

constexpr int kKnnMaxK = 128;        // neighbours per sample the selection lists hold
constexpr int kGradMaxCoef = 136;    // fitted coefficients per sample (without the intercept)

// idx[i*k + r] = index of the r-th nearest row of X to row i (squared Euclidean distance accumulated as sum (a-b)^2 in
// the order of the features, like kdtree::distance::squared_euclidean; ties go to the lower index; the sample itself is
// its own nearest neighbour).  X: n x d row-major with pitch ldx.  k <= min(n, kKnnMaxK).
// With `scratch` (knn_scratch_bytes(n, k, d) bytes of device memory) the search runs in GEMM form on the tensor cores --
// shortlist by |q|^2 + |c|^2 - 2 q.c, exact FP64 re-rank, certificate -- and only uncertified queries (counted in
// *n_exact_fallback; -1 = the exact kernel did the whole job) are redone by the exact kernel: the result is the same
// neighbour lists either way.  The shortlist pass runs on TF32 mma.sync with FP32 accumulation over a TF32-rounded copy
// of the samples (default), or, with CORRLA_B200_KNN_TF32=0 and the engine's packed layout (ldx == 4 mod 8), on the FP64
// tensor pipe; the certificate's error bound follows the precision.  CORRLA_B200_KNN_EXACT=1 forces the exact kernel.
size_t knn_scratch_bytes(int64_t n, int k, int d);
cudaError_t knn_launch(const double* X, int64_t n, int d, int64_t ldx, int k, int* idx, void* scratch, size_t scratch_bytes,
                       int* n_exact_fallback, cudaStream_t s);

// The same search for query points that are not samples (PolyGradientEstimator::grad_at at an arbitrary x0,
// active_subspaces.rs:66-141): idx[q*k + r] = index of the r-th nearest row of X to row q of Q (nq x d, pitch ldq).
// Exact kernel only (query batches are small next to the sample set).
cudaError_t knn_query_launch(const double* X, int64_t n, int d, int64_t ldx, const double* Q, int64_t nq, int64_t ldq, int k,
                             int* idx, cudaStream_t s);

// Local polynomial fit through the k neighbours of every sample and its gradient at the sample:
//   order 1: y ~ b.x + b0                      gradient = b                     (jac_from_lin, stats_corr.rs:164-169)
//   order 2: y ~ b.x + sum_{a<=b} c_ab x_a x_b + b0, gradient taken analytically at x_i (jac_from_quad differentiates
//            the same polynomial by a forward difference with eps = 1e-10, stats_corr.rs:230-249)
// Least squares by Householder QR of the column-centred design matrix (same slopes as the fit with an intercept).
// With Xq != nullptr, n counts the rows of Xq (pitch ldq), idx holds THEIR neighbours and the gradient of fit i is taken
// at row i of Xq instead of at sample i.
// G: n x ldg row-major, row i = gradient at sample i (d entries).  info[0] counts samples whose design matrix was
// numerically rank deficient (their dependent coefficients are set to zero).
size_t poly_grad_smem_bytes(int d, int k, int order);
int poly_grad_num_coef(int d, int order);
cudaError_t poly_grad_launch(const double* X, const double* y, int64_t n, int d, int64_t ldx, const int* idx, int k,
                             int order, double* G, int64_t ldg, int* info, cudaStream_t s, const double* Xq = nullptr,
                             int64_t ldq = 0);

}  // namespace corrla
