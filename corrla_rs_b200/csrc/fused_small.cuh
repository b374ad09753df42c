// Whole-call RSVD of a TINY matrix in ONE kernel: the thin matrix (m x n, a few hundred KB at most) lives in the shared
// memory of a single CTA together with Y, Z and the l x l factors, and the complete schedule of the reference
// (random_svd.rs:15-110: sketch, n_iter power iterations with the QR gate i > 2 and the Frobenius scaling, final thin Q,
// B = Q^T A, SVD of B, U = Q * Ub) runs without leaving the SM: every product on the FP64 tensor pipe (DMMA.8x8x4 fed from
// shared memory), thin-Q by Householder reflectors (the reference's own QR: no Gram matrix, no rank decisions, an
// orthonormal completion for rank-deficient Y for free), the SVD of the l x l core by one-sided Jacobi in one warp.
// The multi-kernel engine needs ~190 launches and 10 host round trips for BASELINE config C1 (100 x 100, n_iters = 12):
// it is launch-latency bound at 1.7 ms, slower than the CPU.  This path is what small callers get instead.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace corrla {

constexpr int kFusedMaxL = 32;            // sketch columns
constexpr size_t kFusedMaxSmem = 224 * 1024;

struct FusedSmallArgs {
  const double* a; int64_t a_rs, a_cs;    // thin matrix (m x n, m >= n) with element strides, device memory
  int m, n, l, k, n_iter, schedule, power_only;
  const double* omega; int64_t om_rs, om_cs;   // n x l injected test matrix (device) or nullptr => Philox(seed)
  uint64_t seed;
  double* u; int64_t u_rs, u_cs;          // thin-U  m x k (may be nullptr)
  double* v; int64_t v_rs, v_cs;          // thin-V  n x k
  double* s;                              // k singular values
  double* qout;                           // power_only: Q, m x l column-major
  int no_chol;                            // 1: Householder QR only (CORRLA_B200_FUSED_NO_CHOL=1; tests compare the two)
  int basis_only;                         // 1: the in-loop QR stops after one Cholesky pass (0: CORRLA_B200_INLOOP_CHOLQR2=1)
  int debug;                              // 1: thread 0 prints a clock64 phase breakdown (CORRLA_B200_FUSED_PROFILE=1)
  int* info;                              // [0] Jacobi sweeps, [1] converged, [2] 1 => result unusable, take the general path
};

// Bytes of dynamic shared memory the kernel needs for this problem, or 0 when it does not apply.
size_t fused_small_smem_bytes(int m, int n, int l, int k);
cudaError_t fused_small_launch(const FusedSmallArgs& args, cudaStream_t stream);

}  // namespace corrla
