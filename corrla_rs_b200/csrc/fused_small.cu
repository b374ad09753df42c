// Single-CTA fused RSVD for tiny matrices; see fused_small.cuh.
//
// Shared-memory layout (doubles).  Every matrix has a pitch == 4 (mod 8): with that pitch the DMMA fragment loads
//   A operand  lane (g, t) reads  a(8i + g, 4s + t)      B operand  lane (g, t) reads  b(4s + t, 8j + g)
// hit 16 distinct 8-byte banks per half-warp whichever of the two indices runs along the pitch, so the SAME copy of
// A serves A*Z and A^T*Y, and Y / Z can be kept column-major (a column is contiguous: what the Householder sweeps want).
//   sA   [m8][pa]   thin matrix, row-major, zero padded to multiples of 8        pa = n8 + 4
//   sY   [l8][pm]   Y / Q, column-major (two buffers: reflectors and the explicit Q) pm = m8 + 4
//   sZ   [l8][pn]   Z / Omega / B^T, column-major (two buffers)                   pn = n8 + 4
//   sX, sV [l][lp]  Jacobi working columns and accumulated rotations
//   sM1, sM2 [k8][pl]  the two l x k right-hand factors of the output products    pl = l8 + 4
#include "fused_small.cuh"

#include <cfloat>

#include "philox.cuh"
#include "ptx.cuh"
#include "small_kernels.cuh"

namespace corrla {

namespace {

constexpr int kFsThreads = 512;
constexpr int kFsWarps = kFsThreads / 32;
constexpr int kFsMaxSweeps = 60;

struct FsLayout {
  int m8, n8, l8, k8, pa, pm, pn, pl, lp;
  size_t oA, oY, oY2, oZ, oZ2, oX, oV, oM1, oM2, oTau, oNrm, oSig, oRed, total;
};

__host__ __device__ inline int fs_up8(int x) { return (x + 7) & ~7; }

__host__ __device__ inline FsLayout fs_layout(int m, int n, int l, int k) {
  FsLayout L;
  L.m8 = fs_up8(m); L.n8 = fs_up8(n); L.l8 = fs_up8(l); L.k8 = fs_up8(k);
  L.pa = L.n8 + 4; L.pm = L.m8 + 4; L.pn = L.n8 + 4; L.pl = L.l8 + 4; L.lp = L.l8 + 1;
  size_t o = 0;
  L.oA = o; o += (size_t)L.m8 * L.pa;
  L.oY = o; o += (size_t)L.l8 * L.pm;
  L.oY2 = o; o += (size_t)L.l8 * L.pm;
  L.oZ = o; o += (size_t)L.l8 * L.pn;
  L.oZ2 = o; o += (size_t)L.l8 * L.pn;
  L.oX = o; o += (size_t)L.l8 * L.lp;
  L.oV = o; o += (size_t)L.l8 * L.lp;
  L.oM1 = o; o += (size_t)L.l8 * L.pl;
  L.oM2 = o; o += (size_t)L.l8 * L.pl;
  L.oTau = o; o += (size_t)L.l8;
  L.oNrm = o; o += (size_t)L.l8;
  L.oSig = o; o += (size_t)L.l8;
  L.oRed = o; o += 64;
  L.total = o;
  return L;
}

// C = alpha * A * B on the FP64 tensor pipe, operands in shared memory:
//   a(i, kk) = A[i*ars + kk*acs]   (i < 8*Mblk, kk < 4*Ksteps, zero padded)
//   b(kk, j) = B[kk*brs + j*bcs]   (j < 8*NB)
//   c(i, j) -> C[i*crs + j*ccs] for i < Mvalid, j < Nvalid
// One warp owns an 8-row strip of C: one A fragment feeds NB DMMAs.
template <int NB>
__device__ __forceinline__ void fs_gemm_nb(const double* A, int ars, int acs, const double* B, int brs, int bcs, double* C,
                                           int64_t crs, int64_t ccs, int Mblk, int Ksteps, int Mvalid, int Nvalid,
                                           double alpha, int warp, int lane) {
  const int g = lane >> 2, t = lane & 3;
  for (int mb = warp; mb < Mblk; mb += kFsWarps) {
    double acc[NB][2];
#pragma unroll
    for (int j = 0; j < NB; ++j) { acc[j][0] = 0.0; acc[j][1] = 0.0; }
    const double* ap = A + (8 * mb + g) * ars + t * acs;
    const double* bp = B + t * brs + g * bcs;
    for (int s = 0; s < Ksteps; ++s) {
      const double af = ap[4 * s * acs];
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const double bf = bp[4 * s * brs + 8 * j * bcs];
        dmma_m8n8k4(acc[j][0], acc[j][1], af, bf);
      }
    }
    const int row = 8 * mb + g;
    if (row < Mvalid) {
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const int col = 8 * j + 2 * t;
        if (col < Nvalid) C[row * crs + col * ccs] = alpha * acc[j][0];
        if (col + 1 < Nvalid) C[row * crs + (col + 1) * ccs] = alpha * acc[j][1];
      }
    }
  }
}

__device__ __forceinline__ void fs_gemm(int nb, const double* A, int ars, int acs, const double* B, int brs, int bcs,
                                        double* C, int64_t crs, int64_t ccs, int Mblk, int Ksteps, int Mvalid, int Nvalid,
                                        double alpha, int warp, int lane) {
  switch (nb) {
    case 1: fs_gemm_nb<1>(A, ars, acs, B, brs, bcs, C, crs, ccs, Mblk, Ksteps, Mvalid, Nvalid, alpha, warp, lane); break;
    case 2: fs_gemm_nb<2>(A, ars, acs, B, brs, bcs, C, crs, ccs, Mblk, Ksteps, Mvalid, Nvalid, alpha, warp, lane); break;
    case 3: fs_gemm_nb<3>(A, ars, acs, B, brs, bcs, C, crs, ccs, Mblk, Ksteps, Mvalid, Nvalid, alpha, warp, lane); break;
    default: fs_gemm_nb<4>(A, ars, acs, B, brs, bcs, C, crs, ccs, Mblk, Ksteps, Mvalid, Nvalid, alpha, warp, lane); break;
  }
}

__device__ __forceinline__ double fs_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Householder reflector of column j of Y (rows j..rows-1), LAPACK dlarfg, by ONE warp: v (v[j] = 1 implied) is left
// below the diagonal, beta = R[j][j] on it, tau[j] in shared memory (0 => H = I).
__device__ __forceinline__ void fs_make_reflector(double* x, int j, int rows, double* tau, int lane) {
  double sig = 0.0;
  for (int r = j + 1 + lane; r < rows; r += 32) sig += x[r] * x[r];
  sig = fs_warp_sum(sig);
  const double alpha = x[j];
  double tj = 0.0;
  __syncwarp();
  if (sig > 0.0) {
    const double nrm = sqrt(alpha * alpha + sig);
    const double beta = (alpha >= 0.0) ? -nrm : nrm;
    tj = (beta - alpha) / beta;
    const double scal = 1.0 / (alpha - beta);
    for (int r = j + 1 + lane; r < rows; r += 32) x[r] *= scal;
    if (lane == 0) x[j] = beta;
  }
  if (lane == 0) tau[j] = tj;
}

// x <- H_j x for one column x (one warp); v = column j of the factored matrix
__device__ __forceinline__ void fs_apply_reflector(const double* v, double tj, double* x, int j, int rows, int lane) {
  double dot = (lane == 0) ? x[j] : 0.0;
  for (int r = j + 1 + lane; r < rows; r += 32) dot += v[r] * x[r];
  dot = fs_warp_sum(dot);
  const double w = tj * dot;
  if (lane == 0) x[j] -= w;
  for (int r = j + 1 + lane; r < rows; r += 32) x[r] -= w * v[r];
  __syncwarp();
}

// In-place Householder QR of the rows x l column-major matrix F (pitch p): reflectors below the diagonal, R on and above
// it.  Column c belongs to warp c % 16; the owner of column j+1 builds the next reflector right after updating that
// column, so a step costs one CTA barrier.  Replaces faer's qr() at random_svd.rs:38 / :57 for this path.
__device__ void fs_house_factor(double* F, int p, int rows, int l, double* tau, int warp, int lane) {
  if (warp == 0) fs_make_reflector(F, 0, rows, tau, lane);
  __syncthreads();
  for (int j = 0; j < l; ++j) {
    const double tj = tau[j];
    const double* v = F + j * p;
    int c = j + 1 + ((warp - (j + 1)) % kFsWarps + kFsWarps) % kFsWarps;      // first column > j owned by this warp
    for (; c < l; c += kFsWarps) {
      double* x = F + c * p;
      if (tj != 0.0) fs_apply_reflector(v, tj, x, j, rows, lane);
      if (c == j + 1) fs_make_reflector(x, j + 1, rows, tau, lane);
    }
    __syncthreads();
  }
}

// Explicit thin Q (rows x l, column-major into Qb, pitch p) from the factored matrix: column c is H_0 ... H_c e_c.
// Columns are independent: no barrier until the end.  (compute_thin_q, random_svd.rs:38 / :57)
__device__ void fs_house_form_q(const double* F, int p, int rows, int l, const double* tau, double* Qb, int warp, int lane) {
  for (int c = warp; c < l; c += kFsWarps) {
    double* qc = Qb + c * p;
    for (int r = lane; r < rows; r += 32) qc[r] = (r == c) ? 1.0 : 0.0;
    __syncwarp();
    for (int j = c; j >= 0; --j) {
      const double tj = tau[j];
      if (tj != 0.0) fs_apply_reflector(F + j * p, tj, qc, j, rows, lane);
    }
  }
  __syncthreads();
}

// One-sided Jacobi SVD of the l x l matrix held as columns X[j*lp + i] (l <= 32), by ONE warp: two lanes per column pair
// (each owns half of the rows), round-robin tournament, no barrier beyond __syncwarp.  V accumulates the rotations.
// Stands in for faer's svd() at random_svd.rs:89 on the QR-preconditioned core.
__device__ void fs_jacobi_warp(double* X, double* V, int lp, int l, double* nrm, int lane, int* sweeps_out, int* conv_out) {
  const int h = (l + 1) >> 1, N1 = 2 * h - 1;
  const int pi = lane >> 1, half = lane & 1;
  const int rsplit = (l + 1) >> 1;
  const int r0 = half ? rsplit : 0, r1 = half ? l : rsplit;
  const double tol2 = (double)l * DBL_EPSILON * DBL_EPSILON;
  int sweeps = 0, converged = 0;
  for (; sweeps < kFsMaxSweeps; ++sweeps) {
    for (int j = lane; j < l; j += 32) {
      double a = 0.0;
      for (int i = 0; i < l; ++i) { const double x = X[j * lp + i]; a += x * x; }
      nrm[j] = a;
    }
    __syncwarp();
    int any = 0, big = 0;
    for (int r = 0; r < N1; ++r) {
      int p = 0, q = l;
      if (pi < h) {
        if (pi == 0) { p = N1; q = r; }
        else { p = r + pi; if (p >= N1) p -= N1; q = r - pi; if (q < 0) q += N1; }
        if (p > q) { const int tmp = p; p = q; q = tmp; }
      }
      const bool valid = (pi < h) && (q < l);
      double c = 0.0;
      if (valid) {
        const double* xp = X + p * lp;
        const double* xq = X + q * lp;
        for (int i = r0; i < r1; ++i) c += xp[i] * xq[i];
      }
      c += __shfl_xor_sync(0xffffffffu, c, 1);          // both lanes of the pair now hold the same bits
      if (valid) {
        const double a = nrm[p], b = nrm[q];
        if (c * c > tol2 * a * b) {
          // division-free rotation (see jacobi_svd_kernel): cos(2 theta) = |d| r, r = 1/sqrt(d^2 + 4c^2)
          const double d = b - a;
          const double rr = rsqrt(fma(d, d, 4.0 * c * c));
          const double uu = fma(0.5 * fabs(d), rr, 0.5);
          const double icu = rsqrt(uu);
          const double cr = fabs(c) * rr;
          const double cs = uu * icu;
          const double sn = copysign(cr * icu, d * c);
          const double t = copysign(cr * icu * icu, d * c);
          double* xp = X + p * lp;
          double* xq = X + q * lp;
          double* vp = V + p * lp;
          double* vq = V + q * lp;
          for (int i = r0; i < r1; ++i) {
            const double x = xp[i], y = xq[i];
            xp[i] = cs * x - sn * y; xq[i] = sn * x + cs * y;
            const double vx = vp[i], vy = vq[i];
            vp[i] = cs * vx - sn * vy; vq[i] = sn * vx + cs * vy;
          }
          if (half == 0) { nrm[p] = fmax(a - t * c, 0.0); nrm[q] = b + t * c; }
          any = 1;
          if (c * c > kJacobiNearCos2 * a * b) big = 1;
        }
      }
      __syncwarp();
    }
    if (!__any_sync(0xffffffffu, any) || !__any_sync(0xffffffffu, big)) { converged = 1; ++sweeps; break; }   // kJacobiNearCos2
  }
  *sweeps_out = sweeps;
  *conv_out = converged;
}

__global__ void __launch_bounds__(kFsThreads, 1)
fused_small_rsvd_kernel(const FusedSmallArgs p) {
  extern __shared__ __align__(16) double fsm[];
  __shared__ int s_perm[kFusedMaxL];
  __shared__ int s_info[4];
  const int m = p.m, n = p.n, l = p.l, k = p.k;
  const FsLayout L = fs_layout(m, n, l, k);
  double* sA = fsm + L.oA;
  double* sY = fsm + L.oY;
  double* sY2 = fsm + L.oY2;
  double* sZ = fsm + L.oZ;
  double* sZ2 = fsm + L.oZ2;
  double* sX = fsm + L.oX;
  double* sV = fsm + L.oV;
  double* sM1 = fsm + L.oM1;
  double* sM2 = fsm + L.oM2;
  double* tau = fsm + L.oTau;
  double* nrm = fsm + L.oNrm;
  double* sig = fsm + L.oSig;
  double* red = fsm + L.oRed;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nbl = L.l8 / 8;

  // ---- stage A (any strides) and Omega; everything else starts as zeros (the pads must stay zero)
  for (size_t idx = tid; idx < L.total - L.oY; idx += kFsThreads) fsm[L.oY + idx] = 0.0;
  for (int idx = tid; idx < L.m8 * L.pa; idx += kFsThreads) {
    const int i = idx / L.pa, j = idx - i * L.pa;
    sA[idx] = (i < m && j < n) ? p.a[(int64_t)i * p.a_rs + (int64_t)j * p.a_cs] : 0.0;
  }
  __syncthreads();
  if (p.omega != nullptr) {
    for (int idx = tid; idx < n * l; idx += kFsThreads) {
      const int i = idx / l, j = idx - i * l;
      sZ[j * L.pn + i] = p.omega[(int64_t)i * p.om_rs + (int64_t)j * p.om_cs];
    }
  } else {
    // the engine's generator: pair pp -> flat elements 2pp, 2pp+1 of the row-major n x l matrix   (random_svd.rs:24)
    const int total = n * l, npairs = (total + 1) >> 1;
    for (int pp = tid; pp < npairs; pp += kFsThreads) {
      double z0, z1;
      philox_normal_pair(pp, p.seed, &z0, &z1);
      const int e0 = 2 * pp, e1 = e0 + 1;
      sZ[(e0 % l) * L.pn + e0 / l] = z0;
      if (e1 < total) sZ[(e1 % l) * L.pn + e1 / l] = z1;
    }
  }
  __syncthreads();

  auto mm_AZ = [&](const double* Zin, double* Yout) {          // Y = A * Z            random_svd.rs:31, :47-51
    fs_gemm(nbl, sA, L.pa, 1, Zin, 1, L.pn, Yout, 1, L.pm, L.m8 / 8, L.n8 / 4, L.m8, L.l8, 1.0, warp, lane);
    __syncthreads();
  };
  auto mm_AtY = [&](const double* Yin, double* Zout) {         // Z = A^T * Y          :42-46, :80
    fs_gemm(nbl, sA, 1, L.pa, Yin, 1, L.pm, Zout, 1, L.pn, L.n8 / 8, L.m8 / 4, L.n8, L.l8, 1.0, warp, lane);
    __syncthreads();
  };
  auto thin_q = [&](double*& F, double*& Other, int pitch, int rows) {    // F <- thin Q of F (buffers swap)
    fs_house_factor(F, pitch, rows, l, tau, warp, lane);
    fs_house_form_q(F, pitch, rows, l, tau, Other, warp, lane);
    double* tmp = F; F = Other; Other = tmp;
  };

  mm_AZ(sZ, sY);
  for (int it = 0; it < p.n_iter; ++it) {                      // :35
    if (p.schedule == 1 || it > 2) thin_q(sY, sY2, L.pm, m);   // :37-39
    mm_AtY(sY, sZ);
    mm_AZ(sZ, sY);
    // Y <- Y / ||Y||_F                                         :53-55
    double ss = 0.0;
    for (int idx = tid; idx < l * L.pm; idx += kFsThreads) { const double y = sY[idx]; ss += y * y; }
    ss = fs_warp_sum(ss);
    if (lane == 0) red[warp] = ss;
    __syncthreads();
    if (tid == 0) {
      double tot = 0.0;
      for (int w = 0; w < kFsWarps; ++w) tot += red[w];
      red[32] = tot > 0.0 ? 1.0 / sqrt(tot) : 0.0;
    }
    __syncthreads();
    const double sc = red[32];
    for (int idx = tid; idx < l * L.pm; idx += kFsThreads) sY[idx] *= sc;
    __syncthreads();
  }
  thin_q(sY, sY2, L.pm, m);                                    // :57   sY = Q

  if (p.power_only) {
    for (int idx = tid; idx < m * l; idx += kFsThreads) {
      const int c = idx / m, r = idx - c * m;
      p.qout[(int64_t)c * m + r] = sY[c * L.pm + r];
    }
    if (tid == 0) { p.info[0] = 0; p.info[1] = 1; p.info[2] = 0; }
    return;
  }

  mm_AtY(sY, sZ);                                              // :80   sZ = B^T (n x l)
  // SVD of B (:89): B^T = Qz R (Householder), one-sided Jacobi on R^T, then U = Q * Vr, V = Qz * Ur
  fs_house_factor(sZ, L.pn, n, l, tau, warp, lane);
  for (int idx = tid; idx < l * l; idx += kFsThreads) {
    const int j = idx / l, i = idx - j * l;                    // X column j = row j of R:  X[j][i] = R[j][i], i >= j
    sX[j * L.lp + i] = (i >= j) ? sZ[i * L.pn + j] : 0.0;
    sV[j * L.lp + i] = (i == j) ? 1.0 : 0.0;
  }
  fs_house_form_q(sZ, L.pn, n, l, tau, sZ2, warp, lane);       // sZ2 = Qz   (ends with a barrier)
  if (warp == 0) {
    int sweeps, conv;
    fs_jacobi_warp(sX, sV, L.lp, l, nrm, lane, &sweeps, &conv);
    // singular values, descending order
    for (int j = lane; j < l; j += 32) {
      double a = 0.0;
      for (int i = 0; i < l; ++i) { const double x = sX[j * L.lp + i]; a += x * x; }
      nrm[j] = sqrt(a);
    }
    __syncwarp();
    int bad = 0;
    for (int j = lane; j < l; j += 32) {
      const double sj = nrm[j];
      int r = 0;
      for (int i = 0; i < l; ++i) r += (nrm[i] > sj || (nrm[i] == sj && i < j)) ? 1 : 0;
      s_perm[r] = j;
      sig[r] = sj;
      if (r < k && !(sj > 0.0)) bad = 1;                       // exact zero (or NaN) among the kept values: no direction
    }
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) { s_info[0] = sweeps; s_info[1] = conv; s_info[2] = bad ? 1 : 0; }
  }
  __syncthreads();
  // R^T * Vacc = Ux * Sigma:  R = Vacc Sigma Ux^T  =>  Ur = Vacc, Vr = Ux;  M1 = Vr[:, :k] (for U), M2 = Ur[:, :k] (for V)
  for (int idx = tid; idx < k * l; idx += kFsThreads) {
    const int c = idx / l, i = idx - c * l;
    const int j = s_perm[c];
    const double sj = sig[c];
    sM1[c * L.pl + i] = sj > 0.0 ? sX[j * L.lp + i] / sj : 0.0;
    sM2[c * L.pl + i] = sV[j * L.lp + i];
  }
  __syncthreads();
  const int nbk = L.k8 / 8;
  if (p.u != nullptr)                                          // :92   U = Q * Ub[:, :k]
    fs_gemm(nbk, sY, 1, L.pm, sM1, 1, L.pl, p.u, p.u_rs, p.u_cs, L.m8 / 8, L.l8 / 4, m, k, 1.0, warp, lane);
  if (p.v != nullptr)
    fs_gemm(nbk, sZ2, 1, L.pn, sM2, 1, L.pl, p.v, p.v_rs, p.v_cs, L.n8 / 8, L.l8 / 4, n, k, 1.0, warp, lane);
  for (int c = tid; c < k; c += kFsThreads) p.s[c] = sig[c];
  if (tid == 0) { p.info[0] = s_info[0]; p.info[1] = s_info[1]; p.info[2] = s_info[2]; }
}

}  // namespace

size_t fused_small_smem_bytes(int m, int n, int l, int k) {
  if (m <= 0 || n <= 0 || l <= 0 || k <= 0 || l > kFusedMaxL || k > l || n > m || l > n) return 0;
  if ((int64_t)m * n > ((int64_t)1 << 16)) return 0;
  const FsLayout L = fs_layout(m, n, l, k);
  const size_t bytes = L.total * sizeof(double);
  return bytes <= kFusedMaxSmem ? bytes : 0;
}

cudaError_t fused_small_launch(const FusedSmallArgs& a, cudaStream_t stream) {
  const size_t smem = fused_small_smem_bytes(a.m, a.n, a.l, a.power_only ? a.l : a.k);
  if (smem == 0) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(fused_small_rsvd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedMaxSmem);
  if (e != cudaSuccess) return e;
  fused_small_rsvd_kernel<<<1, kFsThreads, smem, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace corrla
