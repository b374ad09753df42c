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
#include <cstdio>

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
  size_t oA, oY, oY2, oZ, oZ2, oX, oV, oM1, oM2, oG, oT1, oT2, oTau, oNrm, oSig, oRed, total;
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
  L.oG = o; o += (size_t)L.l8 * L.pl;       // Gram matrix / Cholesky factor of the CholeskyQR2 path
  L.oT1 = o; o += (size_t)L.l8 * L.pl;      // L^-1 (its transpose is R^-1)
  L.oT2 = o; o += (size_t)L.l8 * L.pl;      // second-pass factor 3/2 I - 1/2 G2
  L.oTau = o; o += 2 * (size_t)L.l8;      // hv0 | hden of the Householder reflectors
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

// Householder reflectors, kept UNSCALED so that no square root and no division sits on the critical path of a column
// step (an FP64 sqrt or divide is a ~500-cycle dependent chain here, and a step has nothing to hide it behind):
//   x = column j below row j,  alpha = x[j],  s = alpha^2 + |x[j+1:]|^2,  nrm = sqrt(s) = s * rsqrt(s),
//   beta = -sign(alpha) nrm = R[j][j],   v = x - beta e_j  (v[j] = alpha - beta kept in hv0[j], v[j+1:] = x[j+1:] in place),
//   H = I - v v^T / (nrm (nrm + |alpha|))      since v^T v = 2 nrm (nrm + |alpha|).
// make_reflector leaves beta on the diagonal, v0 in hv0[j] and the DENOMINATOR nrm (nrm + |alpha|) in hden[j] (0 => H = I);
// whoever applies the reflector takes the reciprocal itself, which overlaps with its own dot-product reduction.
__device__ __forceinline__ void fs_make_reflector(double* x, int j, int rows, double* hv0, double* hden, int lane) {
  double sig = 0.0;
  for (int r = j + 1 + lane; r < rows; r += 32) sig += x[r] * x[r];
  sig = fs_warp_sum(sig);
  const double alpha = x[j];
  double v0 = 0.0, den = 0.0;
  __syncwarp();
  if (sig > 0.0) {
    const double s2 = fma(alpha, alpha, sig);
    const double nrm = s2 * rsqrt(s2);
    const double beta = (alpha >= 0.0) ? -nrm : nrm;
    v0 = alpha - beta;
    den = nrm * (nrm + fabs(alpha));
    if (lane == 0) x[j] = beta;
  }
  if (lane == 0) { hv0[j] = v0; hden[j] = den; }
}

// x <- H_j x for one column x (one warp); v = column j of the factored matrix (rows > j), v0 / den as above
__device__ __forceinline__ void fs_apply_reflector(const double* v, double v0, double den, double* x, int j, int rows, int lane) {
  const double tau = 1.0 / den;                        // independent of the reduction below: the two chains overlap
  double dot = (lane == 0) ? v0 * x[j] : 0.0;
  for (int r = j + 1 + lane; r < rows; r += 32) dot += v[r] * x[r];
  dot = fs_warp_sum(dot);
  const double w = tau * dot;
  if (lane == 0) x[j] -= w * v0;
  for (int r = j + 1 + lane; r < rows; r += 32) x[r] -= w * v[r];
  __syncwarp();
}

// In-place Householder QR of the rows x l column-major matrix F (pitch p): reflectors below the diagonal, R on and above
// it.  Column c belongs to warp c % 16; the owner of column j+1 builds the next reflector right after updating that
// column, so a step costs one CTA barrier.  Replaces faer's qr() at random_svd.rs:38 / :57 for this path.
__device__ void fs_house_factor(double* F, int p, int rows, int l, double* hv0, double* hden, int warp, int lane) {
  if (warp == 0) fs_make_reflector(F, 0, rows, hv0, hden, lane);
  __syncthreads();
  for (int j = 0; j < l; ++j) {
    const double v0 = hv0[j], den = hden[j];
    const double* v = F + j * p;
    int c = j + 1 + ((warp - (j + 1)) % kFsWarps + kFsWarps) % kFsWarps;      // first column > j owned by this warp
    for (; c < l; c += kFsWarps) {
      double* x = F + c * p;
      if (den != 0.0) fs_apply_reflector(v, v0, den, x, j, rows, lane);
      if (c == j + 1) fs_make_reflector(x, j + 1, rows, hv0, hden, lane);
    }
    __syncthreads();
  }
}

// Explicit thin Q (rows x l, column-major into Qb, pitch p) from the factored matrix: column c is H_0 ... H_c e_c.
// Columns are independent: no barrier until the end.  (compute_thin_q, random_svd.rs:38 / :57)
__device__ void fs_house_form_q(const double* F, int p, int rows, int l, const double* hv0, const double* hden, double* Qb,
                                int warp, int lane) {
  for (int c = warp; c < l; c += kFsWarps) {
    double* qc = Qb + c * p;
    for (int r = lane; r < rows; r += 32) qc[r] = (r == c) ? 1.0 : 0.0;
    __syncwarp();
    for (int j = c; j >= 0; --j) {
      const double den = hden[j];
      if (den != 0.0) fs_apply_reflector(F + j * p, hv0[j], den, qc, j, rows, lane);
    }
  }
  __syncthreads();
}


// ---- CholeskyQR2 for the well-conditioned case (what the multi-kernel engine does on its fast path) -----------------
// A Householder QR of a 100 x 18 matrix is 18 dependent column steps of ~2000 cycles each on one SM; CholeskyQR2 moves
// the m-sized work into four small DMMA products and leaves one 18-step chain on an l x l matrix.

// G = F^T F (F: rows8 x l8 column-major, pitch p; G: l8 x l8 row-major, pitch pl): one warp per 8 x 8 block of G
__device__ void fs_gram(const double* F, int p, int rows8, int l8, const double* F2, double* G, int pl, int warp, int lane) {
  const int g = lane >> 2, t = lane & 3;
  const int nb = l8 >> 3, units = nb * nb, ks = rows8 >> 2;
  for (int u = warp; u < units; u += kFsWarps) {
    const int mb = u / nb, jb = u - mb * nb;
    const double* ap = F + (8 * mb + g) * p + t;            // a(i, kk) = F [i*p + kk]
    const double* bp = F2 + (8 * jb + g) * p + t;           // b(kk, j) = F2[j*p + kk]
    double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;          // two accumulator pairs: half the dependent DMMA chain
    int s = 0;
    for (; s + 1 < ks; s += 2) {
      dmma_m8n8k4(c0, c1, ap[4 * s], bp[4 * s]);
      dmma_m8n8k4(e0, e1, ap[4 * s + 4], bp[4 * s + 4]);
    }
    if (s < ks) dmma_m8n8k4(c0, c1, ap[4 * s], bp[4 * s]);
    double* o = G + (8 * mb + g) * pl + 8 * jb + 2 * t;
    o[0] = c0 + e0; o[1] = c1 + e1;
  }
}

// Warp 0: Cholesky G = R^T R in place (upper triangle, rows of R) and W = L^-1 (L = R^T) in Wm, so that R^-1 = W^T
// needs no separate triangular inversion.  Left-looking: lane k forms entry (j, k) of row j of R (k >= j) and of W (k <= j)
// as a dot product over the finished rows m < j -- the inner loop only loads (R[m][j] is one broadcast word shared by
// both sums), so nothing serialises behind shared-memory stores; one rsqrt per row is the dependent chain.
// Returns false (warp-uniform) as soon as a pivot falls under 1e-8 of its original diagonal entry: cond(F) beyond ~1e4
// or a rank-deficient F -- the Householder path takes over, nothing has been modified in F.
__device__ bool fs_chol_inv_warp(double* G, double* Wm, int pl, int l, int lane) {
  const bool col = lane < l;
  const double d0 = col ? G[lane * pl + lane] : 1.0;
  for (int j = 0; j < l; ++j) {
    double sg = (col && lane >= j) ? G[j * pl + lane] : 0.0;
    double sw = (lane == j) ? 1.0 : 0.0;
    const bool up = col && lane >= j, lo = lane <= j;
    for (int m = 0; m < j; ++m) {
      const double rmj = G[m * pl + j];
      const double rmk = up ? G[m * pl + lane] : 0.0;
      const double wmk = (lo && lane <= m) ? Wm[m * pl + lane] : 0.0;
      sg = fma(-rmj, rmk, sg);
      sw = fma(-rmj, wmk, sw);
    }
    const double piv = __shfl_sync(0xffffffffu, sg, j);
    const double dj = __shfl_sync(0xffffffffu, d0, j);
    if (!(dj > 0.0) || !(piv > 1e-8 * dj)) return false;
    const double inv = rsqrt(piv);
    if (up) G[j * pl + lane] = sg * inv;
    if (lo) Wm[j * pl + lane] = sw * inv;
    __syncwarp();
  }
  return true;
}

// F (rows x l, column-major, pitch p) <- orthonormal basis of its span by CholeskyQR2 with the first-order second pass
// (T2 = 3/2 I - 1/2 G2, exact to (3/8)|G2 - I|^2).  Returns 0: done, the basis is in F.  1: not applicable, F untouched.
// 2: the second pass found |G2 - I| too large: `Other` holds F * R^-1 (same span, condition number ~1) for Householder.
// basis_only (the QR inside the power loop, whose consumer needs range(F) only -- engine_core.cuh::qr_inplace): stop after
// the first pass and return 3: `Other` holds F * R^-1, a basis of condition 1 + cond(F)^2 eps <= 1 + 1e-8.
__device__ int fs_cholqr2(double* F, double* Other, int p, int rows8, int l, int l8, double* sG, double* sT1, double* sT2,
                          int pl, double* red, int* sflag, int warp, int lane, bool basis_only) {
  const int tid = warp * 32 + lane;
  const int nbl = l8 >> 3;
  fs_gram(F, p, rows8, l8, F, sG, pl, warp, lane);
  for (int idx = tid; idx < l8 * pl; idx += kFsThreads) sT1[idx] = 0.0;
  __syncthreads();
  if (warp == 0) {
    const bool ok = fs_chol_inv_warp(sG, sT1, pl, l, lane);
    if (lane == 0) *sflag = ok ? 1 : 0;
  }
  __syncthreads();
  if (*sflag == 0) return 1;
  // Other = F * R^-1:   b(kk, j) = R^-1[kk][j] = W[j][kk] = sT1[j*pl + kk]
  fs_gemm(nbl, F, 1, p, sT1, 1, pl, Other, 1, p, rows8 >> 3, l8 >> 2, rows8, l8, 1.0, warp, lane);
  __syncthreads();
  if (basis_only) return 3;
  fs_gram(Other, p, rows8, l8, Other, sG, pl, warp, lane);
  __syncthreads();
  double e2 = 0.0;
  for (int idx = tid; idx < l8 * pl; idx += kFsThreads) {
    const int i = idx / pl, j = idx - i * pl;
    double v = 0.0;
    if (i < l && j < l) {
      const double e = sG[idx] - (i == j ? 1.0 : 0.0);
      e2 += e * e;
      v = (i == j ? 1.0 : 0.0) - 0.5 * e;
    }
    sT2[idx] = v;
  }
  e2 = fs_warp_sum(e2);
  if (lane == 0) red[warp] = e2;
  __syncthreads();
  if (tid == 0) {
    double tot = 0.0;
    for (int w = 0; w < kFsWarps; ++w) tot += red[w];
    *sflag = (tot <= 4e-16) ? 1 : 0;                       // NaN compares false: Householder
  }
  __syncthreads();
  if (*sflag == 0) return 2;
  fs_gemm(nbl, Other, 1, p, sT2, pl, 1, F, 1, p, rows8 >> 3, l8 >> 2, rows8, l8, 1.0, warp, lane);
  __syncthreads();
  return 0;
}

// One-sided Jacobi SVD of the l x l matrix held as columns X[j*lp + i] (l <= 32) by the whole CTA: one WARP per column
// pair of the round-robin tournament (at most 16 pairs), lane i owns row i of both columns, so a round is four shared
// loads, one butterfly, the rotation parameters (the two dependent rsqrt are what a round costs) and four stores, then
// one CTA barrier.  V accumulates the rotations.  Stands in for faer's svd() at random_svd.rs:89 on the
// QR-preconditioned core.  flags: two ints of shared memory.
__device__ void fs_jacobi_cta(double* X, double* V, int lp, int l, double* nrm, int* flags, int warp, int lane,
                              int* sweeps_out, int* conv_out) {
  const int h = (l + 1) >> 1, N1 = 2 * h - 1;
  const double tol2 = (double)l * DBL_EPSILON * DBL_EPSILON;
  const bool row = lane < l;
  int sweeps = 0, converged = 0;
  for (; sweeps < kFsMaxSweeps; ++sweeps) {
    for (int j = warp; j < l; j += kFsWarps) {            // exact squared norms once per sweep
      const double x = row ? X[j * lp + lane] : 0.0;
      const double a = fs_warp_sum(x * x);
      if (lane == 0) nrm[j] = a;
    }
    if (warp == 0 && lane == 0) { flags[0] = 0; flags[1] = 0; }
    __syncthreads();
    for (int r = 0; r < N1; ++r) {
      int p = 0, q = l;
      if (warp < h) {
        if (warp == 0) { p = N1; q = r; }
        else { p = r + warp; if (p >= N1) p -= N1; q = r - warp; if (q < 0) q += N1; }
        if (p > q) { const int tmp = p; p = q; q = tmp; }
      }
      if (warp < h && q < l) {                            // warp-uniform
        double* xp = X + p * lp + lane;
        double* xq = X + q * lp + lane;
        const double x = row ? *xp : 0.0, y = row ? *xq : 0.0;
        const double c = fs_warp_sum(x * y);              // identical bits in every lane
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
          if (row) {
            double* vp = V + p * lp + lane;
            double* vq = V + q * lp + lane;
            const double vx = *vp, vy = *vq;
            *xp = cs * x - sn * y; *xq = sn * x + cs * y;
            *vp = cs * vx - sn * vy; *vq = sn * vx + cs * vy;
          }
          if (lane == 0) {
            const double t = copysign(cr * icu * icu, d * c);
            nrm[p] = fmax(a - t * c, 0.0); nrm[q] = b + t * c;
            flags[0] = 1;
            if (c * c > kJacobiNearCos2 * a * b) flags[1] = 1;
          }
        }
      }
      __syncthreads();
    }
    const int any = flags[0], big = flags[1];
    __syncthreads();                                      // everybody has read the flags before the next sweep clears them
    if (!any || !big) { converged = 1; ++sweeps; break; }   // kJacobiNearCos2: no sweep just to confirm convergence
  }
  *sweeps_out = sweeps;
  *conv_out = converged;
}

__global__ void __launch_bounds__(kFsThreads, 1)
fused_small_rsvd_kernel(const FusedSmallArgs p) {
  extern __shared__ __align__(16) double fsm[];
  __shared__ int s_perm[kFusedMaxL];
  __shared__ int s_info[8];
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
  double* sG = fsm + L.oG;
  double* sT1 = fsm + L.oT1;
  double* sT2 = fsm + L.oT2;
  double* hv0 = fsm + L.oTau;
  double* hden = hv0 + L.l8;
  double* nrm = fsm + L.oNrm;
  double* sig = fsm + L.oSig;
  double* red = fsm + L.oRed;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nbl = L.l8 / 8;
  // phase clocks (thread 0, debug only): 0 stage-in, 1 A*Z, 2 A^T*Y, 3 thin Q, 4 normalise, 5 QR of B^T, 6 Jacobi, 7 outputs
  long long tph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tlast = clock64();
  auto tick = [&](int ph) { if (p.debug && tid == 0) { const long long now = clock64(); tph[ph] += now - tlast; tlast = now; } };

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
  tick(0);

  auto mm_AZ = [&](const double* Zin, double* Yout) {          // Y = A * Z            random_svd.rs:31, :47-51
    fs_gemm(nbl, sA, L.pa, 1, Zin, 1, L.pn, Yout, 1, L.pm, L.m8 / 8, L.n8 / 4, L.m8, L.l8, 1.0, warp, lane);
    __syncthreads();
    tick(1);
  };
  auto mm_AtY = [&](const double* Yin, double* Zout) {         // Z = A^T * Y          :42-46, :80
    fs_gemm(nbl, sA, 1, L.pa, Yin, 1, L.pm, Zout, 1, L.pn, L.n8 / 8, L.m8 / 4, L.n8, L.l8, 1.0, warp, lane);
    __syncthreads();
    tick(2);
  };
  // F <- thin Q of F: CholeskyQR2 when its Cholesky probe passes, else Householder (the buffers may swap)
  auto thin_q = [&](double*& F, double*& Other, int pitch, int rows, int rows8, bool basis_only) {
    const int st = p.no_chol ? 1 : fs_cholqr2(F, Other, pitch, rows8, l, L.l8, sG, sT1, sT2, L.pl, red, s_info + 6, warp, lane,
                                              basis_only);
    if (st == 3) { double* tmp = F; F = Other; Other = tmp; }
    else if (st != 0) {
      if (st == 2) { double* tmp = F; F = Other; Other = tmp; }
      fs_house_factor(F, pitch, rows, l, hv0, hden, warp, lane);
      fs_house_form_q(F, pitch, rows, l, hv0, hden, Other, warp, lane);
      double* tmp = F; F = Other; Other = tmp;
    }
    tick(3);
  };

  mm_AZ(sZ, sY);
  for (int it = 0; it < p.n_iter; ++it) {                      // :35
    if (p.schedule == 1 || it > 2) thin_q(sY, sY2, L.pm, m, L.m8, p.basis_only != 0);   // :37-39
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
    tick(4);
  }
  thin_q(sY, sY2, L.pm, m, L.m8, false);                       // :57   sY = Q

  if (p.power_only) {
    for (int idx = tid; idx < m * l; idx += kFsThreads) {
      const int c = idx / m, r = idx - c * m;
      p.qout[(int64_t)c * m + r] = sY[c * L.pm + r];
    }
    if (tid == 0) { p.info[0] = 0; p.info[1] = 1; p.info[2] = 0; }
    return;
  }

  mm_AtY(sY, sZ);                                              // :80   sZ = B^T (n x l)
  // SVD of B (:89): B^T = Qz W with Qz the thin Q of B^T and W = Qz^T B^T (l x l, a small product: no R bookkeeping,
  // whichever QR produced Qz), one-sided Jacobi on W^T, then U = Q * Vr, V = Qz * Ur
  for (int idx = tid; idx < L.l8 * L.pn; idx += kFsThreads) sY2[idx] = sZ[idx];     // keep B^T (sY2 is free: sY holds Q)
  __syncthreads();
  thin_q(sZ, sZ2, L.pn, n, L.n8, false);                       // sZ = Qz
  {
    double* Qz = sZ;
    // W[i][j] = sum_kk Qz[kk][i] * Bt[kk][j]  ->  X column i, row j  (X = W^T as columns: X[i*lp + j] = W[i][j])
    const int g = lane >> 2, t = lane & 3;
    const int nb = L.l8 >> 3, ks = L.n8 >> 2;
    for (int u = warp; u < nb * nb; u += kFsWarps) {
      const int mb = u / nb, jb = u - mb * nb;
      const double* ap = Qz + (8 * mb + g) * L.pn + t;
      const double* bp = sY2 + (8 * jb + g) * L.pn + t;
      double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
      int s2 = 0;
      for (; s2 + 1 < ks; s2 += 2) {
        dmma_m8n8k4(c0, c1, ap[4 * s2], bp[4 * s2]);
        dmma_m8n8k4(e0, e1, ap[4 * s2 + 4], bp[4 * s2 + 4]);
      }
      if (s2 < ks) dmma_m8n8k4(c0, c1, ap[4 * s2], bp[4 * s2]);
      const int i = 8 * mb + g, j = 8 * jb + 2 * t;
      if (i < l) {
        if (j < l) sX[i * L.lp + j] = c0 + e0;
        if (j + 1 < l) sX[i * L.lp + j + 1] = c1 + e1;
      }
    }
    for (int idx = tid; idx < l * l; idx += kFsThreads) {
      const int j = idx / l, i = idx - j * l;
      sV[j * L.lp + i] = (i == j) ? 1.0 : 0.0;
    }
    __syncthreads();
  }
  tick(5);
  int sweeps, conv;
  fs_jacobi_cta(sX, sV, L.lp, l, nrm, s_info + 4, warp, lane, &sweeps, &conv);
  if (warp == 0) {
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
  tick(6);
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
    fs_gemm(nbk, sZ, 1, L.pn, sM2, 1, L.pl, p.v, p.v_rs, p.v_cs, L.n8 / 8, L.l8 / 4, n, k, 1.0, warp, lane);
  for (int c = tid; c < k; c += kFsThreads) p.s[c] = sig[c];
  if (tid == 0) { p.info[0] = s_info[0]; p.info[1] = s_info[1]; p.info[2] = s_info[2]; }
  tick(7);
  if (p.debug && tid == 0)
    printf("fused_small m=%d n=%d l=%d q=%d cycles: stage-in %lld | A*Z %lld | A^T*Y %lld | thin-Q %lld | normalise %lld | QR(B^T) %lld | "
           "Jacobi %lld (%d sweeps) | outputs %lld\n", m, n, l, p.n_iter, tph[0], tph[1], tph[2], tph[3], tph[4], tph[5], tph[6],
           s_info[0], tph[7]);
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
