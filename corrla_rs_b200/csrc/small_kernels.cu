#include "small_kernels.cuh"
#include "philox.cuh"

#include <algorithm>
#include <cfloat>
#include <cmath>

namespace corrla {

namespace {

__global__ void __launch_bounds__(256)
philox_normal_kernel(double* __restrict__ out, int64_t rows, int cols, int64_t ld, uint64_t seed) {
  const int64_t total = rows * cols;
  const int64_t npairs = (total + 1) >> 1;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npairs; p += stride) {
    double z0, z1;
    philox_normal_pair(p, seed, &z0, &z1);
    const int64_t e0 = 2 * p, e1 = e0 + 1;
    const int64_t i0 = e0 / cols;
    out[i0 * ld + (e0 - i0 * cols)] = z0;
    if (e1 < total) {
      const int64_t i1 = e1 / cols;
      out[i1 * ld + (e1 - i1 * cols)] = z1;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Cholesky + deflated triangular inverse, one CTA, matrix resident in shared memory
// ------------------------------------------------------------------------------------------------
constexpr double kTolDeadPerCol = 8.0 * DBL_EPSILON;  // times l
constexpr double kTauShift = 1e-10;

// One CTA of 16x32 threads; thread (ty, tx) keeps the entries (i, k), i = ty + 16a (a < 8), k = tx + 32b (b < 4), so
// l <= 128, of ONE packed matrix M in registers: for k >= i it is the working copy of G that becomes R (right-looking
// Cholesky G = R^T R), for k < i it is W = L^-1 (L = R^T), obtained by running the same elimination on an identity
// matrix.  T = R^-1 = W^T therefore falls out of the factorisation; there is no separate triangular inversion.
// Only pivot row j travels through shared memory (double buffered: one barrier per step).
// Deflation: a pivot under tol_dead * (original diagonal) marks column j dead: row j of R is zero, no update is made,
// and column j of T is zeroed on output, so Q = Y*T has an exact zero column there.
struct CholState {
  double m[8][4];
  double minratio;
};

__device__ __forceinline__ void chol_load(CholState& st, const double* __restrict__ G, int ldg, int l, double shift,
                                          int ty, int tx) {
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int i = ty + 16 * a, k = tx + 32 * b;
      double g = (i < l && k < l && k >= i) ? G[(int64_t)i * ldg + k] : 0.0;     // strictly-lower part: W starts at 0
      if (i == k && g > 0.0) g += shift;
      st.m[a][b] = g;
    }
  st.minratio = DBL_MAX;
}

__device__ __forceinline__ void chol_eliminate(CholState& st, int l, const double* d0, const double* d0inv, double (*rowbuf)[128],
                                               double* tdiag, int* dead_s, bool shifted, double tol_dead, int ty,
                                               int tx) {
  for (int j = 0; j < l; ++j) {
    const int buf = j & 1;
#pragma unroll
    for (int a = 0; a < 8; ++a)
      if (ty + 16 * a == j) {
#pragma unroll
        for (int b = 0; b < 4; ++b) rowbuf[buf][tx + 32 * b] = st.m[a][b];
      }
    __syncthreads();
    const double piv = rowbuf[buf][j], dj = d0[j];
    bool is_dead;
    if (shifted) is_dead = !(dj > 0.0) || !(piv > 0.0);
    else is_dead = !(dj > 0.0) || !(piv > tol_dead * dj);
    const double ratio = (dj > 0.0 && piv > 0.0) ? piv * d0inv[j] : 0.0;    // reciprocals precomputed: no fp64 divide per step
    if (ratio < st.minratio) st.minratio = ratio;
    if (ty == 0 && tx == 0) dead_s[j] = is_dead ? 1 : 0;
    if (is_dead) {
#pragma unroll
      for (int a = 0; a < 8; ++a)
        if (ty + 16 * a == j) {
#pragma unroll
          for (int b = 0; b < 4; ++b) if (tx + 32 * b >= j) st.m[a][b] = 0.0;
        }
      if (ty == 0 && tx == 0) tdiag[j] = 0.0;
      continue;
    }
    const double inv = rsqrt(piv);
    if (ty == 0 && tx == 0) tdiag[j] = inv;
    if (ty + 112 < j) continue;                         // every row this warp owns is already finished
    double ck[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int k = tx + 32 * b;
      ck[b] = (k == j) ? inv : rowbuf[buf][k] * inv;   // k > j: r_jk;  k < j: W[j][k] * inv;  k == j: W[j][j] * inv
    }
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      const int i = ty + 16 * a;
      if (i > j) {
        const double ri = rowbuf[buf][i] * inv;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int k = tx + 32 * b;
          if (k >= i || k <= j) st.m[a][b] -= ri * ck[b];
        }
      } else if (i == j) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int k = tx + 32 * b;
          st.m[a][b] = (k == j) ? piv * inv : ck[b];   // R[j][j] = sqrt(piv); R[j][k>j]; scaled W[j][k<j]
        }
      }
    }
  }
}

__global__ void __launch_bounds__(512, 1)
chol_inv_kernel(const double* __restrict__ G, int ldg, int l, double* __restrict__ T, int Lrows, int ldt, int mode,
                double global_rows, int* flag3, int* info, double* dinfo, int* deadmask, int* flag_dead,
                const int* cond_flag) {
  if (cond_flag != nullptr && *cond_flag == 0) return;
  extern __shared__ double sm[];
  __shared__ double rowbuf[2][128], d0[128], d0inv[128], tdiag[128];
  __shared__ int dead_s[128];
  __shared__ double trace_s;
  const int lp = l + 1;
  double* Wsm = sm;               // l x lp staging of the packed matrix for the transposed write
  const int tid = threadIdx.x, nt = blockDim.x;
  const int tx = tid & 31, ty = tid >> 5;
  const double tol_dead = kTolDeadPerCol * l;

  for (int j = tid; j < 128; j += nt) {
    const double g = (j < l) ? G[(int64_t)j * ldg + j] : 0.0;
    d0[j] = g;
    d0inv[j] = g > 0.0 ? 1.0 / g : 0.0;
  }
  __syncthreads();
  if (tid == 0) {
    double tr = 0.0;
    for (int j = 0; j < l; ++j) tr += d0[j] > 0.0 ? d0[j] : 0.0;
    trace_s = tr;
  }
  CholState st;
  chol_load(st, G, ldg, l, 0.0, ty, tx);
  __syncthreads();
  chol_eliminate(st, l, d0, d0inv, rowbuf, tdiag, dead_s, false, tol_dead, ty, tx);
  int shifted = 0;
  if (mode == kCholAuto) {
    if (st.minratio < kTauShift) {      // uniform: every thread tracked the same pivots
      // shifted CholeskyQR: G + s I with s = 11 (m l + l (l+1)) u ||Y||_2^2, ||Y||_2^2 <= trace(G)
      const double shift = 11.0 * (global_rows * l + (double)l * (l + 1)) * (0.5 * DBL_EPSILON) * trace_s;
      __syncthreads();
      chol_load(st, G, ldg, l, shift, ty, tx);
      chol_eliminate(st, l, d0, d0inv, rowbuf, tdiag, dead_s, true, tol_dead, ty, tx);
      shifted = 1;
    }
    if (tid == 0 && flag3 != nullptr) *flag3 = shifted;
  } else if (mode == kCholCheck) {
    // after the sketch preconditioner cond(Y) is O(1); a small pivot ratio means the embedding was unlucky (or columns
    // were deflated): ask for one more CholeskyQR pass
    if (tid == 0 && flag3 != nullptr) *flag3 = (st.minratio < 1e-3) ? 1 : 0;
  } else if (mode == kCholProbe) {
    if (tid == 0 && flag3 != nullptr) {
      const int robust = (st.minratio < 1e-8) ? 1 : 0;      // dead columns have ratio 0
      flag3[0] = robust; flag3[1] = 1 - robust;
    }
  }
  __syncthreads();
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int i = ty + 16 * a, k = tx + 32 * b;
      if (i < l && k < l) Wsm[i * lp + k] = st.m[a][b];
    }
  __syncthreads();
  for (int idx = tid; idx < Lrows * ldt; idx += nt) {
    const int r = idx / ldt, c = idx - r * ldt;
    double val = 0.0;
    if (r < l && c < l && c >= r && !dead_s[c])
      val = (c == r) ? tdiag[c] : Wsm[c * lp + r];     // T = W^T, W = L^-1
    T[idx] = val;
  }
  if (tid == 0) {
    int live = 0;
    for (int j = 0; j < l; ++j) live += dead_s[j] ? 0 : 1;
    if (info != nullptr) { info[0] = live; info[1] = shifted; }
    if (dinfo != nullptr) dinfo[0] = st.minratio;
    if (flag_dead != nullptr) *flag_dead = (live < l) ? 1 : 0;
  }
  if (deadmask != nullptr)
    for (int j = tid; j < l; j += nt) deadmask[j] = dead_s[j];
}

__global__ void __launch_bounds__(256)
refill_dead_kernel(double* __restrict__ X, int64_t rows, int l, int64_t ld, const int* __restrict__ deadmask,
                   uint64_t seed, uint64_t stream_id, const int* cond_flag) {
  if (cond_flag != nullptr && *cond_flag == 0) return;
  const int64_t total = rows * l;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int64_t i = e / l;
    const int j = (int)(e - i * l);
    if (deadmask[j] == 0) continue;
    const uint64_t ctr = ((stream_id << 40) + (uint64_t)i) * 128ull + (uint64_t)j;
    uint32_t c[4] = {(uint32_t)(ctr & 0xffffffffu), (uint32_t)(ctr >> 32), 0x52454649u /* "REFI" */, 0u};
    philox4x32_10(c, (uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32));
    const uint64_t x = ((uint64_t)c[0] | ((uint64_t)c[1] << 32)) >> 11;
    const uint64_t y = ((uint64_t)c[2] | ((uint64_t)c[3] << 32)) >> 11;
    const double u1 = ((double)x + 1.0) * 0x1.0p-53, u2 = (double)y * 0x1.0p-53;
    X[i * ld + j] = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
  }
}

// ------------------------------------------------------------------------------------------------
// One-sided (Hestenes) Jacobi SVD, one CTA.  Each half-warp owns one column pair of the current round of
// a round-robin tournament (closed-form schedule, no shared ordering state), so a round of up to 64 disjoint pairs
// costs one barrier.  Column norms are carried along and refreshed exactly at the start of every sweep, which
// leaves a single dot product per pair.  The matrix handed in is nearly upper triangular (an R factor); following
// Drmac & Veselic the rotations are applied to its transpose, which converges in fewer sweeps.
// ------------------------------------------------------------------------------------------------
constexpr int kJacobiMaxSweeps = 60;
constexpr int kJacobiLanes = 8;          // lanes per column pair (a quarter warp); each lane moves 16-byte double2s,
                                         // so one group's access is one full 128-byte shared-memory wavefront
constexpr int kJacobiThreads = 512;      // 64 quarter warps >= 64 pairs (l <= 128); larger l loops

// kAllSmem: both matrices live in shared memory (l <= 118) and every access compiles to LDS.128/STS.128; otherwise
// the generic-pointer version runs with one or both matrices in global scratch.
template <bool kAllSmem>
__global__ void __launch_bounds__(kJacobiThreads)
jacobi_svd_kernel(const double* __restrict__ Win, int ldw, int l, double* __restrict__ sigma_out,
                  double* __restrict__ Vr_out, double* __restrict__ Ur_out, int Lrows, int ldo, double* gscratch,
                  int w_smem, int v_smem, int transpose, int* info) {
  extern __shared__ __align__(16) double smj[];
  const int lp = (l + 1) & ~1;         // even pitch: columns are 16-byte aligned, row l (if any) is a zero pad
  const int h = (l + 1) >> 1;          // pairs per round; N = 2h players, player index >= l is a bye
  const int N1 = 2 * h - 1;
  double* Xc;                          // working columns (become U*Sigma)
  double* Vc;                          // accumulated rotations
  double* nrm;                         // squared column norms
  if (kAllSmem) { Xc = smj; Vc = smj + l * lp; nrm = smj + 2 * l * lp; }
  else {
    Xc = w_smem ? smj : gscratch;
    Vc = v_smem ? (smj + (w_smem ? l * lp : 0)) : (gscratch + l * lp);
    nrm = smj + (w_smem ? l * lp : 0) + (v_smem ? l * lp : 0);
  }
  int* rnk = reinterpret_cast<int*>(nrm + l);
  int* anyflag = rnk + l;
  int* bigflag = anyflag + 1;
  const int tid = threadIdx.x, nt = blockDim.x;
  constexpr int LP = kJacobiLanes;
  const int grp = tid / LP, sub = tid % LP, ngrp = nt / LP;
  const unsigned gmask = ((1u << LP) - 1u) << ((tid & 31) & ~(LP - 1));
  const int glead = (tid & 31) & ~(LP - 1);

  // squared norms of the columns of X (= rows of Win when transposed), then their descending rank: the columns are
  // laid out sorted by norm, which shortens the sweep count (de Rijk ordering)
  for (int j = grp; j < l; j += ngrp) {
    double a = 0.0;
    for (int i = sub; i < l; i += LP) {
      const double x = transpose ? Win[(int64_t)j * ldw + i] : Win[(int64_t)i * ldw + j];
      a += x * x;
    }
#pragma unroll
    for (int o = LP / 2; o > 0; o >>= 1) a += __shfl_xor_sync(gmask, a, o);
    if (sub == 0) nrm[j] = a;
  }
  for (int idx = tid; idx < Lrows * ldo; idx += nt) { Vr_out[idx] = 0.0; Ur_out[idx] = 0.0; }
  __syncthreads();
  for (int j = tid; j < l; j += nt) {
    const double sj = nrm[j];
    int r = 0;
    for (int i = 0; i < l; ++i) r += (nrm[i] > sj || (nrm[i] == sj && i < j)) ? 1 : 0;
    rnk[j] = r;
  }
  __syncthreads();
  for (int idx = tid; idx < l * lp; idx += nt) {
    const int j = idx / lp, i = idx - j * lp;                     // source column j, row i (i == l: zero pad)
    const int c = rnk[j];
    double x = 0.0;
    if (i < l) x = transpose ? Win[(int64_t)j * ldw + i] : Win[(int64_t)i * ldw + j];
    Xc[c * lp + i] = x;
    Vc[c * lp + i] = (i == j) ? 1.0 : 0.0;                        // V starts as the permutation
  }
  __syncthreads();

  const double tol = sqrt((double)l) * DBL_EPSILON;
  const double tol2 = tol * tol;
  const int lp2 = lp >> 1;                                        // double2 elements per column
  int sweeps = 0, converged = 0;
  for (; sweeps < kJacobiMaxSweeps; ++sweeps) {
    for (int j = grp; j < l; j += ngrp) {                         // exact norms once per sweep
      const double2* xj = reinterpret_cast<const double2*>(Xc + j * lp);
      double a = 0.0;
      for (int i = sub; i < lp2; i += LP) { const double2 x = xj[i]; a += x.x * x.x + x.y * x.y; }
#pragma unroll
      for (int o = LP / 2; o > 0; o >>= 1) a += __shfl_xor_sync(gmask, a, o);
      if (sub == 0) nrm[j] = a;
    }
    if (tid == 0) { *anyflag = 0; *bigflag = 0; }
    __syncthreads();
    for (int r = 0; r < N1; ++r) {
      for (int pi = grp; pi < h; pi += ngrp) {
        int p, q;
        if (pi == 0) { p = N1; q = r; }
        else { p = r + pi; if (p >= N1) p -= N1; q = r - pi; if (q < 0) q += N1; }
        if (p > q) { const int tmp = p; p = q; q = tmp; }
        if (q >= l) continue;                                     // bye
        double2* xp = reinterpret_cast<double2*>(Xc + p * lp);
        double2* xq = reinterpret_cast<double2*>(Xc + q * lp);
        double c0 = 0.0, c1 = 0.0;
#pragma unroll 4
        for (int i = sub; i < lp2; i += LP) { const double2 x = xp[i], y = xq[i]; c0 += x.x * y.x; c1 += x.y * y.y; }
        double c = c0 + c1;
#pragma unroll
        for (int o = LP / 2; o > 0; o >>= 1) c += __shfl_xor_sync(gmask, c, o);
        c = __shfl_sync(gmask, c, glead);                         // identical bits in all lanes of the group
        const double a = nrm[p], b = nrm[q];
        if (c * c > tol2 * a * b) {
          // t = sign(d) * 2c / (|d| + sqrt(d^2 + 4c^2)), d = b - a  (== sign(zeta)/(|zeta| + sqrt(1 + zeta^2)))
          // division- and sqrt-free form (two rsqrt): cos(2 theta) = |d| r, r = 1/sqrt(d^2 + 4c^2); cs = sqrt(u) with
          // u = (1 + cos 2theta)/2; sn = |c| r / cs; t = sn / cs = |c| r / u, signs from d*c
          const double d = b - a;
          const double rr = rsqrt(fma(d, d, 4.0 * c * c));
          const double u = fma(0.5 * fabs(d), rr, 0.5);
          const double icu = rsqrt(u);
          const double cr = fabs(c) * rr;
          const double cs = u * icu;
          const double sn = copysign(cr * icu, d * c);
          const double t = copysign(cr * icu * icu, d * c);
          double2* vp = reinterpret_cast<double2*>(Vc + p * lp);
          double2* vq = reinterpret_cast<double2*>(Vc + q * lp);
#pragma unroll 4
          for (int i = sub; i < lp2; i += LP) {
            const double2 x = xp[i], y = xq[i];
            xp[i] = make_double2(cs * x.x - sn * y.x, cs * x.y - sn * y.y);
            xq[i] = make_double2(sn * x.x + cs * y.x, sn * x.y + cs * y.y);
            const double2 vx = vp[i], vy = vq[i];
            vp[i] = make_double2(cs * vx.x - sn * vy.x, cs * vx.y - sn * vy.y);
            vq[i] = make_double2(sn * vx.x + cs * vy.x, sn * vx.y + cs * vy.y);
          }
          if (sub == 0) { nrm[p] = fmax(a - t * c, 0.0); nrm[q] = b + t * c; *anyflag = 1; if (c * c > kJacobiNearCos2 * a * b) *bigflag = 1; }
        }
      }
      __syncthreads();
    }
    const int any = *anyflag, big = *bigflag;
    __syncthreads();
    if (!any || !big) { converged = 1; ++sweeps; break; }   // see kJacobiNearCos2
  }

  // singular values, ranks (descending), outputs
  for (int j = grp; j < l; j += ngrp) {
    double a = 0.0;
    for (int i = sub; i < l; i += LP) { const double x = Xc[j * lp + i]; a += x * x; }
#pragma unroll
    for (int o = LP / 2; o > 0; o >>= 1) a += __shfl_xor_sync(gmask, a, o);
    if (sub == 0) nrm[j] = sqrt(a);
  }
  __syncthreads();
  for (int j = tid; j < l; j += nt) {
    const double sj = nrm[j];
    int r = 0;
    for (int i = 0; i < l; ++i) r += (nrm[i] > sj || (nrm[i] == sj && i < j)) ? 1 : 0;
    rnk[j] = r;
    sigma_out[r] = sj;
  }
  __syncthreads();
  // X = Win or Win^T;  X * Vacc = Ux * Sigma.   Win   = Ux S Vacc^T  (no transpose): Ur = Ux, Vr = Vacc
  //                                              Win^T = Ux S Vacc^T  (transpose)  : Ur = Vacc, Vr = Ux
  double* out_ux = transpose ? Vr_out : Ur_out;
  double* out_va = transpose ? Ur_out : Vr_out;
  for (int idx = tid; idx < l * l; idx += nt) {
    const int j = idx / l, i = idx - j * l;
    const int r = rnk[j];
    const double sj = nrm[j];
    out_ux[(int64_t)i * ldo + r] = sj > 0.0 ? Xc[j * lp + i] / sj : 0.0;
    out_va[(int64_t)i * ldo + r] = Vc[j * lp + i];
  }
  // Exactly zero singular values leave zero columns in Ux: complete them to an orthonormal basis (unit vectors, two
  // Gram-Schmidt passes against everything before them), as the full SVD of the reference would (random_svd.rs:89).
  __syncthreads();
  if (tid == 0) { int nzc = 0; for (int j = 0; j < l; ++j) nzc += (nrm[j] > 0.0) ? 1 : 0; *anyflag = nzc; }
  __syncthreads();
  const int nz = *anyflag;
  if (nz < l) {
    double* coef = nrm;                  // sigma is already in sigma_out; coef[r] doubles as the norm slot of column r
    int cand = 0;
    for (int r = nz; r < l; ++r) {
      for (; cand < l; ++cand) {
        __syncthreads();
        for (int i = tid; i < l; i += nt) out_ux[(int64_t)i * ldo + r] = (i == cand) ? 1.0 : 0.0;
        __syncthreads();
        for (int pass = 0; pass < 2; ++pass) {
          for (int q = grp; q < r; q += ngrp) {
            double a = 0.0;
            for (int i = sub; i < l; i += LP) a += out_ux[(int64_t)i * ldo + q] * out_ux[(int64_t)i * ldo + r];
#pragma unroll
            for (int o = LP / 2; o > 0; o >>= 1) a += __shfl_xor_sync(gmask, a, o);
            if (sub == 0) coef[q] = a;
          }
          __syncthreads();
          for (int i = tid; i < l; i += nt) {
            double a = out_ux[(int64_t)i * ldo + r];
            for (int q = 0; q < r; ++q) a -= coef[q] * out_ux[(int64_t)i * ldo + q];
            out_ux[(int64_t)i * ldo + r] = a;
          }
          __syncthreads();
        }
        if (tid < 32) {
          double a = 0.0;
          for (int i = tid; i < l; i += 32) { const double x = out_ux[(int64_t)i * ldo + r]; a += x * x; }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
          if (tid == 0) coef[r] = a;
        }
        __syncthreads();
        const double n2 = coef[r];
        if (n2 > 0.25) {
          const double inv = rsqrt(n2);
          for (int i = tid; i < l; i += nt) out_ux[(int64_t)i * ldo + r] *= inv;
          ++cand;
          break;
        }
      }
    }
  }
  if (tid == 0 && info != nullptr) { info[0] = sweeps; info[1] = converged; }
}

// ------------------------------------------------------------------------------------------------
// strided copies
// ------------------------------------------------------------------------------------------------
// tile 32x32 through shared memory; reads run along whichever source stride is smaller.
__global__ void __launch_bounds__(256)
repack_kernel(const double* __restrict__ src, int64_t rows, int64_t cols, int64_t rs, int64_t cs,
              double* __restrict__ dst, int64_t drs, int64_t dcs, double scale) {
  __shared__ double tile[32][33];
  const int64_t tiles_c = (cols + 31) / 32;
  const int64_t tiles_total = ((rows + 31) / 32) * tiles_c;
  const bool src_row_fast = llabs(cs) <= llabs(rs);   // consecutive j are closer in memory than consecutive i
  const bool dst_row_fast = llabs(dcs) <= llabs(drs);
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int64_t tIdx = blockIdx.x; tIdx < tiles_total; tIdx += gridDim.x) {
    const int64_t i0 = (tIdx / tiles_c) * 32, j0 = (tIdx % tiles_c) * 32;
    for (int r = ty; r < 32; r += 8) {
      const int64_t i = src_row_fast ? i0 + r : i0 + tx;
      const int64_t j = src_row_fast ? j0 + tx : j0 + r;
      if (i < rows && j < cols) tile[i - i0][j - j0] = scale * src[i * rs + j * cs];
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
      const int64_t i = dst_row_fast ? i0 + r : i0 + tx;
      const int64_t j = dst_row_fast ? j0 + tx : j0 + r;
      if (i < rows && j < cols) dst[i * drs + j * dcs] = tile[i - i0][j - j0];
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// sparse sign sketch
// ------------------------------------------------------------------------------------------------
constexpr int kSketchTile = 64;
__constant__ int kSketchMul[kSketchZeta] = {1, 17, 19, 23, 29, 31, 37, 41};   // coprime with 16 * nblk, nblk <= 16

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// 1024 threads: column pair cp = tid & 63 (columns 2cp, 2cp+1; idle if 2cp >= Lc), row lane rj = tid >> 6 handles the
// tile rows rj, rj + 16, rj + 32, rj + 48.  The per-(tile, hash) bucket offset and sign word are drawn by 8 threads
// one tile ahead and shared through a small table; bucket indices advance incrementally (no division in the loop).
__global__ void __launch_bounds__(1024, 1)
sketch_kernel(const double* __restrict__ Y, int64_t rows, int Lc, int64_t ld, int s_rows, uint64_t seed,
              uint64_t stream_id, double* __restrict__ partials, int rows_pad, const int* cond_flag) {
  if (cond_flag != nullptr && *cond_flag == 0) return;
  extern __shared__ __align__(16) double acc[];        // s_rows x Lc
  __shared__ int tab_off[2][kSketchZeta];
  __shared__ unsigned long long tab_sign[2][kSketchZeta];
  __shared__ int tab_step[kSketchZeta];
  const int tid = threadIdx.x;
  const int cp = tid & 63, rj = tid >> 6;
  for (int i = tid; i < s_rows * Lc; i += blockDim.x) acc[i] = 0.0;
  const int64_t tiles = (rows + kSketchTile - 1) / kSketchTile;
  const int64_t per = (tiles + gridDim.x - 1) / gridDim.x;
  const int64_t t_begin = (int64_t)blockIdx.x * per, t_end = min(tiles, t_begin + per);
  const bool active = 2 * cp < Lc;
  const uint64_t key = seed ^ (stream_id * 0xD1B54A32D192ED03ull) ^ ((uint64_t)blockIdx.x << 48);
  auto draw = [&](int64_t tile, int buf) {              // threads 0..7: hash t = tid
    const uint64_t h = splitmix64(key + (uint64_t)(tile - t_begin) * kSketchZeta + tid);
    tab_off[buf][tid] = (int)(((h >> 32) * (uint64_t)s_rows) >> 32);
    tab_sign[buf][tid] = splitmix64(h);
  };
  // per-thread constants: first bucket of my rows and the bucket stride between them, for every hash
  int base[kSketchZeta];
#pragma unroll
  for (int t = 0; t < kSketchZeta; ++t) base[t] = (kSketchMul[t] * rj) % s_rows;
  if (tid < kSketchZeta) tab_step[tid] = (kSketchMul[tid] * 16) % s_rows;
  if (tid < kSketchZeta && t_begin < t_end) draw(t_begin, 0);
  double2 cur[4], nxt[4];
  auto load_tile = [&](int64_t tile, double2 (&v)[4]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t r = tile * kSketchTile + rj + 16 * k;
      v[k] = (active && r < rows) ? *reinterpret_cast<const double2*>(Y + r * ld + 2 * cp) : make_double2(0.0, 0.0);
    }
  };
  if (t_begin < t_end) load_tile(t_begin, cur);
  __syncthreads();
  for (int64_t tile = t_begin; tile < t_end; ++tile) {
    const int buf = (int)((tile - t_begin) & 1);
    if (tile + 1 < t_end) {
      load_tile(tile + 1, nxt);
      if (tid < kSketchZeta) draw(tile + 1, buf ^ 1);
    }
#pragma unroll
    for (int t = 0; t < kSketchZeta; ++t) {
      const int off = tab_off[buf][t];
      const unsigned long long sg = tab_sign[buf][t];
      const int stp = tab_step[t];
      if (active) {
        int x[4];
        x[0] = base[t] + off; if (x[0] >= s_rows) x[0] -= s_rows;
#pragma unroll
        for (int k = 1; k < 4; ++k) { x[k] = x[k - 1] + stp; if (x[k] >= s_rows) x[k] -= s_rows; }
        double2 old[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) old[k] = *reinterpret_cast<const double2*>(acc + x[k] * Lc + 2 * cp);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const bool neg = ((sg >> (rj + 16 * k)) & 1ull) == 0ull;
          old[k].x += neg ? -cur[k].x : cur[k].x;
          old[k].y += neg ? -cur[k].y : cur[k].y;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) *reinterpret_cast<double2*>(acc + x[k] * Lc + 2 * cp) = old[k];
      }
      __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) cur[k] = nxt[k];
  }
  double* dst = partials + (int64_t)blockIdx.x * rows_pad * Lc;
  for (int i = tid; i < rows_pad * Lc; i += blockDim.x) dst[i] = (i < s_rows * Lc) ? acc[i] : 0.0;
}

// ------------------------------------------------------------------------------------------------
// Householder QR of the sketch + deflated triangular inverse, one CTA, sketch resident in shared memory
// ------------------------------------------------------------------------------------------------
constexpr int kHqrThreads = 512;   // 16 warps: column k belongs to warp k & 15 (slot k >> 4, at most 8 slots: l <= 128)
constexpr int kHqrRI = 8;          // row slots per lane: rows lane + 32*i, s_rows <= 256

// Householder reflector of column j acting on rows >= rk, by ONE warp (LAPACK dlarfg).  Leaves v (v[rk] = 1 implied)
// below row rk, beta = R[rk][j] at row rk, tau in tauv[j]; rowof[j] = rk, or -1 if the column is numerically dependent.
__device__ __forceinline__ void hqr_make_reflector(double* x, int rk, int s_rows, double cn0j, double tol, int lane,
                                                   double* tau_out, int* rowof_out) {
  double sig = 0.0;
  for (int r = rk + 1 + lane; r < s_rows; r += 32) sig += x[r] * x[r];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sig += __shfl_xor_sync(0xffffffffu, sig, o);
  const double alpha = (rk < s_rows) ? x[rk] : 0.0;
  const double nrm = sqrt(alpha * alpha + sig);
  const bool dead = (rk >= s_rows) || !(nrm > tol * cn0j) || !(cn0j > 0.0);
  double tau = 0.0;
  __syncwarp();
  if (!dead) {
    const double beta = (alpha >= 0.0) ? -nrm : nrm;
    tau = (beta - alpha) / beta;
    const double scal = 1.0 / (alpha - beta);
    for (int r = rk + 1 + lane; r < s_rows; r += 32) x[r] *= scal;
    if (lane == 0) x[rk] = beta;
  }
  if (lane == 0) { *tau_out = tau; *rowof_out = dead ? -1 : rk; }
}

__global__ void __launch_bounds__(kHqrThreads, 1)
hqr_inv_kernel(const double* __restrict__ SK, int ldsk, int s_rows, int l, double* __restrict__ T, int Lrows, int ldt,
               int* info, int* deadmask, int* flag_dead, const int* cond_flag) {
  if (cond_flag != nullptr && *cond_flag == 0) return;
  extern __shared__ double hsm[];
  const int sp = s_rows | 1;                            // odd pitch: conflict-free both along rows and across columns
  double* Ac = hsm;                                     // column-major: Ac[c*sp + r]
  double* cn0 = Ac + (size_t)l * sp;                    // original column norms
  double* vwork = cn0 + l;                              // l: column of R being inverted
  double* tauv = vwork + l;                             // l: tau of every reflector
  int* rowof = reinterpret_cast<int*>(tauv + l);        // l: R row assigned to column j, or -1 (dead)
  int* live = rowof + l;                                // compact list of live columns
  const int tid = threadIdx.x, nt = blockDim.x, warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  const double tol = 8.0 * l * DBL_EPSILON;

  for (int idx = tid; idx < s_rows * l; idx += nt) {
    const int r = idx / l, c = idx - r * l;
    Ac[c * sp + r] = SK[(int64_t)r * ldsk + c];
  }
  __syncthreads();
  for (int c = warp; c < l; c += nw) {
    double a = 0.0;
    for (int r = lane; r < s_rows; r += 32) { const double x = Ac[c * sp + r]; a += x * x; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) cn0[c] = sqrt(a);
  }
  __syncthreads();
  if (warp == 0) hqr_make_reflector(Ac, 0, s_rows, cn0[0], tol, lane, &tauv[0], &rowof[0]);
  __syncthreads();

  // ---- Householder sweeps with one barrier per column: every warp applies reflector j to the columns it owns, and the
  // owner of column j+1 builds the next reflector right after updating it.  `rk` = next free row of R (does not advance
  // on a dead column).
  int rk = 0;
  for (int j = 0; j < l; ++j) {
    const bool dead_j = rowof[j] < 0;
    const double tau = tauv[j];
    const int rk_next = dead_j ? rk : rk + 1;
    if (!dead_j) {
      double vr[kHqrRI];
      const double* v = Ac + j * sp;
#pragma unroll
      for (int i = 0; i < kHqrRI; ++i) {
        const int r = lane + 32 * i;
        vr[i] = (r < s_rows && r > rk) ? v[r] : ((r == rk) ? 1.0 : 0.0);
      }
      const int i_lo = rk >> 5;                          // row slots below this one are finished rows of R
#pragma unroll 1
      for (int g4 = 0; g4 < 2; ++g4) {                   // my columns, four at a time (independent dot products)
        double vals[4][kHqrRI];
        double dot[4];
        bool on[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int k = warp + 16 * (4 * g4 + q);
          on[q] = (k > j) && (k < l);
          dot[q] = 0.0;
          if (on[q]) {
            const double* y = Ac + k * sp;
#pragma unroll
            for (int i = 0; i < kHqrRI; ++i) {
              const int r = lane + 32 * i;
              vals[q][i] = (i >= i_lo && r < s_rows) ? y[r] : 0.0;
              dot[q] += vr[i] * vals[q][i];
            }
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
          for (int q = 0; q < 4; ++q) dot[q] += __shfl_xor_sync(0xffffffffu, dot[q], o);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (on[q]) {
            const int k = warp + 16 * (4 * g4 + q);
            double* y = Ac + k * sp;
            const double w = tau * dot[q];
#pragma unroll
            for (int i = 0; i < kHqrRI; ++i) {
              const int r = lane + 32 * i;
              if (i >= i_lo && r < s_rows && r >= rk) y[r] = vals[q][i] - w * vr[i];
            }
          }
        }
      }
    }
    if (j + 1 < l && ((j + 1) & 15) == warp) {
      __syncwarp();
      hqr_make_reflector(Ac + (j + 1) * sp, rk_next, s_rows, cn0[j + 1], tol, lane, &tauv[j + 1], &rowof[j + 1]);
    }
    __syncthreads();
    rk = rk_next;
  }
  const int rank = rk;
  if (tid == 0) {
    int n = 0;
    for (int j = 0; j < l; ++j) if (rowof[j] >= 0) live[n++] = j;
  }
  __syncthreads();

  // ---- invert the rank x rank upper-triangular R_c[a][b] = Ac[live[b]*sp + a] in place, column by column
  for (int b = 0; b < rank; ++b) {
    double* colb = Ac + live[b] * sp;
    const double tbb = 1.0 / colb[b];
    for (int a = tid; a < b; a += nt) vwork[a] = colb[a];
    __syncthreads();
    for (int a = warp; a < b; a += nw) {
      double dot = 0.0;
      for (int k = a + lane; k < b; k += 32) dot += Ac[live[k] * sp + a] * vwork[k];   // T_c[a][k], already inverted
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      if (lane == 0) colb[a] = -dot * tbb;
    }
    if (tid == 0) colb[b] = tbb;
    __syncthreads();
  }
  // T[live[a]][live[b]] = T_c[a][b], a <= b; zero elsewhere
  for (int idx = tid; idx < Lrows * ldt; idx += nt) T[idx] = 0.0;
  __syncthreads();
  for (int idx = tid; idx < rank * rank; idx += nt) {
    const int a = idx / rank, b = idx - a * rank;
    if (a <= b) T[(int64_t)live[a] * ldt + live[b]] = Ac[live[b] * sp + a];
  }
  if (tid == 0) {
    if (info != nullptr) { info[0] = rank; info[1] = 0; }
    if (flag_dead != nullptr) *flag_dead = (rank < l) ? 1 : 0;
  }
  if (deadmask != nullptr)
    for (int j = tid; j < l; j += nt) deadmask[j] = rowof[j] < 0 ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------
// column statistics / rank-1 corrections
// ------------------------------------------------------------------------------------------------
constexpr int kSumRowsPerBlock = 4096;

// block (bx over inner tiles of 128, by over row chunks): partial[by][i] = sum of rows of the chunk
__global__ void __launch_bounds__(1024)
sum_over_outer_kernel(const double* __restrict__ p, int64_t inner, int64_t outer, int64_t ld, double* __restrict__ partials) {
  __shared__ double red[8][128];
  const int tx = threadIdx.x & 127, ty = threadIdx.x >> 7;       // 128 columns x 8 row lanes
  const int64_t i = (int64_t)blockIdx.x * 128 + tx;
  const int64_t o0 = (int64_t)blockIdx.y * kSumRowsPerBlock, o1 = min(outer, o0 + kSumRowsPerBlock);
  double a = 0.0;
  if (i < inner)
    for (int64_t o = o0 + ty; o < o1; o += 8) a += p[o * ld + i];
  red[ty][tx] = a;
  __syncthreads();
  if (ty == 0 && i < inner) {
    double t = 0.0;
#pragma unroll
    for (int r = 0; r < 8; ++r) t += red[r][tx];
    partials[(int64_t)blockIdx.y * inner + i] = t;
  }
}
__global__ void __launch_bounds__(256)
sum_partials_kernel(const double* __restrict__ partials, int64_t inner, int nblocks, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= inner) return;
  double t = 0.0;
  for (int b = 0; b < nblocks; ++b) t += partials[(int64_t)b * inner + i];
  out[i] = t;
}
// one block per outer index: sum of a contiguous row
__global__ void __launch_bounds__(256)
sum_over_inner_kernel(const double* __restrict__ p, int64_t inner, int64_t outer, int64_t ld, double* __restrict__ out) {
  __shared__ double red[8];
  for (int64_t o = blockIdx.x; o < outer; o += gridDim.x) {
    double a = 0.0;
    for (int64_t i = threadIdx.x; i < inner; i += blockDim.x) a += p[o * ld + i];
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) a += __shfl_xor_sync(0xffffffffu, a, k);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += red[w];
      out[o] = t;
    }
    __syncthreads();
  }
}
__global__ void scale_vec_kernel(double* v, int64_t n, double scale) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] *= scale;
}
__global__ void __launch_bounds__(1024)
gemv_t_kernel(const double* __restrict__ X, int64_t rows, int cols, int64_t ld, const double* __restrict__ mu,
              double* __restrict__ b) {
  __shared__ double red[8][128];
  const int tx = threadIdx.x & 127, ty = threadIdx.x >> 7;
  double a = 0.0;
  if (tx < cols)
    for (int64_t r = ty; r < rows; r += 8) a += X[r * ld + tx] * mu[r];
  red[ty][tx] = a;
  __syncthreads();
  if (ty == 0 && tx < cols) {
    double t = 0.0;
#pragma unroll
    for (int r = 0; r < 8; ++r) t += red[r][tx];
    b[tx] = t;
  }
}
__global__ void __launch_bounds__(256)
rank1_sub_kernel(double* __restrict__ Z, int64_t rows, int cols, int64_t ld, const double* __restrict__ u,
                 const double* __restrict__ v) {
  const int64_t total = rows * cols;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int64_t r = e / cols;
    const int c = (int)(e - r * cols);
    Z[r * ld + c] -= u[r] * v[c];
  }
}
__global__ void __launch_bounds__(256)
center_copy_kernel(const double* __restrict__ src, int64_t inner, int64_t outer, int64_t lds, double* __restrict__ dst,
                   int64_t ldd, const double* __restrict__ mu, int by_inner) {
  const int64_t total = inner * outer;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int64_t o = e / inner, i = e - o * inner;
    dst[o * ldd + i] = src[o * lds + i] - (by_inner ? mu[i] : mu[o]);
  }
}

}  // namespace

int sketch_rows(int Lc) {
  // s = 2*Lc buckets, capped so that both the s x Lc accumulator of the sketch kernel and the l x s panel of the
  // Householder kernel fit in shared memory; always a multiple of 16 with prime factors <= 13 (coprime with kSketchMul)
  int s = std::max(2 * Lc, 64);
  const int cap = (int)((size_t)220 * 1024 / ((size_t)Lc * 8));
  if (s > cap) s = cap / 16 * 16;
  return s;
}
size_t sketch_ws_bytes(int Lc, int num_sms) {
  const int rows_pad = (sketch_rows(Lc) + 127) / 128 * 128;
  return (size_t)num_sms * rows_pad * Lc * 8;
}
cudaError_t sketch_launch(const double* Y, int64_t rows, int Lc, int64_t ld, uint64_t seed, uint64_t stream_id,
                          double* partials, int num_sms, int* grid_out, const int* cond_flag, cudaStream_t s) {
  const int s_rows = sketch_rows(Lc);
  const int rows_pad = (s_rows + 127) / 128 * 128;
  const size_t smem = (size_t)s_rows * Lc * 8;
  if (Lc > 128 || smem > 224 * 1024) return cudaErrorInvalidValue;
  {   // function attributes are per device: set on every launch (about a microsecond)
    cudaError_t e = cudaFuncSetAttribute(sketch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    if (e != cudaSuccess) return e;
  }
  const int64_t tiles = (rows + kSketchTile - 1) / kSketchTile;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(tiles, num_sms));
  sketch_kernel<<<grid, 1024, smem, s>>>(Y, rows, Lc, ld, s_rows, seed, stream_id, partials, rows_pad, cond_flag);
  if (grid_out) *grid_out = grid;
  return cudaGetLastError();
}
cudaError_t hqr_inv_launch(const double* SK, int ldsk, int s_rows, int l, double* T, int Lrows, int ldt, int* info,
                           int* deadmask, int* flag_dead, const int* cond_flag, cudaStream_t s) {
  const int sp = s_rows | 1;
  const size_t smem = ((size_t)l * sp + 3 * (size_t)l + 4) * 8 + 2 * (size_t)l * 4 + 16;
  if (smem > 226 * 1024) return cudaErrorInvalidValue;
  {   // function attributes are per device: set on every launch (about a microsecond)
    cudaError_t e = cudaFuncSetAttribute(hqr_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e != cudaSuccess) return e;
  }
  hqr_inv_kernel<<<1, kHqrThreads, smem, s>>>(SK, ldsk, s_rows, l, T, Lrows, ldt, info, deadmask, flag_dead, cond_flag);
  return cudaGetLastError();
}

int sum_blocks(int64_t outer) { return (int)((outer + kSumRowsPerBlock - 1) / kSumRowsPerBlock); }

cudaError_t sum_over_outer_launch(const double* p, int64_t inner, int64_t outer, int64_t ld, double* partials,
                                  double* out, cudaStream_t s) {
  if (inner <= 0 || outer <= 0) return cudaSuccess;
  const int nb = sum_blocks(outer);
  dim3 grid((unsigned)((inner + 127) / 128), (unsigned)nb);
  sum_over_outer_kernel<<<grid, 1024, 0, s>>>(p, inner, outer, ld, partials);
  sum_partials_kernel<<<(unsigned)((inner + 255) / 256), 256, 0, s>>>(partials, inner, nb, out);
  return cudaGetLastError();
}
cudaError_t sum_over_inner_launch(const double* p, int64_t inner, int64_t outer, int64_t ld, double* out,
                                  cudaStream_t s) {
  if (inner <= 0 || outer <= 0) return cudaSuccess;
  sum_over_inner_kernel<<<(unsigned)std::min<int64_t>(outer, 148 * 8), 256, 0, s>>>(p, inner, outer, ld, out);
  return cudaGetLastError();
}
cudaError_t scale_vec_launch(double* v, int64_t n, double scale, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  scale_vec_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(v, n, scale);
  return cudaGetLastError();
}
cudaError_t gemv_t_launch(const double* X, int64_t rows, int cols, int64_t ld, const double* mu, double* b,
                          cudaStream_t s) {
  if (cols > 128) return cudaErrorInvalidValue;
  gemv_t_kernel<<<1, 1024, 0, s>>>(X, rows, cols, ld, mu, b);
  return cudaGetLastError();
}
cudaError_t rank1_sub_launch(double* Z, int64_t rows, int cols, int64_t ld, const double* u, const double* v,
                             cudaStream_t s) {
  if (rows <= 0 || cols <= 0) return cudaSuccess;
  const int blocks = (int)std::min<int64_t>((rows * cols + 255) / 256, 148 * 8);
  rank1_sub_kernel<<<blocks, 256, 0, s>>>(Z, rows, cols, ld, u, v);
  return cudaGetLastError();
}
cudaError_t center_copy_launch(const double* src, int64_t inner, int64_t outer, int64_t lds, double* dst, int64_t ldd,
                               const double* mu, int mu_indexed_by_inner, cudaStream_t s) {
  if (inner <= 0 || outer <= 0) return cudaSuccess;
  const int blocks = (int)std::min<int64_t>((inner * outer + 255) / 256, 148 * 16);
  center_copy_kernel<<<blocks, 256, 0, s>>>(src, inner, outer, lds, dst, ldd, mu, mu_indexed_by_inner);
  return cudaGetLastError();
}

cudaError_t philox_normal_launch(double* out, int64_t rows, int cols, int64_t ld, uint64_t seed, cudaStream_t s) {
  const int64_t npairs = (rows * cols + 1) / 2;
  if (npairs <= 0) return cudaSuccess;
  const int blocks = (int)std::min<int64_t>((npairs + 255) / 256, 148 * 16);
  philox_normal_kernel<<<blocks, 256, 0, s>>>(out, rows, cols, ld, seed);
  return cudaGetLastError();
}

cudaError_t chol_inv_launch(const double* G, int ldg, int l, double* T, int Lrows, int ldt, int mode,
                            double global_rows, int* flag3, int* info, double* dinfo, int* deadmask, int* flag_dead,
                            const int* cond_flag, cudaStream_t s) {
  if (l > 128) return cudaErrorInvalidValue;
  const size_t smem = (size_t)l * (l + 1) * 8;
  {   // function attributes are per device: set on every launch (about a microsecond)
    // the kernel also has ~5 KB of static shared memory: leave room for it under the 227 KB limit
    cudaError_t e = cudaFuncSetAttribute(chol_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
  }
  chol_inv_kernel<<<1, 512, smem, s>>>(G, ldg, l, T, Lrows, ldt, mode, global_rows, flag3, info, dinfo, deadmask,
                                       flag_dead, cond_flag);
  return cudaGetLastError();
}

cudaError_t refill_dead_launch(double* X, int64_t rows, int l, int64_t ld, const int* deadmask, uint64_t seed,
                               uint64_t stream_id, const int* cond_flag, cudaStream_t s) {
  if (rows <= 0 || l <= 0) return cudaSuccess;
  const int blocks = (int)std::min<int64_t>((rows * l + 255) / 256, 148 * 16);
  refill_dead_kernel<<<blocks, 256, 0, s>>>(X, rows, l, ld, deadmask, seed, stream_id, cond_flag);
  return cudaGetLastError();
}

cudaError_t jacobi_svd_launch(const double* W, int ldw, int l, double* sigma, double* Vr, double* Ur, int Lrows,
                              int ldo, double* scratch, int* info, cudaStream_t s, int transpose) {
  {
    cudaError_t ce = jacobi_svd_ring_launch(W, ldw, l, sigma, Vr, Ur, Lrows, ldo, info, s, transpose);
    if (ce == cudaSuccess) return ce;
    if (ce != cudaErrorNotSupported) cudaGetLastError();     // launch refused: next kernel in line
    ce = jacobi_svd_cluster_launch(W, ldw, l, sigma, Vr, Ur, Lrows, ldo, info, s, transpose);
    if (ce == cudaSuccess) return ce;
    if (ce != cudaErrorNotSupported) cudaGetLastError();     // launch refused: fall back to the single-CTA kernel
  }
  const int lp = (l + 1) & ~1;
  const size_t mat = (size_t)l * lp * 8;
  const size_t small = (size_t)l * 8 + (size_t)(l + 4) * 4 + 16;
  const size_t cap = 227 * 1024;
  int w_smem = 0, v_smem = 0;
  if (2 * mat + small <= cap) { w_smem = 1; v_smem = 1; }
  else if (mat + small <= cap) { w_smem = 1; }
  const size_t smem = (w_smem + v_smem) * mat + small;
  {   // function attributes are per device: set on every launch (about a microsecond)
    cudaError_t e = cudaFuncSetAttribute(jacobi_svd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(jacobi_svd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap);
    if (e != cudaSuccess) return e;
  }
  if (w_smem && v_smem)
    jacobi_svd_kernel<true><<<1, kJacobiThreads, smem, s>>>(W, ldw, l, sigma, Vr, Ur, Lrows, ldo, scratch, 1, 1, transpose, info);
  else
    jacobi_svd_kernel<false><<<1, kJacobiThreads, smem, s>>>(W, ldw, l, sigma, Vr, Ur, Lrows, ldo, scratch, w_smem, v_smem, transpose, info);
  return cudaGetLastError();
}

// One CTA.  Also measures E = G - I: *redo = 1 when ||E||_F^2 > 4e-16, i.e. when the neglected (3/8)E^2 would show in
// the orthogonality of X*T; the caller then lets the Cholesky kernel overwrite T (launched behind this flag).
__global__ void __launch_bounds__(1024)
lowdin_kernel(const double* __restrict__ G, int ldg, int l, double* __restrict__ T, int Lrows, int ldt, int* redo) {
  __shared__ double red[32];
  double e2 = 0.0;
  for (int idx = threadIdx.x; idx < Lrows * ldt; idx += blockDim.x) {
    const int i = idx / ldt, j = idx - i * ldt;
    double v = 0.0;
    if (i < l && j < l) {
      const double g = (i <= j) ? G[(int64_t)i * ldg + j] : G[(int64_t)j * ldg + i];
      const double e = g - (i == j ? 1.0 : 0.0);
      e2 += e * e;
      v = (i == j ? 1.0 : 0.0) - 0.5 * e;
    }
    T[idx] = v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) e2 += __shfl_xor_sync(0xffffffffu, e2, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = e2;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
    if (redo != nullptr) *redo = (tot > 4e-16 || !(tot == tot)) ? 1 : 0;
  }
}

cudaError_t lowdin_launch(const double* G, int ldg, int l, double* T, int Lrows, int ldt, int* redo, cudaStream_t s) {
  lowdin_kernel<<<1, 1024, 0, s>>>(G, ldg, l, T, Lrows, ldt, redo);
  return cudaGetLastError();
}

cudaError_t repack_launch(const double* src, int64_t rows, int64_t cols, int64_t rs, int64_t cs, double* dst,
                          int64_t ldd, cudaStream_t s, double scale) {
  if (rows <= 0 || cols <= 0) return cudaSuccess;
  const int64_t tiles = ((rows + 31) / 32) * ((cols + 31) / 32);
  const int blocks = (int)std::min<int64_t>(tiles, 148 * 32);
  repack_kernel<<<blocks, 256, 0, s>>>(src, rows, cols, rs, cs, dst, ldd, 1, scale);
  return cudaGetLastError();
}

cudaError_t scatter_launch(const double* src, int64_t rows, int64_t cols, int64_t ld, double* dst, int64_t drs,
                           int64_t dcs, cudaStream_t s) {
  if (rows <= 0 || cols <= 0) return cudaSuccess;
  const int64_t tiles = ((rows + 31) / 32) * ((cols + 31) / 32);
  const int blocks = (int)std::min<int64_t>(tiles, 148 * 32);
  repack_kernel<<<blocks, 256, 0, s>>>(src, rows, cols, ld, 1, dst, drs, dcs, 1.0);
  return cudaGetLastError();
}

}  // namespace corrla
