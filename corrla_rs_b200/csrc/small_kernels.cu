#include "small_kernels.cuh"

#include <algorithm>
#include <cfloat>
#include <cmath>

namespace corrla {

namespace {

constexpr unsigned kFull = 0xffffffffu;

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11) + Box-Muller
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

__global__ void __launch_bounds__(256)
philox_normal_kernel(double* __restrict__ out, int64_t rows, int cols, int64_t ld, uint64_t seed) {
  const int64_t total = rows * cols;
  const int64_t npairs = (total + 1) >> 1;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npairs; p += stride) {
    uint32_t c[4] = {(uint32_t)(p & 0xffffffffu), (uint32_t)((uint64_t)p >> 32), 0u, 0u};
    philox4x32_10(c, (uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32));
    const uint64_t x = ((uint64_t)c[0] | ((uint64_t)c[1] << 32)) >> 11;
    const uint64_t y = ((uint64_t)c[2] | ((uint64_t)c[3] << 32)) >> 11;
    const double u1 = ((double)x + 1.0) * 0x1.0p-53;   // (0, 1]
    const double u2 = (double)y * 0x1.0p-53;           // [0, 1)
    const double r = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    const int64_t e0 = 2 * p, e1 = e0 + 1;
    const int64_t i0 = e0 / cols;
    out[i0 * ld + (e0 - i0 * cols)] = r * cs;
    if (e1 < total) {
      const int64_t i1 = e1 / cols;
      out[i1 * ld + (e1 - i1 * cols)] = r * sn;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Cholesky + deflated triangular inverse, one CTA, matrix resident in shared memory
// ------------------------------------------------------------------------------------------------
constexpr double kTolDeadPerCol = 8.0 * DBL_EPSILON;  // times l
constexpr double kTauShift = 1e-10;

// Factor S (upper triangle, pitch lp) in place; the diagonal of R goes to diag[] (S[j][j] keeps the pivot), so no
// thread overwrites what another still reads and a step needs two barriers.  Smallest pivot ratio -> *minratio_s.
// dead[j] = 1 where the pivot vanished; that row of R is zero.
__device__ void chol_factor(double* S, int lp, int l, const double* d0, double* diag, int* dead, bool shifted,
                            double tol_dead, double* minratio_s) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int tx = tid & 31, ty = tid >> 5, nty = nt >> 5;
  if (tid == 0) *minratio_s = DBL_MAX;
  __syncthreads();
  for (int j = 0; j < l; ++j) {
    const double piv = S[j * lp + j];
    const double dj = d0[j];
    bool is_dead;
    if (shifted) is_dead = !(dj > 0.0) || !(piv > 0.0);
    else is_dead = !(dj > 0.0) || !(piv > tol_dead * dj);
    if (tid == 0) {
      const double ratio = (dj > 0.0 && piv > 0.0) ? piv / dj : 0.0;
      if (ratio < *minratio_s) *minratio_s = ratio;
      dead[j] = is_dead ? 1 : 0;
      diag[j] = is_dead ? 0.0 : sqrt(piv);
    }
    if (!is_dead) {
      const double inv = 1.0 / sqrt(piv);
      for (int k = j + 1 + tid; k < l; k += nt) S[j * lp + k] *= inv;
    } else {
      for (int k = j + 1 + tid; k < l; k += nt) S[j * lp + k] = 0.0;
    }
    __syncthreads();
    if (!is_dead) {
      const double* rj = S + j * lp;
      for (int i = j + 1 + ty; i < l; i += nty) {
        const double ri = rj[i];
        double* si = S + i * lp;
        for (int k = (i & ~31) + tx; k < l; k += 32)
          if (k >= i) si[k] -= ri * rj[k];
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(512)
chol_inv_kernel(const double* __restrict__ G, int ldg, int l, double* __restrict__ T, int Lrows, int ldt, int mode,
                double global_rows, int* flag3, int* info, double* dinfo, int* deadmask, int* flag_dead,
                const int* cond_flag) {
  if (cond_flag != nullptr && *cond_flag == 0) return;
  extern __shared__ double sm[];
  const int lp = l + 1;
  double* S = sm;                 // l x lp
  double* d0 = S + l * lp;        // l
  double* v = d0 + l;             // l
  double* diag = v + l;           // l: diagonal of R
  double* scal = diag + l;        // [0] minratio, [1] trace
  int* dead = reinterpret_cast<int*>(scal + 2);
  const int tid = threadIdx.x, nt = blockDim.x, warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  const double tol_dead = kTolDeadPerCol * l;

  for (int idx = tid; idx < l * l; idx += nt) {
    const int i = idx / l, j = idx - i * l;
    S[i * lp + j] = G[(int64_t)i * ldg + j];
  }
  for (int j = tid; j < l; j += nt) d0[j] = G[(int64_t)j * ldg + j];
  __syncthreads();
  if (tid == 0) {
    double tr = 0.0;
    for (int j = 0; j < l; ++j) tr += d0[j] > 0.0 ? d0[j] : 0.0;
    scal[1] = tr;
  }
  __syncthreads();

  chol_factor(S, lp, l, d0, diag, dead, false, tol_dead, &scal[0]);
  int shifted = 0;
  if (mode == kCholAuto) {
    const double mr = scal[0];
    __syncthreads();
    if (mr < kTauShift) {
      // shifted CholeskyQR: G + s I with s = 11 (m l + l (l+1)) u ||Y||_2^2, ||Y||_2^2 <= trace(G)
      const double shift = 11.0 * (global_rows * l + (double)l * (l + 1)) * (0.5 * DBL_EPSILON) * scal[1];
      for (int idx = tid; idx < l * l; idx += nt) {
        const int i = idx / l, j = idx - i * l;
        double g = G[(int64_t)i * ldg + j];
        if (i == j && g > 0.0) g += shift;
        S[i * lp + j] = g;
      }
      __syncthreads();
      chol_factor(S, lp, l, d0, diag, dead, true, tol_dead, &scal[0]);
      shifted = 1;
    }
    if (tid == 0 && flag3 != nullptr) *flag3 = shifted;
  }

  // In-place inverse of the upper-triangular factor (dead diagonal entries act as 1), column by column:
  // T[0:j, j] = -T[0:j,0:j] * R[0:j, j] * T[j][j]
  for (int j = 0; j < l; ++j) {
    const double tjj = dead[j] ? 1.0 : 1.0 / diag[j];
    for (int k = tid; k < j; k += nt) v[k] = S[k * lp + j];
    __syncthreads();
    for (int i = warp; i < j; i += nw) {
      double dot = 0.0;
      for (int k = i + lane; k < j; k += 32) dot += S[i * lp + k] * v[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(kFull, dot, o);
      if (lane == 0) S[i * lp + j] = -dot * tjj;
    }
    if (tid == 0) S[j * lp + j] = tjj;
    __syncthreads();
  }
  for (int idx = tid; idx < Lrows * ldt; idx += nt) {
    const int i = idx / ldt, c = idx - i * ldt;
    double val = 0.0;
    if (i < l && c < l && c >= i && !dead[c]) val = S[i * lp + c];
    T[idx] = val;
  }
  if (tid == 0) {
    int live = 0;
    for (int j = 0; j < l; ++j) live += dead[j] ? 0 : 1;
    if (info != nullptr) { info[0] = live; info[1] = shifted; }
    if (dinfo != nullptr) dinfo[0] = scal[0];
    if (flag_dead != nullptr) *flag_dead = (live < l) ? 1 : 0;
  }
  if (deadmask != nullptr)
    for (int j = tid; j < l; j += nt) deadmask[j] = dead[j];
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

__global__ void __launch_bounds__(1024)
jacobi_svd_kernel(const double* __restrict__ Win, int ldw, int l, double* __restrict__ sigma_out,
                  double* __restrict__ Vr_out, double* __restrict__ Ur_out, int Lrows, int ldo, double* gscratch,
                  int w_smem, int v_smem, int transpose, int* info) {
  extern __shared__ double sm[];
  const int lp = l | 1;
  const int h = (l + 1) >> 1;          // pairs per round; N = 2h players, player index >= l is a bye
  const int N1 = 2 * h - 1;
  double* Xc = w_smem ? sm : gscratch;                                   // working columns (become U*Sigma)
  double* Vc = v_smem ? (sm + (w_smem ? l * lp : 0)) : (gscratch + l * lp);   // accumulated rotations
  double* nrm = sm + (w_smem ? l * lp : 0) + (v_smem ? l * lp : 0);      // squared column norms
  int* rnk = reinterpret_cast<int*>(nrm + l);
  const int tid = threadIdx.x, nt = blockDim.x;
  const int hw = tid >> 4, sub = tid & 15, nhw = nt >> 4;
  const unsigned hmask = 0xffffu << (tid & 16);

  for (int idx = tid; idx < l * l; idx += nt) {
    const int i = idx / l, j = idx - i * l;
    const double w = Win[(int64_t)i * ldw + j];
    if (transpose) Xc[i * lp + j] = w; else Xc[j * lp + i] = w;
    Vc[j * lp + i] = (i == j) ? 1.0 : 0.0;
  }
  for (int idx = tid; idx < Lrows * ldo; idx += nt) { Vr_out[idx] = 0.0; Ur_out[idx] = 0.0; }
  __syncthreads();

  const double tol = sqrt((double)l) * DBL_EPSILON;
  int sweeps = 0, converged = 0;
  for (; sweeps < kJacobiMaxSweeps; ++sweeps) {
    // exact norms
    for (int j = hw; j < l; j += nhw) {
      double a = 0.0;
      for (int i = sub; i < l; i += 16) { const double x = Xc[j * lp + i]; a += x * x; }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) a += __shfl_xor_sync(hmask, a, o);
      if (sub == 0) nrm[j] = a;
    }
    __syncthreads();
    int any = 0;
    for (int r = 0; r < N1; ++r) {
      int rotated = 0;
      for (int pi = hw; pi < h; pi += nhw) {
        int p, q;
        if (pi == 0) { p = N1; q = r; }
        else { p = r + pi; if (p >= N1) p -= N1; q = r - pi; if (q < 0) q += N1; }
        if (p > q) { const int tmp = p; p = q; q = tmp; }
        if (q >= l) continue;                                     // bye
        double* xp = Xc + p * lp; double* xq = Xc + q * lp;
        double c = 0.0;
        for (int i = sub; i < l; i += 16) c += xp[i] * xq[i];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) c += __shfl_xor_sync(hmask, c, o);
        c = __shfl_sync(hmask, c, tid & 16);                      // identical bits in all 16 lanes
        const double a = nrm[p], b = nrm[q];
        if (c != 0.0 && fabs(c) > tol * sqrt(a) * sqrt(b)) {
          const double zeta = (b - a) / (2.0 * c);
          const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
          double* vp = Vc + p * lp; double* vq = Vc + q * lp;
          for (int i = sub; i < l; i += 16) {
            const double x = xp[i], y = xq[i];
            xp[i] = cs * x - sn * y; xq[i] = sn * x + cs * y;
            const double vx = vp[i], vy = vq[i];
            vp[i] = cs * vx - sn * vy; vq[i] = sn * vx + cs * vy;
          }
          if (sub == 0) { nrm[p] = fmax(a - t * c, 0.0); nrm[q] = b + t * c; }
          rotated = 1;
        }
      }
      any |= __syncthreads_or(rotated);
    }
    if (!any) { converged = 1; ++sweeps; break; }
  }

  // singular values, ranks (descending), outputs
  for (int j = hw; j < l; j += nhw) {
    double a = 0.0;
    for (int i = sub; i < l; i += 16) { const double x = Xc[j * lp + i]; a += x * x; }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) a += __shfl_xor_sync(hmask, a, o);
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

}  // namespace

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
  const size_t smem = ((size_t)l * (l + 1) + 3 * (size_t)l + 2) * 8 + (size_t)l * 4;
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(chol_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
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
  const int lp = l | 1;
  const size_t mat = (size_t)l * lp * 8;
  const size_t small = (size_t)l * 8 + (size_t)(l + 2) * 4 + 16;
  const size_t cap = 227 * 1024;
  int w_smem = 0, v_smem = 0;
  if (2 * mat + small <= cap) { w_smem = 1; v_smem = 1; }
  else if (mat + small <= cap) { w_smem = 1; }
  const size_t smem = (w_smem + v_smem) * mat + small;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(jacobi_svd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  jacobi_svd_kernel<<<1, 1024, smem, s>>>(W, ldw, l, sigma, Vr, Ur, Lrows, ldo, scratch, w_smem, v_smem, transpose, info);
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
