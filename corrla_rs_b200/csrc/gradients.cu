#include "gradients.cuh"

#include <algorithm>
#include <cfloat>

namespace corrla {

namespace {

// ------------------------------------------------------------------------------------------------
// exact brute-force k nearest neighbours
// ------------------------------------------------------------------------------------------------
constexpr int kQT = 64;           // queries per CTA: 8 warps x 8 queries
constexpr int kCT = 128;          // candidates per tile: 4 per lane
constexpr int kDC = 32;           // features per staged chunk
constexpr int kKnnThreads = 256;
constexpr int kQPitch = kQT + 1, kCPitch = kCT + 1;   // odd pitches: the transposing stores are conflict free

// Warp-collective insertion of the lanes flagged in `mask` (lowest lane first = increasing candidate index) into the
// sorted list (ld, li) of length k; returns the new k-th best distance.  Equal distances keep the lower index first.
__device__ __noinline__ double knn_insert(double* ld, int* li, int k, double dist, int cand, double thr, unsigned mask,
                                          int lane) {
  while (mask) {
    const int src = __ffs(mask) - 1;
    const double dv = __shfl_sync(0xffffffffu, dist, src);
    const int iv = __shfl_sync(0xffffffffu, cand, src);
    if (dv < thr) {
      int cnt = 0;
      for (int p = lane; p < k; p += 32) cnt += (ld[p] <= dv) ? 1 : 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      const int pos = cnt;
      double td[4]; int ti[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int p = lane + 32 * j;
        td[j] = 0.0; ti[j] = 0;
        if (p > pos && p < k) { td[j] = ld[p - 1]; ti[j] = li[p - 1]; }
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int p = lane + 32 * j;
        if (p > pos && p < k) { ld[p] = td[j]; li[p] = ti[j]; }
        else if (p == pos && p < k) { ld[p] = dv; li[p] = iv; }
      }
      __syncwarp();
      thr = ld[k - 1];
    }
    mask &= mask - 1;
    mask &= __ballot_sync(0xffffffffu, dist < thr);
  }
  return thr;
}

__global__ void __launch_bounds__(kKnnThreads)
knn_kernel(const double* __restrict__ X, int64_t n, int d, int64_t ldx, int k, int* __restrict__ idx_out) {
  extern __shared__ __align__(16) double smk[];
  double* Qs = smk;                                   // [kDC][kQPitch]  query chunk, feature-major
  double* Cs = Qs + kDC * kQPitch;                    // [kDC][kCPitch]  candidate chunk, feature-major
  double* Ld = Cs + kDC * kCPitch;                    // [kQT][k]        sorted distances of the current best k
  int* Li = reinterpret_cast<int*>(Ld + (size_t)kQT * k);   // [kQT][k]  their indices
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t q0 = (int64_t)blockIdx.x * kQT;
  for (int i = tid; i < kQT * k; i += kKnnThreads) { Ld[i] = DBL_MAX; Li[i] = -1; }

  for (int64_t c0 = 0; c0 < n; c0 += kCT) {
    double acc[8][4];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
    for (int d0 = 0; d0 < d; d0 += kDC) {
      __syncthreads();
      for (int i = tid; i < kDC * kQT; i += kKnnThreads) {
        const int q = i / kDC, dd = i - q * kDC;
        const int64_t row = q0 + q;
        Qs[dd * kQPitch + q] = (row < n && d0 + dd < d) ? X[row * ldx + d0 + dd] : 0.0;
      }
      for (int i = tid; i < kDC * kCT; i += kKnnThreads) {
        const int cc = i / kDC, dd = i - cc * kDC;
        const int64_t row = c0 + cc;
        Cs[dd * kCPitch + cc] = (row < n && d0 + dd < d) ? X[row * ldx + d0 + dd] : 0.0;
      }
      __syncthreads();
#pragma unroll 4
      for (int dd = 0; dd < kDC; ++dd) {
        double cv[4], qv[8];
#pragma unroll
        for (int b = 0; b < 4; ++b) cv[b] = Cs[dd * kCPitch + lane + 32 * b];
#pragma unroll
        for (int a = 0; a < 8; ++a) qv[a] = Qs[dd * kQPitch + 8 * warp + a];
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) { const double t = qv[a] - cv[b]; acc[a][b] = fma(t, t, acc[a][b]); }
      }
    }
    // selection: warp w owns queries 8w .. 8w+7; candidates are visited in increasing index order.  Only the threshold
    // test is unrolled; the (rare) insertion is one out-of-line routine, which keeps the hot loop in the instruction cache
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      const int ql = 8 * warp + a;
      if (q0 + ql < n) {                                              // uniform in the warp
        double* ld = Ld + (size_t)ql * k;
        int* li = Li + (size_t)ql * k;
        double thr = ld[k - 1];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int64_t cg = c0 + lane + 32 * b;
          const double dist = (cg < n) ? acc[a][b] : DBL_MAX;
          const unsigned mask = __ballot_sync(0xffffffffu, dist < thr);
          if (mask) thr = knn_insert(ld, li, k, dist, (int)cg, thr, mask, lane);
        }
      }
    }
  }
  __syncwarp();
  for (int a = 0; a < 8; ++a) {
    const int ql = 8 * warp + a;
    const int64_t row = q0 + ql;
    if (row >= n) break;
    for (int p = lane; p < k; p += 32) idx_out[row * k + p] = Li[(size_t)ql * k + p];
  }
}

// ------------------------------------------------------------------------------------------------
// batched local least squares: one CTA per sample
// ------------------------------------------------------------------------------------------------
constexpr int kGradThreads = 128;

__global__ void __launch_bounds__(kGradThreads)
poly_grad_kernel(const double* __restrict__ X, const double* __restrict__ y, int64_t n, int d, int64_t ldx,
                 const int* __restrict__ idx, int k, int order, int p, double* __restrict__ G, int64_t ldg,
                 int* __restrict__ info) {
  extern __shared__ __align__(16) double smg[];
  const int kp = k | 1;                              // odd column pitch: thread-per-column accesses are conflict free
  double* A = smg;                                   // [p + 1][kp] column-major: p design columns, then y
  double* diagR = A + (size_t)(p + 1) * kp;          // [p]
  double* beta = diagR + p;                          // [p]
  double* scal = beta + p;                           // [4] sigma, alpha, tau, max |R_jj|
  const int tid = threadIdx.x, nt = blockDim.x;
  const int64_t i = blockIdx.x;
  const int* nb = idx + i * k;

  for (int e = tid; e < k * d; e += nt) {
    const int r = e / d, j = e - r * d;
    A[(size_t)j * kp + r] = X[(int64_t)nb[r] * ldx + j];
  }
  for (int r = tid; r < k; r += nt) A[(size_t)p * kp + r] = y[nb[r]];
  __syncthreads();
  if (order == 2) {
    // interaction columns in the order of mat_col_interactions (stats_corr.rs:112-142): (a, b) with a <= b
    for (int c = d + tid; c < p; c += nt) {
      int a = 0, rem = c - d;
      while (rem >= d - a) { rem -= d - a; ++a; }
      const int b = a + rem;
      for (int r = 0; r < k; ++r) A[(size_t)c * kp + r] = A[(size_t)a * kp + r] * A[(size_t)b * kp + r];
    }
    __syncthreads();
  }
  // centre every column (the slopes of a least-squares fit with an intercept are those of the centred problem)
  for (int c = tid; c <= p; c += nt) {
    double m = 0.0;
    for (int r = 0; r < k; ++r) m += A[(size_t)c * kp + r];
    m /= (double)k;
    for (int r = 0; r < k; ++r) A[(size_t)c * kp + r] -= m;
  }
  __syncthreads();
  // Householder QR of [design | y]
  const int steps = min(p, k - 1);
  for (int j = 0; j < steps; ++j) {
    if (tid < 32) {
      double s = 0.0;
      for (int r = j + tid; r < k; r += 32) { const double v = A[(size_t)j * kp + r]; s += v * v; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (tid == 0) {
        const double ajj = A[(size_t)j * kp + j];
        const double nrm = sqrt(s);
        const double alpha = (ajj > 0.0) ? -nrm : nrm;
        const double vtv = s - 2.0 * ajj * alpha + alpha * alpha;     // |a - alpha e|^2
        scal[1] = alpha;
        scal[2] = (vtv > 0.0) ? 2.0 / vtv : 0.0;
        A[(size_t)j * kp + j] = ajj - alpha;                          // v in place
        diagR[j] = alpha;
      }
    }
    __syncthreads();
    const double tau = scal[2];
    for (int c = j + 1 + tid; c <= p; c += nt) {
      double dot = 0.0;
      for (int r = j; r < k; ++r) dot += A[(size_t)j * kp + r] * A[(size_t)c * kp + r];
      dot *= tau;
      for (int r = j; r < k; ++r) A[(size_t)c * kp + r] -= dot * A[(size_t)j * kp + r];
    }
    __syncthreads();
  }
  if (tid == 0) {
    double mx = 0.0;
    for (int j = 0; j < steps; ++j) mx = fmax(mx, fabs(diagR[j]));
    scal[3] = mx;
  }
  for (int j = tid; j < p; j += nt) beta[j] = 0.0;
  __syncthreads();
  // back substitution R beta = Q^T y (rows 0 .. steps-1 of the transformed y column)
  const double tol = 1e-12 * scal[3];
  int deficient = (steps < p) ? 1 : 0;
  for (int j = steps - 1; j >= 0; --j) {
    const double rjj = diagR[j];
    const bool live = fabs(rjj) > tol;
    if (!live) deficient = 1;
    const double bj = live ? A[(size_t)p * kp + j] / rjj : 0.0;
    if (tid == 0) beta[j] = bj;
    for (int r = tid; r < j; r += nt) A[(size_t)p * kp + r] -= A[(size_t)j * kp + r] * bj;   // R_rj lives in column j, row r < j
    __syncthreads();
  }
  if (tid == 0 && deficient && info != nullptr) atomicAdd(info, 1);
  // gradient at the sample itself
  for (int j = tid; j < d; j += nt) {
    double g = beta[j];
    if (order == 2) {
      int c = d;
      for (int a = 0; a < d; ++a)
        for (int b = a; b < d; ++b, ++c) {
          if (a == j && b == j) g += 2.0 * beta[c] * X[i * ldx + j];
          else if (a == j) g += beta[c] * X[i * ldx + b];
          else if (b == j) g += beta[c] * X[i * ldx + a];
        }
    }
    G[i * ldg + j] = g;
  }
}

}  // namespace

cudaError_t knn_launch(const double* X, int64_t n, int d, int64_t ldx, int k, int* idx, cudaStream_t s) {
  if (n <= 0 || d <= 0 || k <= 0 || k > kKnnMaxK || k > n) return cudaErrorInvalidValue;
  const size_t smem = ((size_t)kDC * kQPitch + (size_t)kDC * kCPitch + (size_t)kQT * k) * 8 + (size_t)kQT * k * 4;
  cudaError_t e = cudaFuncSetAttribute(knn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) return e;
  const int64_t blocks = (n + kQT - 1) / kQT;
  knn_kernel<<<(unsigned)blocks, kKnnThreads, smem, s>>>(X, n, d, ldx, k, idx);
  return cudaGetLastError();
}

int poly_grad_num_coef(int d, int order) { return order == 2 ? d + d * (d + 1) / 2 : d; }

size_t poly_grad_smem_bytes(int d, int k, int order) {
  const int p = poly_grad_num_coef(d, order);
  return ((size_t)(p + 1) * (k | 1) + 2 * (size_t)p + 4) * 8;
}

cudaError_t poly_grad_launch(const double* X, const double* y, int64_t n, int d, int64_t ldx, const int* idx, int k,
                             int order, double* G, int64_t ldg, int* info, cudaStream_t s) {
  if (n <= 0 || d <= 0 || k <= 1 || (order != 1 && order != 2)) return cudaErrorInvalidValue;
  const int p = poly_grad_num_coef(d, order);
  const size_t smem = poly_grad_smem_bytes(d, k, order);
  if (p > kGradMaxCoef || smem > 200 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(poly_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) return e;
  poly_grad_kernel<<<(unsigned)n, kGradThreads, smem, s>>>(X, y, n, d, ldx, idx, k, order, p, G, ldg, info);
  return cudaGetLastError();
}

}  // namespace corrla
