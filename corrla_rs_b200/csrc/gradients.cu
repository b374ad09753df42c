#include "gradients.cuh"
#include "ptx.cuh"

#include <algorithm>
#include <cfloat>
#include <cstdlib>

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
constexpr int kKnnSlots = 5;      // list entries per lane in the insertion routine: lists of up to 160 entries
constexpr int kKnnMargin = 16;    // extra shortlist entries of the GEMM-form search (see knn_gemm_kernel)

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
      double td[kKnnSlots]; int ti[kKnnSlots];
#pragma unroll
      for (int j = 0; j < kKnnSlots; ++j) {
        const int p = lane + 32 * j;
        td[j] = 0.0; ti[j] = 0;
        if (p > pos && p < k) { td[j] = ld[p - 1]; ti[j] = li[p - 1]; }
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < kKnnSlots; ++j) {
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
knn_kernel(const double* __restrict__ X, int64_t n, int d, int64_t ldx, int k, int* __restrict__ idx_out,
           const int* __restrict__ qlist, const int* __restrict__ qcount, const double* __restrict__ Qext, int64_t nq_ext,
           int64_t ldq) {
  extern __shared__ __align__(16) double smk[];
  double* Qs = smk;                                   // [kDC][kQPitch]  query chunk, feature-major
  double* Cs = Qs + kDC * kQPitch;                    // [kDC][kCPitch]  candidate chunk, feature-major
  double* Ld = Cs + kDC * kCPitch;                    // [kQT][k]        sorted distances of the current best k
  int* Li = reinterpret_cast<int*>(Ld + (size_t)kQT * k);   // [kQT][k]  their indices
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t q0 = (int64_t)blockIdx.x * kQT;
  // with a query list (the queries the GEMM-form search could not certify) block b takes entries 64b .. 64b + 63 of it
  // with external query points (grad_at away from the samples) the queries are the rows of Qext, the candidates stay X
  const double* __restrict__ Qsrc = (Qext != nullptr) ? Qext : X;
  const int64_t qrows = (Qext != nullptr) ? nq_ext : n, ldqs = (Qext != nullptr) ? ldq : ldx;
  const int64_t nq = (qlist != nullptr) ? (int64_t)*qcount : qrows;
  if (q0 >= nq) return;
  auto qrow = [&](int ql) -> int64_t {
    const int64_t e = q0 + ql;
    if (e >= nq) return qrows;                             // past the end: behaves like a row beyond the matrix
    return (qlist != nullptr) ? (int64_t)qlist[e] : e;
  };
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
        const int64_t row = qrow(q);
        Qs[dd * kQPitch + q] = (row < qrows && d0 + dd < d) ? Qsrc[row * ldqs + d0 + dd] : 0.0;
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
      if (q0 + ql < nq) {                                             // uniform in the warp
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
    const int64_t row = qrow(ql);
    if (row >= qrows) break;
    for (int p = lane; p < k; p += 32) idx_out[row * k + p] = Li[(size_t)ql * k + p];
  }
}


// ------------------------------------------------------------------------------------------------
// GEMM-form nearest-neighbour search on the FP64 tensor pipe
// ------------------------------------------------------------------------------------------------
// |q - c|^2 = |q|^2 + |c|^2 - 2 q.c : the N^2 d inner products are a GEMM, done with DMMA.8x8x4 at one FMA per feature
// pair where the exact kernel above spends a subtraction and an FMA on the plain FP64 pipe.  Those distances carry a
// rounding error of order d eps (|q|^2 + |c|^2), so this pass only builds a SHORTLIST of k + kKnnMargin candidates per
// query; knn_rerank_kernel then recomputes the exact sum (a-b)^2 of the shortlist in feature order (the kd-tree's
// squared_euclidean), sorts it with the exact tie rule, and certifies the result: every candidate left out has an
// approximate distance >= the shortlist's last one, hence an exact distance >= that - delta; if the exact k-th distance
// is below that bound the k neighbours are provably the exact ones.  Queries that cannot be certified (more than
// kKnnMargin near-ties at the k-th distance: duplicated or lattice data) are redone by the exact kernel.
// A CTA owns 64 queries and streams all candidates in tiles of CT rows, one cp.async.bulk per tile (a tile of the packed
// sample matrix is contiguous) through a 3-stage mbarrier pipeline fed by a producer warp.  The 8 consumer warps are
// AUTONOMOUS: warp w owns queries 8w .. 8w+7 from the DMMA to the selection lists -- its A fragments (the 8 query rows)
// stay in registers for the whole kernel, its 8 x CT distances go through a private slab of shared memory (the C
// fragment layout is transposed into "lane = candidate" for the ballot-based selection), and nothing but the candidate
// stages is shared, so there is no CTA barrier: a warp busy inserting only delays the others once the pipeline is full.
// The candidate tile is row-major with the engine pitch ld == 4 (mod 8): bank-conflict-free DMMA B-fragment loads.
constexpr int kGQ = 64;
constexpr int kGThreads = 9 * 32;
constexpr int kGStages = 3;

template <int CT, int KS>
__global__ void __launch_bounds__(kGThreads, 1)
knn_gemm_kernel(const double* __restrict__ X, const double* __restrict__ nrm2, int64_t n, int d8, int64_t ldx, int kp,
                int* __restrict__ short_idx, double* __restrict__ short_thr, int dbg, unsigned stagger_ns) {
  constexpr int JW = CT / 8;                          // n-blocks of 8 candidates: every warp covers the whole tile
  constexpr int DP = CT + 8;                          // pitch of a distance slab: the 16-byte stores of a quarter warp spread
  extern __shared__ __align__(16) unsigned char smraw[];
  const int ld = (int)ldx;
  double* Cs = reinterpret_cast<double*>(smraw);                       // [stages][CT][ld]
  double* Cn = Cs + (size_t)kGStages * CT * ld;                        // [stages][CT]   squared norms of the tile
  double* Dt = Cn + kGStages * CT;                                     // [8 warps][8][DP]
  double* Ld = Dt + (size_t)kGQ * DP;                                  // [64][kp]
  int* Li = reinterpret_cast<int*>(Ld + (size_t)kGQ * kp);             // [64][kp]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(Li + (size_t)kGQ * kp + ((kGQ * kp) & 1));   // full[3], empty[3]
  const uint32_t sBar = smem_u32(bars);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t q0 = (int64_t)blockIdx.x * kGQ;
  const int64_t tiles = (n + CT - 1) / CT;

  if (tid == 0) {
    for (int s = 0; s < kGStages; ++s) { mbar_init(sBar + 8 * s, 1); mbar_init(sBar + 8 * (kGStages + s), 8); }
    mbar_fence_init();
  }
  for (int i = tid; i < kGQ * kp; i += kGThreads) { Ld[i] = DBL_MAX; Li[i] = -1; }
  __syncthreads();

  if (warp == 8) {
    // ------------------------------ producer: one bulk copy per candidate tile (+ its norms) ------------------------------
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int64_t t = 0; t < tiles; ++t) {
        mbar_wait(sBar + 8 * (kGStages + stage), phase ^ 1u);
        const int64_t c0 = t * CT;
        const int rows = (int)min((int64_t)CT, n - c0);
        const uint32_t bytes = (uint32_t)rows * (uint32_t)ld * 8u;
        const uint32_t nbytes = (uint32_t)((rows + 1) & ~1) * 8u;        // 16-byte granules (the norm array is padded)
        mbar_arrive_expect_tx(sBar + 8 * stage, bytes + nbytes);
        bulk_load(smem_u32(Cs + (size_t)stage * CT * ld), X + c0 * ldx, bytes, sBar + 8 * stage);
        bulk_load(smem_u32(Cn + stage * CT), nrm2 + c0, nbytes, sBar + 8 * stage);
        if (++stage == kGStages) { stage = 0; phase ^= 1u; }
      }
    }
    return;
  }
  // ------------------------------ consumers ------------------------------
  const int g = lane >> 2, t4 = lane & 3;
  const int ksteps = d8 >> 2;
  const int64_t myrow = q0 + 8 * warp + g;                               // the query of this lane's fragment row
  double aq[KS];
#pragma unroll
  for (int s = 0; s < KS; ++s) aq[s] = (s < ksteps && myrow < n) ? X[myrow * ldx + 4 * s + t4] : 0.0;
  const double qn = (myrow < n) ? nrm2[myrow] : 0.0;
  double* Dw = Dt + (size_t)warp * 8 * DP;
  const double* thr_p = Ld + (size_t)(8 * warp + g) * kp + kp - 1;
  uint32_t stage = 0, phase = 0;
  // Warps w and w + 4 share a scheduler and its FP64 pipe.  Started together they stay in lockstep -- both in the MMA
  // loop (each at half rate), then both in the epilogue (pipe idle).  Half a tile of head start for one of them keeps
  // one warp's epilogue under the other's MMAs.
  if (warp >= 4 && stagger_ns > 0) __nanosleep(stagger_ns);
  for (int64_t t = 0; t < tiles; ++t) {
    const int64_t c0 = t * CT;
    mbar_wait(sBar + 8 * stage, phase);
    // B fragments: candidate 8j + g, feature 4s + t4.  Software pipeline one k-step deep: the fragments of step s + 1
    // are requested before the MMAs of step s (kept in program order) issue.
    const double* bp0 = Cs + (size_t)stage * CT * ld + (size_t)g * ld + t4;
    double acc[JW][2];
    double bf[2][JW];
#pragma unroll
    for (int j = 0; j < JW; ++j) { acc[j][0] = 0.0; acc[j][1] = 0.0; bf[0][j] = bp0[8 * j * ld]; }
#pragma unroll
    for (int s = 0; s < KS; ++s) {
      if (s < ksteps) {
        if (s + 1 < KS && s + 1 < ksteps) {
#pragma unroll
          for (int j = 0; j < JW; ++j) bf[(s + 1) & 1][j] = bp0[4 * (s + 1) + 8 * j * ld];
        }
#pragma unroll
        for (int j = 0; j < JW; ++j) dmma_m8n8k4_ordered(acc[j][0], acc[j][1], aq[s], bf[s & 1][j]);
      }
    }
    const double* cn = Cn + stage * CT;
    const double thr = *thr_p;
    bool hit = false;
#pragma unroll
    for (int j = 0; j < JW; ++j) {
      const int cl = 8 * j + 2 * t4;
      const double2 cnv = *reinterpret_cast<const double2*>(cn + cl);
      const double d0 = fma(-2.0, acc[j][0], qn + cnv.x);
      const double d1 = fma(-2.0, acc[j][1], qn + cnv.y);
      *reinterpret_cast<double2*>(Dw + g * DP + cl) = make_double2(d0, d1);
      hit |= (d0 < thr && c0 + cl < n) | (d1 < thr && c0 + cl + 1 < n);
    }
    __syncwarp();                                                       // the slab stores are visible to the whole warp
    const unsigned hb = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) mbar_arrive(sBar + 8 * (kGStages + stage));          // candidate tile and norms are free again
    if (++stage == kGStages) { stage = 0; phase ^= 1u; }
    // selection, only for the queries with a candidate under their threshold (after the first tiles: ~1 in 20 per tile);
    // candidates in increasing index order, ties keep the lower index
    if (hb != 0u && !(dbg == 1 && t >= 64)) {              // dbg 1: timing experiment without the selection (wrong results)
#pragma unroll 1
      for (int a = 0; a < 8; ++a) {
        if (((hb >> (4 * a)) & 0xfu) == 0u) continue;
        const int ql = 8 * warp + a;
        if (q0 + ql >= n) continue;
        double* ldq = Ld + (size_t)ql * kp;
        int* liq = Li + (size_t)ql * kp;
        double th = ldq[kp - 1];
#pragma unroll
        for (int b = 0; b < CT / 32; ++b) {
          const int64_t cg = c0 + lane + 32 * b;
          const double dist = (cg < n) ? Dw[a * DP + lane + 32 * b] : DBL_MAX;
          const unsigned mask = __ballot_sync(0xffffffffu, dist < th);
          if (mask) th = knn_insert(ldq, liq, kp, dist, (int)cg, th, mask, lane);
        }
      }
    }
    __syncwarp();                                                       // the slab is rewritten by the next tile
  }
  for (int a = 0; a < 8; ++a) {
    const int ql = 8 * warp + a;
    const int64_t row = q0 + ql;
    if (row >= n) break;
    for (int p = lane; p < kp; p += 32) short_idx[row * kp + p] = Li[(size_t)ql * kp + p];
    if (lane == 0) short_thr[row] = Ld[(size_t)ql * kp + kp - 1];
  }
}

// squared norms of the rows (the same d-term sums the distance identity needs) and their maximum
__global__ void __launch_bounds__(256)
knn_norms_kernel(const double* __restrict__ X, int64_t n, int d, int64_t ldx, double* __restrict__ nrm2, int64_t n_pad,
                 unsigned long long* __restrict__ max_bits) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double a = 0.0;
  if (i < n) {
    const double* x = X + i * ldx;
    for (int c = 0; c < d; ++c) a = fma(x[c], x[c], a);
  }
  if (i < n_pad) nrm2[i] = a;
  // non-negative doubles order like their bit patterns
  double m = a;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(max_bits, (unsigned long long)__double_as_longlong(m));
}

// One warp per query: exact distances of its shortlist (sum of (a-b)^2 in feature order, like the exact kernel and the
// kd-tree), the k nearest in (distance, index) order, and the certificate described above.
__global__ void __launch_bounds__(256)
knn_rerank_kernel(const double* __restrict__ X, const double* __restrict__ nrm2, int64_t n, int d, int64_t ldx, int k, int kp,
                  const int* __restrict__ short_idx, const double* __restrict__ short_thr,
                  const unsigned long long* __restrict__ max_bits, int* __restrict__ idx_out, int* __restrict__ qlist,
                  int* __restrict__ qcount) {
  __shared__ double sd[8][160];
  __shared__ int si[8][160];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t q = (int64_t)blockIdx.x * 8 + w;
  if (q >= n) return;
  const double* xq = X + q * ldx;
  for (int p = lane; p < kp; p += 32) {
    const int c = short_idx[q * kp + p];
    double a = DBL_MAX;
    if (c >= 0) {
      const double* xc = X + (int64_t)c * ldx;
      a = 0.0;
      for (int f = 0; f < d; ++f) { const double t = xq[f] - xc[f]; a = fma(t, t, a); }
    }
    sd[w][p] = a;
    si[w][p] = c >= 0 ? c : 0x7fffffff;
  }
  __syncwarp();
  double dk = DBL_MAX;
  for (int p = lane; p < kp; p += 32) {
    const double a = sd[w][p];
    const int c = si[w][p];
    int r = 0;
    for (int o = 0; o < kp; ++o) r += (sd[w][o] < a || (sd[w][o] == a && si[w][o] < c)) ? 1 : 0;
    if (r < k) idx_out[q * k + r] = c;
    if (r == k - 1) dk = a;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dk = fmin(dk, __shfl_xor_sync(0xffffffffu, dk, o));
  if (lane == 0) {
    // |approximate - exact| <= (2d + 8) eps (|q|^2 + |c|^2) for every pair; a candidate outside the shortlist has an
    // approximate distance >= thr.  With fewer than kp samples the shortlist is the whole set: nothing was left out.
    const double nmax = __longlong_as_double((long long)*max_bits);
    const double delta = (2.0 * d + 8.0) * DBL_EPSILON * (nrm2[q] + nmax);
    const double thr = short_thr[q];
    const bool certified = (n <= kp) || (dk < thr - 2.0 * delta);
    if (!certified) qlist[atomicAdd(qcount, 1)] = (int)q;
  }
}

// ------------------------------------------------------------------------------------------------
// batched local least squares: one CTA per sample
// ------------------------------------------------------------------------------------------------
constexpr int kGradThreads = 128;

__global__ void __launch_bounds__(kGradThreads)
poly_grad_kernel(const double* __restrict__ X, const double* __restrict__ y, int64_t n, int d, int64_t ldx,
                 const int* __restrict__ idx, int k, int order, int p, double* __restrict__ G, int64_t ldg,
                 int* __restrict__ info, const double* __restrict__ Xq, int64_t ldq) {
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
  // gradient at the evaluation point: the sample itself, or row i of Xq (grad_at away from the samples)
  const double* __restrict__ x0 = (Xq != nullptr) ? Xq + i * ldq : X + i * ldx;
  for (int j = tid; j < d; j += nt) {
    double g = beta[j];
    if (order == 2) {
      int c = d;
      for (int a = 0; a < d; ++a)
        for (int b = a; b < d; ++b, ++c) {
          if (a == j && b == j) g += 2.0 * beta[c] * x0[j];
          else if (a == j) g += beta[c] * x0[b];
          else if (b == j) g += beta[c] * x0[a];
        }
    }
    G[i * ldg + j] = g;
  }
}

}  // namespace

static cudaError_t knn_exact_launch(const double* X, int64_t n, int d, int64_t ldx, int k, int* idx, const int* qlist,
                                    const int* qcount, int64_t n_queries, cudaStream_t s, const double* Qext = nullptr,
                                    int64_t ldq = 0) {
  const size_t smem = ((size_t)kDC * kQPitch + (size_t)kDC * kCPitch + (size_t)kQT * k) * 8 + (size_t)kQT * k * 4;
  cudaError_t e = cudaFuncSetAttribute(knn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) return e;
  const int64_t blocks = (n_queries + kQT - 1) / kQT;
  if (blocks <= 0) return cudaSuccess;
  knn_kernel<<<(unsigned)blocks, kKnnThreads, smem, s>>>(X, n, d, ldx, k, idx, qlist, qcount, Qext, n_queries, ldq);
  return cudaGetLastError();
}

static size_t knn_gemm_smem(int ct, int ld, int kp) {
  size_t doubles = (size_t)kGStages * ct * ld + (size_t)kGStages * ct + (size_t)kGQ * (ct + 8) + (size_t)kGQ * kp;
  size_t bytes = doubles * 8 + ((size_t)kGQ * kp + 1) / 2 * 2 * 4 + 2 * kGStages * 8;
  return bytes;
}

size_t knn_scratch_bytes(int64_t n, int k) {
  const int kp = k + kKnnMargin;
  const size_t n_pad = (size_t)((n + 63) / 64 * 64 + 64);
  // norms | shortlist thresholds | shortlist indices | query list | counters
  return (n_pad + (size_t)n) * 8 + ((size_t)n * kp + (size_t)n) * 4 + 64 + 256;
}

cudaError_t knn_launch(const double* X, int64_t n, int d, int64_t ldx, int k, int* idx, void* scratch, size_t scratch_bytes,
                       int* n_exact_fallback, cudaStream_t s) {
  if (n <= 0 || d <= 0 || k <= 0 || k > kKnnMaxK || k > n) return cudaErrorInvalidValue;
  if (n_exact_fallback) *n_exact_fallback = -1;                           // -1: the exact kernel did everything
  const char* env = getenv("CORRLA_B200_KNN_EXACT");
  const int kp = k + kKnnMargin;
  const int d8 = (d + 7) / 8 * 8;
  int ct = 64;
  if (knn_gemm_smem(ct, (int)ldx, kp) > 225 * 1024) ct = 32;
  const bool gemm_ok = (env == nullptr || env[0] != '1') && scratch != nullptr && scratch_bytes >= knn_scratch_bytes(n, k) &&
                       (ldx % 8) == 4 && ldx >= d8 && d8 <= 128 && n >= 2048 && n < ((int64_t)1 << 31) &&
                       knn_gemm_smem(ct, (int)ldx, kp) <= 225 * 1024;
  if (!gemm_ok) return knn_exact_launch(X, n, d, ldx, k, idx, nullptr, nullptr, n, s);

  const size_t n_pad = (size_t)((n + 63) / 64 * 64 + 64);
  unsigned char* base = static_cast<unsigned char*>(scratch);
  double* nrm2 = reinterpret_cast<double*>(base);
  double* thr = nrm2 + n_pad;
  int* sidx = reinterpret_cast<int*>(thr + n);
  int* qlist = sidx + (size_t)n * kp;
  unsigned long long* counters = reinterpret_cast<unsigned long long*>((reinterpret_cast<uintptr_t>(qlist + n) + 63) & ~(uintptr_t)63);
  cudaError_t e = cudaMemsetAsync(counters, 0, 64, s);
  if (e != cudaSuccess) return e;
  knn_norms_kernel<<<(unsigned)((n_pad + 255) / 256), 256, 0, s>>>(X, n, d, ldx, nrm2, (int64_t)n_pad, counters);
  const size_t smem = knn_gemm_smem(ct, (int)ldx, kp);
  const unsigned blocks = (unsigned)((n + kGQ - 1) / kGQ);
  auto launch = [&](auto kern) -> cudaError_t {
    cudaError_t ee = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
    if (ee != cudaSuccess) return ee;
    const char* dbg_env = getenv("CORRLA_B200_KNN_DEBUG");
    const char* stg_env = getenv("CORRLA_B200_KNN_STAGGER_NS");
    kern<<<blocks, kGThreads, smem, s>>>(X, nrm2, n, d8, ldx, kp, sidx, thr, dbg_env ? atoi(dbg_env) : 0,
                                         stg_env ? (unsigned)atoi(stg_env) : 1200u);
    return cudaGetLastError();
  };
  if (ct == 64) e = (d8 <= 64) ? launch(knn_gemm_kernel<64, 16>) : launch(knn_gemm_kernel<64, 32>);
  else e = (d8 <= 64) ? launch(knn_gemm_kernel<32, 16>) : launch(knn_gemm_kernel<32, 32>);
  if (e != cudaSuccess) return e;
  int* qcount = reinterpret_cast<int*>(counters + 1);
  knn_rerank_kernel<<<(unsigned)((n + 7) / 8), 256, 0, s>>>(X, nrm2, n, d, ldx, k, kp, sidx, thr, counters, idx, qlist, qcount);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  // queries without a certificate go through the exact kernel (normally none)
  int hcount = 0;
  e = cudaMemcpyAsync(&hcount, qcount, sizeof(int), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return e;
  if (n_exact_fallback) *n_exact_fallback = hcount;
  if (hcount > 0) return knn_exact_launch(X, n, d, ldx, k, idx, qlist, qcount, hcount, s);
  return cudaSuccess;
}

cudaError_t knn_query_launch(const double* X, int64_t n, int d, int64_t ldx, const double* Q, int64_t nq, int64_t ldq, int k,
                             int* idx, cudaStream_t s) {
  if (n <= 0 || nq <= 0 || d <= 0 || k <= 0 || k > kKnnMaxK || k > n || Q == nullptr) return cudaErrorInvalidValue;
  return knn_exact_launch(X, n, d, ldx, k, idx, nullptr, nullptr, nq, s, Q, ldq);
}

int poly_grad_num_coef(int d, int order) { return order == 2 ? d + d * (d + 1) / 2 : d; }

size_t poly_grad_smem_bytes(int d, int k, int order) {
  const int p = poly_grad_num_coef(d, order);
  return ((size_t)(p + 1) * (k | 1) + 2 * (size_t)p + 4) * 8;
}

cudaError_t poly_grad_launch(const double* X, const double* y, int64_t n, int d, int64_t ldx, const int* idx, int k,
                             int order, double* G, int64_t ldg, int* info, cudaStream_t s, const double* Xq, int64_t ldq) {
  if (n <= 0 || d <= 0 || k <= 1 || (order != 1 && order != 2)) return cudaErrorInvalidValue;
  const int p = poly_grad_num_coef(d, order);
  const size_t smem = poly_grad_smem_bytes(d, k, order);
  if (p > kGradMaxCoef || smem > 200 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(poly_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) return e;
  poly_grad_kernel<<<(unsigned)n, kGradThreads, smem, s>>>(X, y, n, d, ldx, idx, k, order, p, G, ldg, info, Xq, ldq);
  return cudaGetLastError();
}

}  // namespace corrla
