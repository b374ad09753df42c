#include "gradients.cuh"
#include "ptx.cuh"

#include <algorithm>
#include <cfloat>
#include <cstdio>
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
constexpr int kCPitch = kCT + 1;   // odd pitches (also the query chunk: QT + 1): the transposing stores are conflict free
constexpr int kKnnSlots = 5;      // list entries per lane in the insertion routine: lists of up to 160 entries
constexpr int kKnnMargin = 16;    // extra shortlist entries of the GEMM-form search (see knn_gemm_kernel)

// Warp-collective insertion of the lanes flagged in `mask` (lowest lane first = increasing candidate index) into the
// sorted list (ld, li) of length k; returns the new k-th best distance.  Equal distances keep the lower index first.
template <typename T>
__device__ __noinline__ T knn_insert_t(T* ld, int* li, int k, T dist, int cand, T thr, unsigned mask, int lane) {
  while (mask) {
    const int src = __ffs(mask) - 1;
    const T dv = __shfl_sync(0xffffffffu, dist, src);
    const int iv = __shfl_sync(0xffffffffu, cand, src);
    if (dv < thr) {
      int cnt = 0;
      for (int p = lane; p < k; p += 32) cnt += (ld[p] <= dv) ? 1 : 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      const int pos = cnt;
      T td[kKnnSlots]; int ti[kKnnSlots];
#pragma unroll
      for (int j = 0; j < kKnnSlots; ++j) {
        const int p = lane + 32 * j;
        td[j] = (T)0; ti[j] = 0;
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
__device__ __forceinline__ double knn_insert(double* ld, int* li, int k, double dist, int cand, double thr, unsigned mask, int lane) {
  return knn_insert_t<double>(ld, li, k, dist, cand, thr, mask, lane);
}

// QPW = queries per warp: 8 for the full search; 1 for short query lists (the uncertified queries of the GEMM-form search:
// a CTA then streams the candidates ~8x faster, and eight times as many CTAs share the list).
template <int QPW>
__global__ void __launch_bounds__(kKnnThreads)
knn_kernel(const double* __restrict__ X, int64_t n, int d, int64_t ldx, int k, int* __restrict__ idx_out,
           const int* __restrict__ qlist, const int* __restrict__ qcount, const double* __restrict__ Qext, int64_t nq_ext,
           int64_t ldq) {
  extern __shared__ __align__(16) double smk[];
  constexpr int QT = 8 * QPW, QP = QT + 1;             // queries per CTA, odd pitch of the query chunk
  double* Qs = smk;                                   // [kDC][QP]  query chunk, feature-major
  double* Cs = Qs + kDC * QP;                    // [kDC][kCPitch]  candidate chunk, feature-major
  double* Ld = Cs + kDC * kCPitch;                    // [QT][k]         sorted distances of the current best k
  int* Li = reinterpret_cast<int*>(Ld + (size_t)QT * k);    // [QT][k]   their indices
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t q0 = (int64_t)blockIdx.x * QT;
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
  for (int i = tid; i < QT * k; i += kKnnThreads) { Ld[i] = DBL_MAX; Li[i] = -1; }

  for (int64_t c0 = 0; c0 < n; c0 += kCT) {
    double acc[QPW][4];
#pragma unroll
    for (int a = 0; a < QPW; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
    for (int d0 = 0; d0 < d; d0 += kDC) {
      __syncthreads();
      for (int i = tid; i < kDC * QT; i += kKnnThreads) {
        const int q = i / kDC, dd = i - q * kDC;
        const int64_t row = qrow(q);
        Qs[dd * QP + q] = (row < qrows && d0 + dd < d) ? Qsrc[row * ldqs + d0 + dd] : 0.0;
      }
      for (int i = tid; i < kDC * kCT; i += kKnnThreads) {
        const int cc = i / kDC, dd = i - cc * kDC;
        const int64_t row = c0 + cc;
        Cs[dd * kCPitch + cc] = (row < n && d0 + dd < d) ? X[row * ldx + d0 + dd] : 0.0;
      }
      __syncthreads();
#pragma unroll 4
      for (int dd = 0; dd < kDC; ++dd) {
        double cv[4], qv[QPW];
#pragma unroll
        for (int b = 0; b < 4; ++b) cv[b] = Cs[dd * kCPitch + lane + 32 * b];
#pragma unroll
        for (int a = 0; a < QPW; ++a) qv[a] = Qs[dd * QP + QPW * warp + a];
#pragma unroll
        for (int a = 0; a < QPW; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) { const double t = qv[a] - cv[b]; acc[a][b] = fma(t, t, acc[a][b]); }
      }
    }
    // selection: warp w owns queries QPW w .. QPW w + QPW - 1; candidates are visited in increasing index order.  Only the threshold
    // test is unrolled; the (rare) insertion is one out-of-line routine, which keeps the hot loop in the instruction cache
#pragma unroll
    for (int a = 0; a < QPW; ++a) {
      const int ql = QPW * warp + a;
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
  for (int a = 0; a < QPW; ++a) {
    const int ql = QPW * warp + a;
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

// ------------------------------------------------------------------------------------------------
// The same shortlist search on the TF32 tensor-core path
// ------------------------------------------------------------------------------------------------
// The shortlist only has to CONTAIN the k nearest; the exact re-rank and the certificate decide.  So the N^2 d inner
// products do not need 53 bits: with the samples rounded to TF32 (cvt.rna, relative error 2^-11 per element) and
// mma.sync.m16n8k8 accumulating in FP32, |approximate - exact| <= delta with
//   delta = kTf32DeltaRel(d8) * (|q|^2 + max|c|^2),  kTf32DeltaRel = 1.5 * 2^-10 + (d8 + 24) * 2^-22
// (2 * [2^-10 + 2^-22] |q||c| from the two roundings of every product, |q||c| <= (|q|^2 + |c|^2)/2, the FP32
// accumulation and the FP32 norms in the second term, a factor 1.5 of slack on the first).  The certificate of
// knn_rerank_kernel takes that delta; the shortlist carries kKnnMarginTf32 = 32 extra entries for the candidates the
// coarser distances can misplace (Gaussian samples in 64 dimensions, k = 72 of 2^20: ~12 candidates within 2 delta of the
// k-th distance).  One m16n8k8 does 1024 FMAs where one DMMA.8x8x4 does 256 at a sixteenth of the issue rate.
// Layout: a float copy of the samples with pitch d8 + 4 (== 4 mod 8: conflict-free B-fragment loads), zero padded to
// whole tiles; a CTA owns 128 queries (8 autonomous warps x 16), otherwise as knn_gemm_kernel.
constexpr int kTQ = 128;
constexpr int kKnnMarginTf32 = 32;

__host__ __device__ inline double knn_tf32_delta_rel(int d8) { return 1.5 * 0x1.0p-10 + (double)(d8 + 24) * 0x1.0p-22; }

__device__ __forceinline__ void mma_tf32_m16n8k8(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Distances do not change under a common shift, the error bound above does (it scales with the norms): the float copy
// holds the samples minus their column means.  Column sums first (any summation order will do: whichever shift comes
// out, it is applied consistently) ...
__global__ void __launch_bounds__(256)
knn_colsum_kernel(const double* __restrict__ X, int64_t n, int d, int64_t ldx, double* __restrict__ colsum) {
  __shared__ double part[8][128];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int64_t i = (int64_t)blockIdx.x * 8 + w; i < n; i += (int64_t)gridDim.x * 8)
#pragma unroll
    for (int b = 0; b < 4; ++b) { const int c = lane + 32 * b; if (c < d) acc[b] += X[i * ldx + c]; }
#pragma unroll
  for (int b = 0; b < 4; ++b) part[w][lane + 32 * b] = acc[b];
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += 256) {
    double t = 0.0;
    for (int ww = 0; ww < 8; ++ww) t += part[ww][c];
    atomicAdd(colsum + c, t);
  }
}

// ... then, one warp per row: Xf[i][c] = tf32(x_ic - mean_c) as float bits (rows >= n and columns >= d zero), the FP64
// squared norm of the shifted row, its float copy, and the maximum of those norms (what the certificate needs)
__global__ void __launch_bounds__(256)
knn_center_tf32_kernel(const double* __restrict__ X, const double* __restrict__ colsum, int64_t n, int d, int64_t ldx,
                       float* __restrict__ Xf, float* __restrict__ nf, double* __restrict__ nrm2c, int64_t n_pad, int ldf,
                       unsigned long long* __restrict__ max_bits) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= n_pad) return;
  const double inv_n = 1.0 / (double)n;
  double a = 0.0;
  for (int c = lane; c < ldf; c += 32) {
    float v = 0.0f;
    if (i < n && c < d) {
      const double t = X[i * ldx + c] - colsum[c] * inv_n;
      a = fma(t, t, a);
      uint32_t r;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"((float)t));
      v = __uint_as_float(r);
    }
    Xf[i * ldf + c] = v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane == 0) {
    nrm2c[i] = a;
    nf[i] = (float)a;
    atomicMax(max_bits, (unsigned long long)__double_as_longlong(a));      // non-negative doubles order like their bits
  }
}

template <int CT, int KS>                               // KS k-steps of 8 features (d8 <= 8 KS)
__global__ void __launch_bounds__(kGThreads, 1)
knn_tf32_kernel(const float* __restrict__ Xf, const float* __restrict__ nf, int64_t n, int d8, int ldf, int kp,
                int* __restrict__ short_idx, double* __restrict__ short_thr) {
  constexpr int JW = CT / 8;
  constexpr int DP = CT + 8;                            // == 8 (mod 32): the 8-byte slab stores of a half warp hit 32 banks
  extern __shared__ __align__(16) unsigned char smraw[];
  float* Cs = reinterpret_cast<float*>(smraw);                           // [stages][CT][ldf]
  float* Cn = Cs + (size_t)kGStages * CT * ldf;                          // [stages][CT]
  float* Dt = Cn + kGStages * CT;                                        // [8 warps][16][DP]
  float* Ld = Dt + (size_t)kTQ * DP;                                     // [128][kp]
  int* Li = reinterpret_cast<int*>(Ld + (size_t)kTQ * kp);               // [128][kp]
  float* Thr = reinterpret_cast<float*>(Li + (size_t)kTQ * kp);          // [128] largest distance of each pool = its threshold
  int* Mp = reinterpret_cast<int*>(Thr + kTQ);                           // [128] where that entry sits
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(
      (reinterpret_cast<uintptr_t>(Mp + kTQ) + 7) & ~(uintptr_t)7);                                  // full[3], empty[3]
  const uint32_t sBar = smem_u32(bars);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t q0 = (int64_t)blockIdx.x * kTQ;
  const int64_t tiles = (n + CT - 1) / CT;

  if (tid == 0) {
    for (int s = 0; s < kGStages; ++s) { mbar_init(sBar + 8 * s, 1); mbar_init(sBar + 8 * (kGStages + s), 8); }
    mbar_fence_init();
  }
  for (int i = tid; i < kTQ * kp; i += kGThreads) { Ld[i] = FLT_MAX; Li[i] = -1; }
  for (int i = tid; i < kTQ; i += kGThreads) { Thr[i] = FLT_MAX; Mp[i] = 0; }
  __syncthreads();

  if (warp == 8) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      const uint32_t bytes = (uint32_t)CT * (uint32_t)ldf * 4u, nbytes = (uint32_t)CT * 4u;    // whole tiles: Xf is padded
      for (int64_t t = 0; t < tiles; ++t) {
        mbar_wait(sBar + 8 * (kGStages + stage), phase ^ 1u);
        const int64_t c0 = t * CT;
        mbar_arrive_expect_tx(sBar + 8 * stage, bytes + nbytes);
        bulk_load(smem_u32(Cs + (size_t)stage * CT * ldf), Xf + c0 * ldf, bytes, sBar + 8 * stage);
        bulk_load(smem_u32(Cn + stage * CT), nf + c0, nbytes, sBar + 8 * stage);
        if (++stage == kGStages) { stage = 0; phase ^= 1u; }
      }
    }
    return;
  }
  // ------------------------------ consumers: warp w owns queries 16w .. 16w + 15 ------------------------------
  const int g = lane >> 2, t4 = lane & 3;
  const int ksteps = d8 >> 3;
  const int64_t row_a = q0 + 16 * warp + g, row_b = row_a + 8;           // the two query rows of this lane's fragments
  uint32_t aq[KS][4];
#pragma unroll
  for (int s = 0; s < KS; ++s) {
    const bool on = s < ksteps;
    // rows beyond n are zero rows of the padded copy (n_pad >= tiles * CT + 128 is guaranteed by the launcher)
    aq[s][0] = on ? __float_as_uint(Xf[row_a * ldf + 8 * s + t4]) : 0u;
    aq[s][1] = on ? __float_as_uint(Xf[row_b * ldf + 8 * s + t4]) : 0u;
    aq[s][2] = on ? __float_as_uint(Xf[row_a * ldf + 8 * s + t4 + 4]) : 0u;
    aq[s][3] = on ? __float_as_uint(Xf[row_b * ldf + 8 * s + t4 + 4]) : 0u;
  }
  const float qn_a = nf[row_a], qn_b = nf[row_b];
  float* Dw = Dt + (size_t)warp * 16 * DP;
  const float* thr_pa = Thr + 16 * warp + g;
  const float* thr_pb = thr_pa + 8;
  uint32_t stage = 0, phase = 0;
  for (int64_t t = 0; t < tiles; ++t) {
    const int64_t c0 = t * CT;
    mbar_wait(sBar + 8 * stage, phase);
    const float* bp0 = Cs + (size_t)stage * CT * ldf + (size_t)g * ldf + t4;     // candidate 8j + g, feature 8s + t4 (+4)
    float acc[JW][4];
#pragma unroll
    for (int j = 0; j < JW; ++j) { acc[j][0] = 0.0f; acc[j][1] = 0.0f; acc[j][2] = 0.0f; acc[j][3] = 0.0f; }
#pragma unroll
    for (int s = 0; s < KS; ++s) {
      if (s < ksteps) {
#pragma unroll
        for (int j = 0; j < JW; ++j) {
          const float* bp = bp0 + 8 * j * ldf + 8 * s;
          mma_tf32_m16n8k8(acc[j], aq[s], __float_as_uint(bp[0]), __float_as_uint(bp[4]));
        }
      }
    }
    const float* cn = Cn + stage * CT;
    const float thr_a = *thr_pa, thr_b = *thr_pb;
    bool hit_a = false, hit_b = false;
#pragma unroll
    for (int j = 0; j < JW; ++j) {
      const int cl = 8 * j + 2 * t4;
      const float2 cnv = *reinterpret_cast<const float2*>(cn + cl);
      const float a0 = fmaf(-2.0f, acc[j][0], qn_a + cnv.x), a1 = fmaf(-2.0f, acc[j][1], qn_a + cnv.y);
      const float b0 = fmaf(-2.0f, acc[j][2], qn_b + cnv.x), b1 = fmaf(-2.0f, acc[j][3], qn_b + cnv.y);
      *reinterpret_cast<float2*>(Dw + g * DP + cl) = make_float2(a0, a1);
      *reinterpret_cast<float2*>(Dw + (g + 8) * DP + cl) = make_float2(b0, b1);
      const bool in0 = c0 + cl < n, in1 = c0 + cl + 1 < n;
      hit_a |= (a0 < thr_a && in0) | (a1 < thr_a && in1);
      hit_b |= (b0 < thr_b && in0) | (b1 < thr_b && in1);
    }
    __syncwarp();
    const unsigned hba = __ballot_sync(0xffffffffu, hit_a), hbb = __ballot_sync(0xffffffffu, hit_b);
    if (lane == 0) mbar_arrive(sBar + 8 * (kGStages + stage));
    if (++stage == kGStages) { stage = 0; phase ^= 1u; }
    if ((hba | hbb) != 0u) {
#pragma unroll 1
      for (int a = 0; a < 16; ++a) {
        const unsigned hb = (a < 8) ? hba : hbb;
        if (((hb >> (4 * (a & 7))) & 0xfu) == 0u) continue;
        const int ql = 16 * warp + a;
        if (q0 + ql >= n) continue;
        // The shortlist is an UNSORTED pool (the re-rank sorts by exact distance anyway): a candidate under the threshold
        // replaces the pool's largest entry, then the new largest entry is found with one pass over the pool and one
        // REDUX on order-preserving keys -- about a third of the work of a sorted insertion.
        float* ldq = Ld + (size_t)ql * kp;
        int* liq = Li + (size_t)ql * kp;
        float th = Thr[ql];
        int mp = Mp[ql];
        bool changed = false;
#pragma unroll
        for (int b = 0; b < CT / 32; ++b) {
          const int64_t cg = c0 + lane + 32 * b;
          const float dist = (cg < n) ? Dw[a * DP + lane + 32 * b] : FLT_MAX;
          unsigned mask = __ballot_sync(0xffffffffu, dist < th);
          while (mask) {
            const int src = __ffs(mask) - 1;
            const float dv = __shfl_sync(0xffffffffu, dist, src);
            if (lane == 0) { ldq[mp] = dv; liq[mp] = (int)(c0 + src + 32 * b); }
            __syncwarp();
            unsigned best = 0u;
            int bpos = 0;
            for (int p = lane; p < kp; p += 32) {
              const unsigned bits = __float_as_uint(ldq[p]);
              const unsigned key = (bits & 0x80000000u) ? ~bits : (bits | 0x80000000u);      // unsigned order == float order
              if (key > best) { best = key; bpos = p; }
            }
            const unsigned kmax = __reduce_max_sync(0xffffffffu, best);
            const unsigned who = __ballot_sync(0xffffffffu, best == kmax);
            mp = __shfl_sync(0xffffffffu, bpos, __ffs(who) - 1);
            th = __uint_as_float((kmax & 0x80000000u) ? (kmax & 0x7fffffffu) : ~kmax);
            changed = true;
            mask &= mask - 1;
            mask &= __ballot_sync(0xffffffffu, dist < th);
          }
        }
        if (changed && lane == 0) { Thr[ql] = th; Mp[ql] = mp; }
      }
    }
    __syncwarp();
  }
  for (int a = 0; a < 16; ++a) {
    const int ql = 16 * warp + a;
    const int64_t row = q0 + ql;
    if (row >= n) break;
    for (int p = lane; p < kp; p += 32) short_idx[row * kp + p] = Li[(size_t)ql * kp + p];
    if (lane == 0) short_thr[row] = (double)Thr[ql];
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
                  int* __restrict__ qcount, double delta_rel) {
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
    // |approximate - exact| <= delta_rel (|q|^2 + |c|^2) for every pair ((2d + 8) eps for the DMMA shortlist,
    // knn_tf32_delta_rel for the TF32 one); a candidate outside the shortlist has an approximate distance >= thr.
    // With fewer than kp samples the shortlist is the whole set: nothing was left out.
    const double nmax = __longlong_as_double((long long)*max_bits);
    const double delta = delta_rel * (nrm2[q] + nmax);
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
  // short query lists (fewer than one query per warp of a full grid): one query per warp, eight times the CTAs
  const bool few = n_queries <= 8 * 148;
  const int qt = few ? 8 : kQT;
  const size_t smem = ((size_t)kDC * (qt + 1) + (size_t)kDC * kCPitch + (size_t)qt * k) * 8 + (size_t)qt * k * 4;
  cudaError_t e = cudaFuncSetAttribute(knn_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(knn_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) return e;
  const int64_t blocks = (n_queries + qt - 1) / qt;
  if (blocks <= 0) return cudaSuccess;
  if (few) knn_kernel<1><<<(unsigned)blocks, kKnnThreads, smem, s>>>(X, n, d, ldx, k, idx, qlist, qcount, Qext, n_queries, ldq);
  else knn_kernel<8><<<(unsigned)blocks, kKnnThreads, smem, s>>>(X, n, d, ldx, k, idx, qlist, qcount, Qext, n_queries, ldq);
  return cudaGetLastError();
}

static size_t knn_gemm_smem(int ct, int ld, int kp) {
  size_t doubles = (size_t)kGStages * ct * ld + (size_t)kGStages * ct + (size_t)kGQ * (ct + 8) + (size_t)kGQ * kp;
  size_t bytes = doubles * 8 + ((size_t)kGQ * kp + 1) / 2 * 2 * 4 + 2 * kGStages * 8;
  return bytes;
}

static size_t knn_tf32_smem(int ct, int ldf, int kp) {
  return ((size_t)kGStages * ct * ldf + (size_t)kGStages * ct + (size_t)kTQ * (ct + 8)) * 4 + (size_t)kTQ * kp * 8 +
         (size_t)kTQ * 8 + 8 + 2 * kGStages * 8;
}

// scratch layout shared by knn_scratch_bytes and knn_launch
struct KnnScratch {
  size_t n_pad, off_nrm2, off_thr, off_sidx, off_qlist, off_counters, off_mu, off_nrm2c, off_xf, off_nf, total;
  int ldf;
  KnnScratch(int64_t n, int k, int d) {
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    n_pad = (size_t)((n + 127) / 128 * 128 + 128);
    ldf = (d + 7) / 8 * 8 + 4;
    const size_t kp = (size_t)k + kKnnMarginTf32;                          // the larger of the two margins
    size_t o = 0;
    off_nrm2 = o; o = up(o + n_pad * 8);
    off_thr = o; o = up(o + (size_t)n * 8);
    off_sidx = o; o = up(o + (size_t)n * kp * 4);
    off_qlist = o; o = up(o + (size_t)n * 4);
    off_counters = o; o = up(o + 64);                                      // [0] max |x|^2, [1] query count, [2] max centred |x|^2
    off_mu = o; o = up(o + 128 * 8);
    off_nrm2c = o; o = up(o + n_pad * 8);
    off_xf = o; o = up(o + n_pad * (size_t)ldf * 4);
    off_nf = o; o = up(o + n_pad * 4);
    total = o;
  }
};

size_t knn_scratch_bytes(int64_t n, int k, int d) { return KnnScratch(n, k, d).total + 256; }

cudaError_t knn_launch(const double* X, int64_t n, int d, int64_t ldx, int k, int* idx, void* scratch, size_t scratch_bytes,
                       int* n_exact_fallback, cudaStream_t s) {
  if (n <= 0 || d <= 0 || k <= 0 || k > kKnnMaxK || k > n) return cudaErrorInvalidValue;
  if (n_exact_fallback) *n_exact_fallback = -1;                           // -1: the exact kernel did everything
  const char* env = getenv("CORRLA_B200_KNN_EXACT");
  const char* env32 = getenv("CORRLA_B200_KNN_TF32");                     // "0": shortlist on the FP64 tensor pipe instead
  const int d8 = (d + 7) / 8 * 8;
  const int ldf = d8 + 4;
  // TF32 shortlist (default): tile of 64 or 32 candidates, whichever fits next to the 128 selection lists
  int kp = k + kKnnMarginTf32;
  int ct32 = 64;
  if (knn_tf32_smem(ct32, ldf, kp) > 225 * 1024) ct32 = 32;
  bool tf32_ok = !(env32 != nullptr && env32[0] == '0') && knn_tf32_smem(ct32, ldf, kp) <= 225 * 1024;
  if (!tf32_ok) kp = k + kKnnMargin;
  int ct = 64;
  if (knn_gemm_smem(ct, (int)ldx, kp) > 225 * 1024) ct = 32;
  const bool common_ok = (env == nullptr || env[0] != '1') && scratch != nullptr && scratch_bytes >= knn_scratch_bytes(n, k, d) &&
                         d8 <= 128 && n >= 2048 && n < ((int64_t)1 << 31);
  const bool dmma_ok = (ldx % 8) == 4 && ldx >= d8 && knn_gemm_smem(ct, (int)ldx, kp) <= 225 * 1024;
  if (!common_ok || (!tf32_ok && !dmma_ok)) return knn_exact_launch(X, n, d, ldx, k, idx, nullptr, nullptr, n, s);

  const KnnScratch L(n, k, d);
  const size_t n_pad = L.n_pad;
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(scratch) + 255) & ~(uintptr_t)255);
  double* nrm2 = reinterpret_cast<double*>(base + L.off_nrm2);
  double* thr = reinterpret_cast<double*>(base + L.off_thr);
  int* sidx = reinterpret_cast<int*>(base + L.off_sidx);
  int* qlist = reinterpret_cast<int*>(base + L.off_qlist);
  unsigned long long* counters = reinterpret_cast<unsigned long long*>(base + L.off_counters);
  double* mu = reinterpret_cast<double*>(base + L.off_mu);
  double* nrm2c = reinterpret_cast<double*>(base + L.off_nrm2c);
  float* Xf = reinterpret_cast<float*>(base + L.off_xf);
  float* nf = reinterpret_cast<float*>(base + L.off_nf);
  cudaError_t e = cudaMemsetAsync(counters, 0, 64, s);
  if (e == cudaSuccess) e = cudaMemsetAsync(mu, 0, 128 * 8, s);
  if (e != cudaSuccess) return e;
  knn_norms_kernel<<<(unsigned)((n_pad + 255) / 256), 256, 0, s>>>(X, n, d, ldx, nrm2, (int64_t)n_pad, counters);
  const double* cert_nrm2 = nrm2;                        // norms and their maximum as the certificate's bound sees them
  const unsigned long long* cert_max = counters;
  double delta_rel;
  if (tf32_ok) {
    knn_colsum_kernel<<<148 * 4, 256, 0, s>>>(X, n, d, ldx, mu);
    knn_center_tf32_kernel<<<(unsigned)((n_pad + 7) / 8), 256, 0, s>>>(X, mu, n, d, ldx, Xf, nf, nrm2c, (int64_t)n_pad, ldf, counters + 2);
    cert_nrm2 = nrm2c;
    cert_max = counters + 2;
    const size_t smem = knn_tf32_smem(ct32, ldf, kp);
    const unsigned blocks = (unsigned)((n + kTQ - 1) / kTQ);
    auto launch = [&](auto kern) -> cudaError_t {
      cudaError_t ee = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
      if (ee != cudaSuccess) return ee;
      kern<<<blocks, kGThreads, smem, s>>>(Xf, nf, n, d8, ldf, kp, sidx, thr);
      return cudaGetLastError();
    };
    if (ct32 == 64) e = (d8 <= 64) ? launch(knn_tf32_kernel<64, 8>) : launch(knn_tf32_kernel<64, 16>);
    else e = (d8 <= 64) ? launch(knn_tf32_kernel<32, 8>) : launch(knn_tf32_kernel<32, 16>);
    delta_rel = knn_tf32_delta_rel(d8);
  } else {
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
    delta_rel = (2.0 * d + 8.0) * DBL_EPSILON;
  }
  if (e != cudaSuccess) return e;
  int* qcount = reinterpret_cast<int*>(counters + 1);
  knn_rerank_kernel<<<(unsigned)((n + 7) / 8), 256, 0, s>>>(X, cert_nrm2, n, d, ldx, k, kp, sidx, thr, cert_max, idx, qlist, qcount,
                                                            delta_rel);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  // queries without a certificate go through the exact kernel (normally none)
  int hcount = 0;
  e = cudaMemcpyAsync(&hcount, qcount, sizeof(int), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return e;
  if (n_exact_fallback) *n_exact_fallback = hcount;
  { const char* v = getenv("CORRLA_B200_KNN_VERBOSE");
    if (v != nullptr && v[0] == '1')
      fprintf(stderr, "[knn] n=%lld d=%d k=%d shortlist=%d on %s: %d queries without a certificate go to the exact kernel\n",
              (long long)n, d, k, kp, tf32_ok ? "TF32 mma.sync" : "FP64 DMMA", hcount); }
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
