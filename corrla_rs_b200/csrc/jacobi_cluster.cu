// One-sided Jacobi SVD of the l x l core on a thread-block CLUSTER: the rows of the working matrix X (= U*Sigma) and of
// the accumulated rotations V are split into slabs, one per CTA of the cluster, each held in that CTA's shared memory.
// A rotation of the column pair (p, q) needs the full-length dot product x_p . x_q: every CTA computes the part over
// its rows, pushes it into every peer's shared memory through DSMEM, one cluster barrier later all CTAs hold the same
// partial sums, add them in rank order (bit-identical decisions everywhere, no broadcast) and rotate their own rows.
// Per round: one cluster barrier + one CTA barrier, and 1/C of the shared-memory traffic of the single-CTA kernel
// (small_kernels.cu), which sits at the shared-memory bandwidth floor of one SM.  The slabs of an l = 256 core fit in
// the shared memory of 8 SMs, where the single-CTA kernel has to work out of L2.
#include <cooperative_groups.h>

#include <cfloat>
#include <cstdlib>

#include "small_kernels.cuh"

namespace cg = cooperative_groups;

namespace corrla {

namespace {

constexpr int kCJThreads = 512;          // 64 groups of 8 lanes: every pair of a round (l <= 128) in one step
constexpr int kCJLanes = 8;
constexpr int kCJMaxSweeps = 60;
constexpr int kCJMaxCluster = 8;
constexpr unsigned kCJSpinLimit = 1u << 22;   // bounded waits: a lost transaction becomes an error flag, not a hang

__device__ __forceinline__ uint32_t cj_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cj_mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// 8-byte store into a peer's shared memory that completes 8 bytes of the transaction count of the peer's mbarrier
__device__ __forceinline__ void cj_st_async(uint32_t raddr, double v, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
               ::"r"(raddr), "l"(__double_as_longlong(v)), "r"(rbar) : "memory");
}
__device__ __forceinline__ void cj_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void cj_mbar_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool cj_mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}

// kAsync: the partial sums travel as st.async stores that complete transaction bytes on the RECEIVER's mbarrier, and a
// CTA waits only on its own mbarrier -- no cluster-wide barrier and no release fence per round.  Two buffers / two
// mbarriers alternate; a CTA re-arms the one it has just consumed, which always happens before any peer can send the
// exchange after next (a peer sends exchange e+2 only after it has received this CTA's exchange e+1).
template <bool kAsync>
__global__ void __launch_bounds__(kCJThreads)
jacobi_cluster_kernel(const double* __restrict__ Win, int ldw, int l, double* __restrict__ sigma_out,
                      double* __restrict__ Vr_out, double* __restrict__ Ur_out, int Lrows, int ldo, int transpose,
                      int* info, unsigned spin_limit) {
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ int fail_s;                   // any thread of this CTA gave up waiting for a peer's partial sums
  const int C = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  extern __shared__ __align__(16) double smc[];
  const int h = (l + 1) >> 1;              // pairs per round; 2h players, player index >= l is a bye
  const int N1 = 2 * h - 1;
  int lr = (l + C - 1) / C;
  lr = (lr + 1) & ~1;                      // slab rows, even: columns are 16-byte aligned, pad rows are zero
  const int r0 = min(l, rank * lr), r1 = min(l, r0 + lr);
  double* Xs = smc;                        // [l][lr]  my rows of the working columns
  double* Vs = Xs + (size_t)l * lr;        // [l][lr]  my rows of the accumulated rotations
  double2* rot = reinterpret_cast<double2*>(Vs + (size_t)l * lr);   // [h]  (cos, sin) of the round's rotations
  double* nrm = reinterpret_cast<double*>(rot + h);                 // [l]  squared column norms (replicated, updated identically)
  double* recv = nrm + l;                  // [2][C][l] partial sums received from every rank, double buffered
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(recv + (size_t)2 * C * l);   // [2] mbarriers (kAsync)
  int* rnk = reinterpret_cast<int*>(bars + 2);
  int* anyflag = rnk + l;
  int* bigflag = anyflag + 1;               // some pair of this sweep was further than kJacobiNearCos from orthogonal
  const int tid = threadIdx.x, nt = blockDim.x;
  constexpr int LP = kCJLanes;
  const int grp = tid / LP, sub = tid % LP, ngrp = nt / LP;
  const unsigned gmask = ((1u << LP) - 1u) << ((tid & 31) & ~(LP - 1));
  const int glead = (tid & 31) & ~(LP - 1);
  // lane `sub` of every group pushes to rank `sub` (C <= lanes per group)
  double* const peer_recv = (sub < C) ? cluster.map_shared_rank(recv, sub) : recv;
  const uint32_t peer_recv_a = cj_mapa(cj_smem_u32(recv), (uint32_t)(sub < C ? sub : rank));
  const uint32_t peer_bar_a = cj_mapa(cj_smem_u32(bars), (uint32_t)(sub < C ? sub : rank));
  const uint32_t my_bar_a = cj_smem_u32(bars);
  const uint32_t bytes_cols = (uint32_t)(C * l * 8), bytes_pairs = (uint32_t)(C * h * 8);
  unsigned uses0 = 0, uses1 = 0;              // completed uses of the two mbarriers (phase parity)
  int failed = 0;
  if (tid == 0) fail_s = (spin_limit == 0) ? 1 : 0;      // spin_limit == 0: test hook, the failure path from the start
  if (kAsync && tid == 0) {
    cj_mbar_init(my_bar_a, 1);
    cj_mbar_init(my_bar_a + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    cj_mbar_expect(my_bar_a, bytes_cols);         // exchange 0: the column norms of the first sweep
    cj_mbar_expect(my_bar_a + 8, bytes_pairs);    // exchange 1: the dot products of round 0
  }
  // squared norms of the source columns (all rows; every CTA computes the same numbers), descending rank
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
  if (rank == 0)
    for (int idx = tid; idx < Lrows * ldo; idx += nt) { Vr_out[idx] = 0.0; Ur_out[idx] = 0.0; }
  __syncthreads();
  for (int j = tid; j < l; j += nt) {
    const double sj = nrm[j];
    int r = 0;
    for (int i = 0; i < l; ++i) r += (nrm[i] > sj || (nrm[i] == sj && i < j)) ? 1 : 0;
    rnk[j] = r;
  }
  __syncthreads();
  for (int idx = tid; idx < l * lr; idx += nt) {
    const int j = idx / lr, il = idx - j * lr;           // source column j, local row il
    const int i = r0 + il, c = rnk[j];
    double x = 0.0;
    if (i < r1) x = transpose ? Win[(int64_t)j * ldw + i] : Win[(int64_t)i * ldw + j];
    Xs[(size_t)c * lr + il] = x;
    Vs[(size_t)c * lr + il] = (i < r1 && i == j) ? 1.0 : 0.0;   // V starts as the permutation
  }
  __syncthreads();
  // every CTA of the cluster must have started before anyone writes into a peer's shared memory
  cluster.sync();

  const double tol = sqrt((double)l) * DBL_EPSILON;
  const double tol2 = tol * tol;
  const int lr2 = lr >> 1;
  int par = 0, sweeps = 0, converged = 0;
  // lane sub < C: store one partial sum into slot `slot` of buffer `par` on rank `sub`
  auto push = [&](int slot, double v) {
    const size_t idx = ((size_t)par * C + rank) * l + slot;
    if (kAsync) cj_st_async(peer_recv_a + (uint32_t)(idx * 8), v, peer_bar_a + (uint32_t)(par * 8));
    else peer_recv[idx] = v;
  };
  // all partial sums of the current exchange have arrived in buffer `par`.  next2_cols: the exchange after next (which
  // reuses this buffer and its mbarrier) carries column sums (l values per rank) rather than dot products (h values)
  auto arrived = [&](bool next2_cols) {
    if (kAsync) {
      const uint32_t bar = my_bar_a + (uint32_t)(par * 8);
      const unsigned phase = (par == 0 ? uses0 : uses1) & 1u;
      unsigned spins = 0;
      // after a timeout the CTA leaves the iteration at its next barrier (fail_s is read CTA-uniformly there), stops
      // pushing and re-arming, and meets its peers -- who time out in turn -- at the final cluster barrier: info[1] = -1
      while (!failed && !cj_mbar_try_wait(bar, phase)) { if (++spins > spin_limit) { failed = 1; fail_s = 1; } }
      if (par == 0) ++uses0; else ++uses1;
      if (tid == 0 && !failed) cj_mbar_expect(bar, next2_cols ? bytes_cols : bytes_pairs);
    } else {
      cluster.sync();
    }
  };
  // sum over the ranks of one value per column: partial over my rows -> every peer -> fixed-order sum
  auto column_sums = [&](bool take_sqrt) {
    for (int j = grp; j < l; j += ngrp) {
      const double2* xj = reinterpret_cast<const double2*>(Xs + (size_t)j * lr);
      double a = 0.0;
      for (int i = sub; i < lr2; i += LP) { const double2 x = xj[i]; a += x.x * x.x + x.y * x.y; }
#pragma unroll
      for (int o = LP / 2; o > 0; o >>= 1) a += __shfl_xor_sync(gmask, a, o);
      if (sub < C) push(j, a);                              // the butterfly left the total in every lane
    }
    arrived(false);                                         // two exchanges later: round 1 of this sweep (or nothing)
    for (int j = tid; j < l; j += nt) {
      double a = 0.0;
      for (int src = 0; src < C; ++src) a += recv[((size_t)par * C + src) * l + j];
      nrm[j] = take_sqrt ? sqrt(a) : a;
    }
    if (tid == 0) { *anyflag = 0; *bigflag = 0; }
    __syncthreads();
    par ^= 1;
  };

  for (; sweeps < kCJMaxSweeps; ++sweeps) {
    column_sums(false);                                   // exact norms once per sweep
    if (fail_s) break;                                    // CTA-uniform: read behind the barrier that ends column_sums
    for (int r = 0; r < N1; ++r) {
      for (int pi = grp; pi < h; pi += ngrp) {
        int p, q;
        if (pi == 0) { p = N1; q = r; }
        else { p = r + pi; if (p >= N1) p -= N1; q = r - pi; if (q < 0) q += N1; }
        if (p > q) { const int tmp = p; p = q; q = tmp; }
        double c = 0.0;                                   // a bye sends 0: every exchange has a fixed byte count
        if (q < l) {
          const double2* xp = reinterpret_cast<const double2*>(Xs + (size_t)p * lr);
          const double2* xq = reinterpret_cast<const double2*>(Xs + (size_t)q * lr);
          double c0 = 0.0, c1 = 0.0;
          for (int i = sub; i < lr2; i += LP) { const double2 x = xp[i], y = xq[i]; c0 += x.x * y.x; c1 += x.y * y.y; }
          c = c0 + c1;
#pragma unroll
          for (int o = LP / 2; o > 0; o >>= 1) c += __shfl_xor_sync(gmask, c, o);
          c = __shfl_sync(gmask, c, glead);
        }
        if (sub < C) push(pi, c);
      }
      arrived(r == N1 - 2);                               // exchange after next: column sums iff this is round N1 - 2
      // rotation parameters: ONE thread per pair (the fp64 divide / square roots cost ~150 instructions; done by all
      // lanes of every group they would occupy the FP64 pipe for longer than everything else in the round)
      for (int pi = tid; pi < h; pi += nt) {
        int p, q;
        if (pi == 0) { p = N1; q = r; }
        else { p = r + pi; if (p >= N1) p -= N1; q = r - pi; if (q < 0) q += N1; }
        if (p > q) { const int tmp = p; p = q; q = tmp; }
        double2 cs_sn = make_double2(1.0, 0.0);
        if (q < l) {
          double c = 0.0;
          for (int src = 0; src < C; ++src) c += recv[((size_t)par * C + src) * l + pi];   // same order on every CTA
          const double a = nrm[p], b = nrm[q];
          if (c * c > tol2 * a * b) {
            // division- and sqrt-free form of t = sign(d c) 2|c| / (|d| + sqrt(d^2 + 4c^2)), cs = 1/sqrt(1 + t^2):
            // cos(2 theta) = |d| r with r = 1/sqrt(d^2 + 4c^2); cs = sqrt(u), u = (1 + cos 2theta)/2 in [1/2, 1];
            // sn = sin(2 theta) / (2 cs) = |c| r / cs; t = sn / cs = |c| r / u.  Two rsqrt instead of sqrt + div + rsqrt.
            const double d = b - a;
            const double r = rsqrt(fma(d, d, 4.0 * c * c));
            const double u = fma(0.5 * fabs(d), r, 0.5);
            const double icu = rsqrt(u);
            const double cr = fabs(c) * r;
            const double cs = u * icu;
            const double sn = copysign(cr * icu, d * c);
            const double t = copysign(cr * icu * icu, d * c);
            cs_sn = make_double2(cs, sn);
            nrm[p] = fmax(a - t * c, 0.0); nrm[q] = b + t * c;
            *anyflag = 1;
            if (c * c > kJacobiNearCos2 * a * b) *bigflag = 1;
          }
        }
        rot[pi] = cs_sn;
      }
      __syncthreads();
      if (fail_s) break;                                  // CTA-uniform (set only between the previous barrier and this one)
      for (int pi = grp; pi < h; pi += ngrp) {
        const double2 cs_sn = rot[pi];
        const double cs = cs_sn.x, sn = cs_sn.y;
        if (sn == 0.0) continue;                          // no rotation (or a bye)
        int p, q;
        if (pi == 0) { p = N1; q = r; }
        else { p = r + pi; if (p >= N1) p -= N1; q = r - pi; if (q < 0) q += N1; }
        if (p > q) { const int tmp = p; p = q; q = tmp; }
        double2* xp = reinterpret_cast<double2*>(Xs + (size_t)p * lr);
        double2* xq = reinterpret_cast<double2*>(Xs + (size_t)q * lr);
        double2* vp = reinterpret_cast<double2*>(Vs + (size_t)p * lr);
        double2* vq = reinterpret_cast<double2*>(Vs + (size_t)q * lr);
        for (int i = sub; i < lr2; i += LP) {
          const double2 x = xp[i], y = xq[i];
          xp[i] = make_double2(cs * x.x - sn * y.x, cs * x.y - sn * y.y);
          xq[i] = make_double2(sn * x.x + cs * y.x, sn * x.y + cs * y.y);
          const double2 vx = vp[i], vy = vq[i];
          vp[i] = make_double2(cs * vx.x - sn * vy.x, cs * vx.y - sn * vy.y);
          vq[i] = make_double2(sn * vx.x + cs * vy.x, sn * vx.y + cs * vy.y);
        }
      }
      __syncthreads();
      par ^= 1;
    }
    if (fail_s) break;
    // Quadratic convergence: when every rotation of a sweep was by less than kJacobiNearCos, what is left afterwards is
    // of the order of its square, below the tolerance -- no further sweep is needed just to confirm it.
    const int any = *anyflag, big = *bigflag;             // identical on every CTA: same sums, same decisions
    if (!any || !big) { converged = 1; ++sweeps; break; }
  }

  // singular values (column norms over all rows), descending ranks, outputs for my rows
  if (!fail_s) column_sums(true);
  // Agree on the outcome across the cluster: a CTA that gave up stops pushing, so its peers time out in turn, but the
  // barriers below must be taken by everybody or by nobody.  After this barrier every CTA reads every fail flag.
  __syncthreads();
  cluster.sync();
  int anyfail = 0;
  for (int r = 0; r < C; ++r) anyfail |= *cluster.map_shared_rank(&fail_s, r);
  const bool dead = anyfail != 0;                         // cluster-uniform: nothing is written, info[1] = -1
  for (int j = tid; j < l && !dead; j += nt) {
    const double sj = nrm[j];
    int r = 0;
    for (int i = 0; i < l; ++i) r += (nrm[i] > sj || (nrm[i] == sj && i < j)) ? 1 : 0;
    rnk[j] = r;
    if (rank == 0) sigma_out[r] = sj;
  }
  __syncthreads();
  double* out_ux = transpose ? Vr_out : Ur_out;
  double* out_va = transpose ? Ur_out : Vr_out;
  for (int idx = tid; idx < l * lr && !dead; idx += nt) {
    const int j = idx / lr, il = idx - j * lr;
    const int i = r0 + il;
    if (i >= r1) continue;
    const int r = rnk[j];
    const double sj = nrm[j];
    out_ux[(int64_t)i * ldo + r] = sj > 0.0 ? Xs[(size_t)j * lr + il] / sj : 0.0;
    out_va[(int64_t)i * ldo + r] = Vs[(size_t)j * lr + il];
  }
  if (rank == 0 && tid == 0 && info != nullptr) { info[0] = sweeps; info[1] = dead ? -1 : converged; }
  // exactly zero singular values leave zero columns in Ux: rank 0 completes them to an orthonormal basis (unit
  // vectors, two Gram-Schmidt passes), like the single-CTA kernel
  __syncthreads();
  int nz = 0;
  for (int j = 0; j < l; ++j) nz += (nrm[j] > 0.0) ? 1 : 0;          // replicated data: uniform everywhere
  // (after a failed exchange the replicated data may differ between CTAs: `dead` is what keeps this branch uniform)
  if (nz < l && !dead) {
    __threadfence();
    cluster.sync();                                                   // all slabs of Ux are in global memory
    if (rank == 0) {
      double* coef = nrm;
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
  }
  // nobody may exit while a peer can still write into its shared memory (the last pushes precede the last barrier)
  __threadfence();
  cluster.sync();
}

}  // namespace

// Returns cudaErrorNotSupported when the cluster variant does not apply (size, environment switch); any other error is a
// launch failure the caller may answer with the single-CTA kernel.
cudaError_t jacobi_svd_cluster_launch(const double* W, int ldw, int l, double* sigma, double* Vr, double* Ur, int Lrows,
                                      int ldo, int* info, cudaStream_t s, int transpose) {
  // CORRLA_B200_JACOBI_CLUSTER = 0 disables the cluster variant, 2 / 4 / 8 pick the cluster size (default 8)
  static const int req = [] { const char* e = getenv("CORRLA_B200_JACOBI_CLUSTER"); return e == nullptr ? kCJMaxCluster : atoi(e); }();
  if (req <= 0 || l < 32) return cudaErrorNotSupported;
  const int C = (req == 2 || req == 4) ? req : kCJMaxCluster;
  int lr = (l + C - 1) / C;
  lr = (lr + 1) & ~1;
  const size_t smem = ((size_t)2 * l * lr + (size_t)l + (size_t)2 * C * l + (size_t)(l + 1) + 2) * 8 + ((size_t)l + 4) * 4;
  if (smem > (size_t)220 * 1024) return cudaErrorNotSupported;
  // CORRLA_B200_JACOBI_EXCHANGE=barrier selects the cluster-barrier exchange instead of st.async + mbarrier
  static const int use_async = [] { const char* e = getenv("CORRLA_B200_JACOBI_EXCHANGE"); return (e != nullptr && e[0] == 'b') ? 0 : 1; }();
  auto kern = use_async ? jacobi_cluster_kernel<true> : jacobi_cluster_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)C, 1, 1);
  cfg.blockDim = dim3(kCJThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // CORRLA_B200_TEST_JACOBI_SPIN_LIMIT: test hook -- a tiny limit forces the bounded-wait timeout path
  static const unsigned spin_limit = [] { const char* e = getenv("CORRLA_B200_TEST_JACOBI_SPIN_LIMIT"); return e == nullptr ? kCJSpinLimit : (unsigned)strtoul(e, nullptr, 10); }();
  return cudaLaunchKernelEx(&cfg, kern, W, ldw, l, sigma, Vr, Ur, Lrows, ldo, transpose, info, spin_limit);
}

}  // namespace corrla
