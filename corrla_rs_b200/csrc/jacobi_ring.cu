// One-sided Jacobi SVD of the l x l core as a SYSTOLIC RING of warps (odd-even transposition ordering).
//
// jacobi_cluster.cu splits the ROWS over the CTAs of a cluster, so every round needs a cluster-wide sum of the 55 partial
// dot products before anybody can rotate: 3000 cycles per round, almost all of it exchange latency, CTA barriers and the
// one-thread-per-pair rotation parameters.  Here the COLUMN PAIRS are split instead: warp g of the cluster holds the
// two columns of positions (2g, 2g+1) -- X part and accumulated-rotation part, whole columns -- in REGISTERS (lane = row
// mod 32).  A step is then warp-local: three dot products by butterfly, the rotation parameters computed redundantly by
// all lanes (no shared-memory round trip, no barrier), the rotation in registers.  Between steps ONE column per warp
// moves to a neighbouring warp:
//
//   even step : pairs (2g, 2g+1)    rotate, then the column left in position 2g   goes to warp g-1
//   odd  step : pairs (2g+1, 2g+2)  rotate, then the column left in position 2g+2 goes to warp g+1
//
// with the exchange of the two columns of a pair after every rotation folded into the naming of the registers (the
// rotation is done in place, the "left" register set is sent after odd steps and the "right" one after even steps).
// This is the odd-even transposition sort with unconditional exchanges: in n steps the n columns reverse their order and
// every pair of columns meets exactly once -- one sweep.  The first warp parks the idle column of position 0 during the
// odd steps, the last warp idles with position n-1.
// A column travels through a mailbox in the receiver's shared memory: plain stores + an mbarrier arrive inside a CTA,
// st.async stores completing transaction bytes on the receiver's mbarrier across CTAs (one warp in wpc has a remote
// neighbour on each side, and since a dependency chain crosses a CTA boundary only every wpc-th step, the DSMEM latency is
// paid once per wpc steps).  No mailbox needs an "empty" signal: between two columns from the same sender the
// receiver sends one back that depends on the first (see the comments at send / receive).  Per sweep one cluster barrier
// collects the convergence flags.  Waits are bounded; a lost column becomes info[1] = -1, never a hang.
#include <cooperative_groups.h>

#include <cfloat>
#include <cstdio>
#include <cstdlib>

#include "small_kernels.cuh"

namespace cg = cooperative_groups;

namespace corrla {

namespace {

constexpr int kRJMaxSweeps = 60;
constexpr int kRJMaxWarps = 16;               // warps per CTA
constexpr unsigned kRJSpinLimit = 1u << 22;

__device__ __forceinline__ uint32_t rj_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t rj_mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void rj_st_async2(uint32_t raddr, double a, double b, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];"
               ::"r"(raddr), "l"(__double_as_longlong(a)), "l"(__double_as_longlong(b)), "r"(rbar) : "memory");
}
__device__ __forceinline__ void rj_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void rj_mbar_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rj_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool rj_mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}

__device__ __forceinline__ double rj_warp_sum(double a) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  return a;
}

// NR = rows per lane (rows 32k + lane, k < NR): l <= 32 NR.  A mailbox holds one column: [2 NR][32] doubles stored as
// 16-byte pairs (k2, lane) -> X rows (2 k2, 2 k2 + 1) of that lane, then the same for the rotation part.
template <int NR, bool kDbg>
__global__ void __launch_bounds__(32 * kRJMaxWarps, 1)
jacobi_ring_kernel(const double* __restrict__ Win, int ldw, int l, double* __restrict__ sigma_out,
                   double* __restrict__ Vr_out, double* __restrict__ Ur_out, int Lrows, int ldo, int transpose, int* info,
                   unsigned spin_limit, long long* dbg, int fail_at_sweep) {
  long long t_rot = 0, t_send = 0, t_wait = 0, t_loop = 0, t_mark = 0;      // kDbg: cycles per phase of this warp
  auto tick = [&](long long& acc) { if (kDbg) { const long long now = clock64(); acc += now - t_mark; t_mark = now; } };
  cg::cluster_group cluster = cg::this_cluster();
  constexpr int COLW = 2 * NR * 32;                 // doubles per mailbox
  constexpr uint32_t kColBytes = COLW * 8;
  __shared__ int fail_s;
  __shared__ int flags_s[3][2];                     // [sweep mod 3][any, big]
  const int C = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, wpc = nt >> 5;
  const int h = (l + 1) >> 1, n = 2 * h;            // warps in the ring, positions (position l is a zero column if l is odd)
  const int g = rank * wpc + warp;                  // my place in the ring
  const bool active = g < h;
  const bool first = (g == 0), last = (g == h - 1);

  extern __shared__ __align__(16) double smr[];
  double* box_l = smr;                              // [wpc][COLW] mailbox filled by my left neighbour
  double* box_r = box_l + (size_t)wpc * COLW;       // [wpc][COLW] mailbox filled by my right neighbour
  double* park = box_r + (size_t)wpc * COLW;        // [COLW]      position 0 during the odd steps (first warp only)
  double* nrm = park + COLW;                        // [n] squared norms of the source columns, later final singular values
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(nrm + n);    // [wpc][2] mailbox-full barriers
  int* inv = reinterpret_cast<int*>(bars + 2 * wpc);                            // [n] position -> source column (-1: zero column)

  const uint32_t my_box_l = rj_smem_u32(box_l + (size_t)warp * COLW), my_box_r = rj_smem_u32(box_r + (size_t)warp * COLW);
  const uint32_t my_bar_l = rj_smem_u32(bars + 2 * warp), my_bar_r = my_bar_l + 8;
  // my neighbours' mailboxes: the left neighbour receives "from the right" and vice versa
  const bool left_remote = active && !first && warp == 0;              // warp g-1 lives in the previous CTA
  const bool right_remote = active && !last && warp == wpc - 1;        // warp g+1 lives in the next CTA
  uint32_t left_box, left_bar, right_box, right_bar;
  {
    const int lw = (warp == 0) ? wpc - 1 : warp - 1, rw = (warp == wpc - 1) ? 0 : warp + 1;
    const uint32_t lb = rj_smem_u32(box_r + (size_t)lw * COLW), lbar = rj_smem_u32(bars + 2 * lw) + 8;
    const uint32_t rb = rj_smem_u32(box_l + (size_t)rw * COLW), rbar = rj_smem_u32(bars + 2 * rw);
    left_box = left_remote ? rj_mapa(lb, (uint32_t)(rank - 1)) : lb;
    left_bar = left_remote ? rj_mapa(lbar, (uint32_t)(rank - 1)) : lbar;
    right_box = right_remote ? rj_mapa(rb, (uint32_t)(rank + 1)) : rb;
    right_bar = right_remote ? rj_mapa(rbar, (uint32_t)(rank + 1)) : rbar;
  }

  if (tid == 0) { fail_s = (spin_limit == 0) ? 1 : 0; for (int i = 0; i < 6; ++i) (&flags_s[0][0])[i] = 0; }
  if (lane == 0) {
    rj_mbar_init(my_bar_l, 1);
    rj_mbar_init(my_bar_r, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // a mailbox fed from another CTA is armed by its owner: the st.async stores complete its transaction bytes
    if (left_remote) rj_mbar_expect(my_bar_l, kColBytes);
    if (right_remote) rj_mbar_expect(my_bar_r, kColBytes);
  }
  // squared norms of the source columns (every CTA computes the same numbers) and their descending rank = start position
  for (int j = warp; j < n; j += wpc) {
    double a = 0.0;
    if (j < l)
      for (int i = lane; i < l; i += 32) {
        const double x = transpose ? Win[(int64_t)j * ldw + i] : Win[(int64_t)i * ldw + j];
        a += x * x;
      }
    a = rj_warp_sum(a);
    if (lane == 0) nrm[j] = a;
  }
  if (rank == 0)
    for (int idx = tid; idx < Lrows * ldo; idx += nt) { Vr_out[idx] = 0.0; Ur_out[idx] = 0.0; }
  __syncthreads();
  for (int j = tid; j < n; j += nt) {
    if (j >= l) { inv[j] = -1; continue; }          // the zero column of an odd l starts (and stays irrelevant) at the end
    const double sj = nrm[j];
    int r = 0;
    for (int i = 0; i < l; ++i) r += (nrm[i] > sj || (nrm[i] == sj && i < j)) ? 1 : 0;
    inv[r] = j;
  }
  __syncthreads();

  double xl[NR], xr[NR], vl[NR], vr[NR];            // the "left" and "right" register sets (see the header comment)
#pragma unroll
  for (int k = 0; k < NR; ++k) { xl[k] = xr[k] = vl[k] = vr[k] = 0.0; }
  if (active) {
    const int jl = inv[2 * g], jr = inv[2 * g + 1];
#pragma unroll
    for (int k = 0; k < NR; ++k) {
      const int i = 32 * k + lane;
      if (i < l) {
        if (jl >= 0) { xl[k] = transpose ? Win[(int64_t)jl * ldw + i] : Win[(int64_t)i * ldw + jl]; vl[k] = (i == jl) ? 1.0 : 0.0; }
        if (jr >= 0) { xr[k] = transpose ? Win[(int64_t)jr * ldw + i] : Win[(int64_t)i * ldw + jr]; vr[k] = (i == jr) ? 1.0 : 0.0; }
      }
    }
  }
  // every CTA of the cluster has initialised its barriers before anyone writes into a peer's shared memory
  cluster.sync();

  const double tol = sqrt((double)l) * DBL_EPSILON;
  const double tol2 = tol * tol;
  unsigned phase_l = 0, phase_r = 0;               // columns received so far from the left / right (mbarrier phase parity)
  int dead = 0;                                     // warp-uniform: this warp gave up (or saw the CTA's fail flag)
  int any = 0, big = 0;

  auto fail_everywhere = [&]() {
    if (lane < C) *cluster.map_shared_rank(&fail_s, lane) = 1;
  };
  // Column into a neighbour's mailbox.  The mailbox is free: the previous column this warp sent there was consumed before
  // the neighbour sent back the column this warp has just rotated (that column depends on it).
  auto send = [&](const double (&x)[NR], const double (&v)[NR], uint32_t box, uint32_t bar, bool remote) {
    if (remote) {
#pragma unroll
      for (int k2 = 0; k2 < NR / 2; ++k2) {
        rj_st_async2(box + (uint32_t)((k2 * 32 + lane) * 16), x[2 * k2], x[2 * k2 + 1], bar);
        rj_st_async2(box + (uint32_t)(((NR / 2 + k2) * 32 + lane) * 16), v[2 * k2], v[2 * k2 + 1], bar);
      }
    } else {
#pragma unroll
      for (int k2 = 0; k2 < NR / 2; ++k2) {
        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(box + (uint32_t)((k2 * 32 + lane) * 16)), "d"(x[2 * k2]), "d"(x[2 * k2 + 1]) : "memory");
        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(box + (uint32_t)(((NR / 2 + k2) * 32 + lane) * 16)), "d"(v[2 * k2]), "d"(v[2 * k2 + 1]) : "memory");
      }
      __syncwarp();
      if (lane == 0) rj_mbar_arrive(bar);
    }
  };
  // Column out of my mailbox; false after a timeout (or when some warp of the cluster has failed).
  auto receive = [&](double (&x)[NR], double (&v)[NR], uint32_t box, uint32_t bar, unsigned& phase, bool remote) -> bool {
    unsigned spins = 0;
    while (!rj_mbar_try_wait(bar, phase & 1u)) {
      ++spins;
      if ((spins & 63u) == 0 && *reinterpret_cast<volatile int*>(&fail_s)) return false;
      if (spins > spin_limit) { fail_everywhere(); return false; }
    }
    ++phase;
#pragma unroll
    for (int k2 = 0; k2 < NR / 2; ++k2) {
      asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x[2 * k2]), "=d"(x[2 * k2 + 1]) : "r"(box + (uint32_t)((k2 * 32 + lane) * 16)) : "memory");
      asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v[2 * k2]), "=d"(v[2 * k2 + 1]) : "r"(box + (uint32_t)(((NR / 2 + k2) * 32 + lane) * 16)) : "memory");
    }
    // re-arm for the next column from the other CTA (it cannot be on its way before this warp has sent one back)
    if (remote && lane == 0) rj_mbar_expect(bar, kColBytes);
    return true;
  };
  // rotation of the pair held in (xl, vl), (xr, vr), in place
  auto rotate = [&]() {
    double a = 0.0, b = 0.0, c = 0.0;
#pragma unroll
    for (int k = 0; k < NR; ++k) { a = fma(xl[k], xl[k], a); b = fma(xr[k], xr[k], b); c = fma(xl[k], xr[k], c); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
      c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if (c * c > tol2 * a * b) {
      // same division-free rotation parameters as the other two kernels (see jacobi_cluster.cu)
      const double d = b - a;
      const double r = rsqrt(fma(d, d, 4.0 * c * c));
      const double u = fma(0.5 * fabs(d), r, 0.5);
      const double icu = rsqrt(u);
      const double cr = fabs(c) * r;
      const double cs = u * icu;
      const double sn = copysign(cr * icu, d * c);
      any = 1;
      if (c * c > kJacobiNearCos2 * a * b) big = 1;
#pragma unroll
      for (int k = 0; k < NR; ++k) {
        const double x = xl[k], y = xr[k];
        xl[k] = cs * x - sn * y; xr[k] = sn * x + cs * y;
        const double p = vl[k], q = vr[k];
        vl[k] = cs * p - sn * q; vr[k] = sn * p + cs * q;
      }
    }
  };

  int sweeps = 0, converged = 0;
  if (*reinterpret_cast<volatile int*>(&fail_s)) dead = 1;
  for (; sweeps < kRJMaxSweeps; ++sweeps) {
    // test hook: warp 1 behaves as if a wait had timed out at the start of this sweep -- its neighbours are left waiting
    if (sweeps == fail_at_sweep && g == 1 && !dead) { fail_everywhere(); dead = 1; }
    if (active && !dead) {
      if (kDbg) { t_mark = clock64(); t_loop -= t_mark; }
      for (int t = 0; t < n && !dead; t += 2) {
        // even step: pairs (2g, 2g+1).  Afterwards the column now in position 2g is the RIGHT register set.
        rotate();
        tick(t_rot);
        if (first) {
#pragma unroll
          for (int k = 0; k < NR; ++k) { park[(2 * k) * 32 + lane] = xr[k]; park[(2 * k + 1) * 32 + lane] = vr[k]; }
        } else {
          send(xr, vr, left_box, left_bar, left_remote);
        }
        tick(t_send);
        bool have_right = false;
        if (!last) { have_right = receive(xr, vr, my_box_r, my_bar_r, phase_r, right_remote); if (!have_right) { dead = 1; break; } }
        tick(t_wait);
        // odd step: pairs (2g+1, 2g+2); the last warp idles with position n-1 in its LEFT set
        if (!last) rotate();
        tick(t_rot);
        // afterwards the column in position 2g+2 is the LEFT register set: it goes right, position 2g comes from the left
        if (!last) send(xl, vl, right_box, right_bar, right_remote);
        else {
#pragma unroll
          for (int k = 0; k < NR; ++k) { xr[k] = xl[k]; vr[k] = vl[k]; }
        }
        tick(t_send);
        if (first) {
#pragma unroll
          for (int k = 0; k < NR; ++k) { xl[k] = park[(2 * k) * 32 + lane]; vl[k] = park[(2 * k + 1) * 32 + lane]; }
        } else {
          if (!receive(xl, vl, my_box_l, my_bar_l, phase_l, left_remote)) { dead = 1; break; }
        }
        tick(t_wait);
      }
      if (kDbg) t_loop += clock64();
      if (lane < C && (any | big)) atomicOr(cluster.map_shared_rank(&flags_s[sweeps % 3][0], lane), any);
      if (lane < C && big) atomicOr(cluster.map_shared_rank(&flags_s[sweeps % 3][1], lane), big);
    }
    cluster.sync();                                  // flags (and a failure, if any) are visible everywhere
    if (*reinterpret_cast<volatile int*>(&fail_s)) { dead = 1; break; }   // cluster-uniform: set before the barrier, everywhere
    const int sany = *reinterpret_cast<volatile int*>(&flags_s[sweeps % 3][0]);
    const int sbig = *reinterpret_cast<volatile int*>(&flags_s[sweeps % 3][1]);
    if (tid == 0) { flags_s[(sweeps + 2) % 3][0] = 0; flags_s[(sweeps + 2) % 3][1] = 0; }
    any = 0; big = 0;
    // quadratic convergence: a sweep of small rotations only leaves residuals below the tolerance (small_kernels.cuh)
    if (!sany || !sbig) { converged = 1; ++sweeps; break; }
  }

  // singular values: every warp publishes the norms of its two columns in every CTA (-1 marks the zero column of an odd l)
  if (active && !dead) {
    double a = 0.0, b = 0.0, va = 0.0, vb = 0.0;
#pragma unroll
    for (int k = 0; k < NR; ++k) { a = fma(xl[k], xl[k], a); b = fma(xr[k], xr[k], b); va = fma(vl[k], vl[k], va); vb = fma(vr[k], vr[k], vb); }
    a = rj_warp_sum(a); b = rj_warp_sum(b); va = rj_warp_sum(va); vb = rj_warp_sum(vb);
    if (lane < C) {
      double* peer = cluster.map_shared_rank(nrm, lane);
      peer[2 * g] = (va > 0.0) ? sqrt(a) : -1.0;
      peer[2 * g + 1] = (vb > 0.0) ? sqrt(b) : -1.0;
    }
  }
  __syncthreads();
  cluster.sync();
  int anyfail = 0;
  for (int r = 0; r < C; ++r) anyfail |= *cluster.map_shared_rank(&fail_s, r);
  const bool failed = anyfail != 0;                 // cluster-uniform
  double* out_ux = transpose ? Vr_out : Ur_out;
  double* out_va = transpose ? Ur_out : Vr_out;
  if (active && !failed) {
#pragma unroll
    for (int side = 0; side < 2; ++side) {
      const int pos = 2 * g + side;
      const double sj = nrm[pos];
      if (sj < 0.0) continue;                       // the zero column
      int r = 0;
      for (int j = lane; j < n; j += 32) { const double o = nrm[j]; r += (o > sj || (o == sj && j < pos)) ? 1 : 0; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
      if (lane == 0) sigma_out[r] = sj;
#pragma unroll
      for (int k = 0; k < NR; ++k) {
        const int i = 32 * k + lane;
        if (i < l) {
          out_ux[(int64_t)i * ldo + r] = sj > 0.0 ? (side == 0 ? xl[k] : xr[k]) / sj : 0.0;
          out_va[(int64_t)i * ldo + r] = (side == 0 ? vl[k] : vr[k]);
        }
      }
    }
  }
  if (rank == 0 && tid == 0 && info != nullptr) { info[0] = sweeps; info[1] = failed ? -1 : converged; }
  if (kDbg && dbg != nullptr && active && lane == 0) {
    dbg[4 * g + 0] = t_rot; dbg[4 * g + 1] = t_send; dbg[4 * g + 2] = t_wait; dbg[4 * g + 3] = t_loop;
  }
  // exactly zero singular values leave zero columns in Ux: rank 0 completes them to an orthonormal basis (unit vectors,
  // two Gram-Schmidt passes), like the other kernels
  int nz = 0;
  for (int j = 0; j < n; ++j) nz += (nrm[j] > 0.0) ? 1 : 0;            // replicated data: uniform everywhere
  if (nz < l && !failed) {
    __threadfence();
    cluster.sync();                                                   // all columns of Ux are in global memory
    if (rank == 0) {
      double* coef = nrm;
      int cand = 0;
      for (int r = nz; r < l; ++r) {
        for (; cand < l; ++cand) {
          __syncthreads();
          for (int i = tid; i < l; i += nt) out_ux[(int64_t)i * ldo + r] = (i == cand) ? 1.0 : 0.0;
          __syncthreads();
          for (int pass = 0; pass < 2; ++pass) {
            for (int q = warp; q < r; q += wpc) {
              double a = 0.0;
              for (int i = lane; i < l; i += 32) a += out_ux[(int64_t)i * ldo + q] * out_ux[(int64_t)i * ldo + r];
              a = rj_warp_sum(a);
              if (lane == 0) coef[q] = a;
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
            a = rj_warp_sum(a);
            if (tid == 0) coef[r] = a;
          }
          __syncthreads();
          const double n2 = coef[r];
          if (n2 > 0.25) {
            const double inv2 = rsqrt(n2);
            for (int i = tid; i < l; i += nt) out_ux[(int64_t)i * ldo + r] *= inv2;
            ++cand;
            break;
          }
        }
      }
    }
  }
  // nobody may exit while a peer can still write into its shared memory
  __threadfence();
  cluster.sync();
}

}  // namespace

// Returns cudaErrorNotSupported when the ring variant does not apply (size, environment switch); any other error is a
// launch failure the caller may answer with the next kernel in line.
cudaError_t jacobi_svd_ring_launch(const double* W, int ldw, int l, double* sigma, double* Vr, double* Ur, int Lrows,
                                   int ldo, int* info, cudaStream_t s, int transpose) {
  // CORRLA_B200_JACOBI_RING=0 disables the ring kernel (the row-slab cluster kernel / single-CTA kernel take over)
  static const int enabled = [] { const char* e = getenv("CORRLA_B200_JACOBI_RING"); return e == nullptr ? 1 : atoi(e); }();
  if (!enabled || l < 4 || l > 256) return cudaErrorNotSupported;
  const int h = (l + 1) / 2, n = 2 * h;
  // One warp per SM sub-partition when the cluster can grow that far: the step is a chain of dependent FP64 instructions
  // and two warps on one sub-partition get in each other's way (l = 110: 7 warps per CTA on 8 CTAs 0.97 M cycles, 4 warps
  // per CTA on 16 CTAs 0.89 M).  A cluster of 16 is a non-portable size: it is used only when the occupancy query says
  // the device can place it.  CORRLA_B200_JACOBI_RING_WPC=n / CORRLA_B200_JACOBI_RING_MAXC=8 override the plan.
  static const int wpc_target = [] { const char* e = getenv("CORRLA_B200_JACOBI_RING_WPC"); const int v = e == nullptr ? 4 : atoi(e); return (v >= 1 && v <= kRJMaxWarps) ? v : 4; }();
  static const int c_limit = [] { const char* e = getenv("CORRLA_B200_JACOBI_RING_MAXC"); const int v = e == nullptr ? 16 : atoi(e); return (v == 1 || v == 2 || v == 4 || v == 8) ? v : 16; }();
  const int nr = (l <= 64) ? 2 : (l <= 128 ? 4 : 8);
  const size_t colw = (size_t)2 * nr * 32;
  // CORRLA_B200_JACOBI_DEBUG=1: per-warp phase clocks of the step loop, printed to stderr (synchronises the stream)
  static const int debug = [] { const char* e = getenv("CORRLA_B200_JACOBI_DEBUG"); return e == nullptr ? 0 : atoi(e); }();
  static long long* dbg_dev = nullptr;
  if (debug && dbg_dev == nullptr && cudaMalloc(&dbg_dev, 4 * 128 * sizeof(long long)) != cudaSuccess) { cudaGetLastError(); dbg_dev = nullptr; }
  const bool dbg_on = debug && dbg_dev != nullptr;
  auto kern = dbg_on ? ((nr == 2) ? jacobi_ring_kernel<2, true> : (nr == 4 ? jacobi_ring_kernel<4, true> : jacobi_ring_kernel<8, true>))
                     : ((nr == 2) ? jacobi_ring_kernel<2, false> : (nr == 4 ? jacobi_ring_kernel<4, false> : jacobi_ring_kernel<8, false>));
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  int C = 0, wpc = 0;
  for (int cmax = c_limit; cmax >= 1; cmax = (cmax > 8 ? 8 : 0)) {
    C = 1;
    while (C < cmax && (h + C - 1) / C > wpc_target) C *= 2;
    wpc = (h + C - 1) / C;
    if (wpc > kRJMaxWarps) { if (cmax > 8) continue; return cudaErrorNotSupported; }
    cfg = cudaLaunchConfig_t{};
    cfg.gridDim = dim3((unsigned)C, 1, 1);
    cfg.blockDim = dim3((unsigned)(32 * wpc), 1, 1);
    cfg.dynamicSmemBytes = ((size_t)2 * wpc * colw + colw + (size_t)n) * 8 + (size_t)2 * wpc * 8 + (size_t)n * 4 + 16;
    cfg.stream = s;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (C <= 8) break;
    // 16 CTAs: allowed and placeable?  (asked once per kernel variant and device)
    static int ok16[3][64];                          // [variant][device]: 0 unknown, 1 yes, -1 no
    int dev = 0;
    cudaGetDevice(&dev);
    int& ok = ok16[nr == 2 ? 0 : (nr == 4 ? 1 : 2)][dev & 63];
    if (ok == 0) {
      int nclusters = 0;
      cudaError_t q = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      if (q == cudaSuccess) q = cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg);
      if (q != cudaSuccess) cudaGetLastError();
      ok = (q == cudaSuccess && nclusters >= 1) ? 1 : -1;
    }
    if (ok == 1) break;
  }
  // CORRLA_B200_TEST_JACOBI_SPIN_LIMIT: test hook -- 0 forces the failure path from the start
  static const unsigned spin_limit = [] { const char* e = getenv("CORRLA_B200_TEST_JACOBI_SPIN_LIMIT"); return e == nullptr ? kRJSpinLimit : (unsigned)strtoul(e, nullptr, 10); }();
  // CORRLA_B200_TEST_JACOBI_FAIL_SWEEP=s: test hook -- one warp gives up at the start of sweep s (failure from inside the iteration)
  static const int fail_at = [] { const char* e = getenv("CORRLA_B200_TEST_JACOBI_FAIL_SWEEP"); return e == nullptr ? -1 : atoi(e); }();
  e = cudaLaunchKernelEx(&cfg, kern, W, ldw, l, sigma, Vr, Ur, Lrows, ldo, transpose, info, spin_limit, dbg_on ? dbg_dev : nullptr, fail_at);
  if (dbg_on && e == cudaSuccess && cudaStreamSynchronize(s) == cudaSuccess) {
    static long long hb[4 * 128];
    if (cudaMemcpy(hb, dbg_dev, sizeof(hb), cudaMemcpyDeviceToHost) == cudaSuccess) {
      fprintf(stderr, "[jacobi_ring] l=%d cluster=%d warps/CTA=%d; cycles per warp in the step loops (rotate, send, wait, total):\n", l, C, wpc);
      for (int g = 0; g < h; ++g)
        if (g < 2 || g >= h - 2 || g % wpc == 0 || g % wpc == wpc - 1 || g == h / 2)
          fprintf(stderr, "  warp %3d: %9lld %9lld %9lld %9lld\n", g, hb[4 * g], hb[4 * g + 1], hb[4 * g + 2], hb[4 * g + 3]);
    }
  }
  return e;
}

}  // namespace corrla
