// Skinny f64 GEMM: persistent, warp-specialised (1 TMA producer warp + 8 DMMA warps), split-K with a
// deterministic two-stage reduction.  See skinny_gemm.cuh for the contract.
//
// Shared-memory tile formats (both written by TMA with CU_TENSOR_MAP_SWIZZLE_128B, 128-byte rows):
//   KC ("K-contiguous", reduce_inner): one box [128 Mside-rows][16 k]; element (r,k) lives at
//        r*128 + (((k>>1) ^ (r&7)) << 4) + (k&1)*8.
//        MMA row g of an 8-row block reads tile row 2*(g&3) + (g>>2): each half-warp then touches
//        rows {0,2,4,6} or {1,3,5,7}, whose XOR patterns map the 16 lanes onto 16 distinct 8-byte
//        bank pairs -> conflict-free LDS.64.
//   MC ("M-contiguous", reduce_outer): eight boxes [16 k][16 Mside-cols] of 2 KB; element (k,c) of a box at
//        k*128 + (((c>>1) ^ (k&7)) << 4) + (c&1)*8.
//        MMA row g of m-block `mbl` (two per box) reads box column (g&1) + 8*((g>>1)&1) + 2*(g>>2) + 4*mbl,
//        again 16 distinct bank pairs per half-warp.
// The output rows follow the same permutations, so no data is ever shuffled.
// B slab: [16 k][ldb] row-major, ldb == 4 (mod 8): lane (g,t) reads word (4s+t)*2*ldb + 2*(8*nb+g);
// t*2*ldb mod 32 is {0,8,16,24} (or its mirror) -> conflict-free.
#include "skinny_gemm.cuh"
#include "ptx.cuh"

#include <algorithm>
#include <cstdio>
#include <mutex>

namespace corrla {

namespace {

constexpr uint32_t kAStageBytes = kTileM * kChunkK * 8;  // 16 KB

__device__ __forceinline__ int rowperm(int g) { return 2 * (g & 3) + (g >> 2); }
__device__ __forceinline__ int colperm(int g, int mbl) {
  return (g & 1) + 8 * ((g >> 1) & 1) + 2 * (g >> 2) + 4 * mbl;
}


// Register budget: 12 warps are launched (two consumer warpgroups + one producer warpgroup); the producer group gives
// its registers back (setmaxnreg.dec) and the consumers grow to 232, which holds 112 accumulator registers plus two
// sets of fragments without spilling.
constexpr int kConsumerRegs = 232;
constexpr int kProducerRegs = 40;
constexpr int kThreads = 12 * 32;

// DMMAs of one k4-step for the n-block slots j >= J0 only.  The triangular modes skip whole n-blocks; the skip has to be
// a real (warp-uniform) branch -- as a predicate on each DMMA the skipped instructions still take their issue slots and
// the mode runs at the dense rate (measured: 459 us against 250 us of kept work for the 512k x 112 triangular apply).
template <int MI, int JW, int J0, bool FULL>
__device__ __forceinline__ void dmma_step_from(double (&acc)[MI][JW][2], const double (&af)[MI], const double (&bf)[JW], int jn) {
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = J0; j < JW; ++j)
      if (FULL || j < jn) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
}
template <int MI, int JW, bool FULL>
__device__ __forceinline__ void dmma_step_skip(int j0, double (&acc)[MI][JW][2], const double (&af)[MI], const double (&bf)[JW], int jn) {
  switch (j0) {
    case 0: dmma_step_from<MI, JW, 0, FULL>(acc, af, bf, jn); break;
    case 1: if (JW > 1) dmma_step_from<MI, JW, (JW > 1 ? 1 : 0), FULL>(acc, af, bf, jn); break;
    case 2: if (JW > 2) dmma_step_from<MI, JW, (JW > 2 ? 2 : 0), FULL>(acc, af, bf, jn); break;
    case 3: if (JW > 3) dmma_step_from<MI, JW, (JW > 3 ? 3 : 0), FULL>(acc, af, bf, jn); break;
    case 4: if (JW > 4) dmma_step_from<MI, JW, (JW > 4 ? 4 : 0), FULL>(acc, af, bf, jn); break;
    case 5: if (JW > 5) dmma_step_from<MI, JW, (JW > 5 ? 5 : 0), FULL>(acc, af, bf, jn); break;
    case 6: if (JW > 6) dmma_step_from<MI, JW, (JW > 6 ? 6 : 0), FULL>(acc, af, bf, jn); break;
    case 7: if (JW > 7) dmma_step_from<MI, JW, (JW > 7 ? 7 : 0), FULL>(acc, af, bf, jn); break;
    default: break;                                       // every n-block of this warp is skipped
  }
}

// (The symmetric mode keeps per-DMMA predicates: its skip pattern differs per accumulator row, and one switch per row
// and step measured slower -- 445 us against 262 us for the 512k x 112 Gram matrix -- than the predicated form.)
// One consumer warp, all work items of this CTA.  FULL: every n-block of the warp is valid (no predicates on the DMMAs).
// MODE 0: dense.  MODE 1 (MC only): the output is symmetric (Gram matrix) and only its upper triangle is consumed, so
// 8x8 blocks entirely below the diagonal are skipped and m-blocks are dealt round-robin to the warps to balance what
// is left.  MODE 2 (KC only): B is upper triangular (a Cholesky/Householder inverse), so k4-steps whose rows of B are
// zero for a given n-block are skipped.  n-blocks are always interleaved between the WN warps (nb = wn + WN*j), which
// balances both triangular cases.
template <bool KC, int WM, int WN, int MI, int JW, bool FULL, int MODE>
__device__ __forceinline__ void consumer_loop(const GemmArgs& p, unsigned char* smem, unsigned char* smemB, uint32_t sBar,
                                              double* red, int warp, int lane, int jn) {
  constexpr int NCW = WM * WN;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp % WM, wn = warp / WM;

  // A-fragment byte offsets inside a stage
  int a_base;          // KC: row part for i = 0;   MC: box part for i = 0
  int a_xo[2][4];      // KC: [0][s];  MC: [i&1][s]
  if (KC) {
    const int rp = rowperm(g);
    a_base = (wm * MI * 8 + rp) * 128;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      a_xo[0][s] = ((((s << 1) | (t >> 1)) ^ rp) << 4) + ((t & 1) << 3);
      a_xo[1][s] = a_xo[0][s];
    }
  } else {
    a_base = (MODE == 1) ? (wm >> 1) * 2048 : ((wm * MI) >> 1) * 2048;
#pragma unroll
    for (int par = 0; par < 2; ++par) {
      const int mc = colperm(g, par);
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const int kr = 4 * s + t;
        a_xo[par][s] = kr * 128 + ((((mc >> 1) ^ (kr & 7))) << 4) + ((mc & 1) << 3);
      }
    }
  }
  const int b_off = t * p.ldb * 8 + (8 * wn + g) * 8;      // n-block of slot j: nb = wn + WN*j
  const int b_step = 4 * p.ldb * 8;

  double alpha = 1.0;
  if (p.alpha_sumsq != nullptr) alpha = rsqrt(*p.alpha_sumsq);

  auto load_frags = [&](double (&af)[MI], double (&bf)[JW], const unsigned char* a, const unsigned char* b, int s) {
#pragma unroll
    for (int i = 0; i < MI; ++i) {
      const int off = KC ? (i * 1024 + a_xo[0][s])
                         : (MODE == 1 ? (2 * i * 2048 + a_xo[wm & 1][s]) : ((i >> 1) * 2048 + a_xo[i & 1][s]));
      af[i] = *reinterpret_cast<const double*>(a + off);
    }
#pragma unroll
    for (int j = 0; j < JW; ++j)
      if (FULL || j < jn) bf[j] = *reinterpret_cast<const double*>(b + s * b_step + j * (WN * 64));
  };
  // does DMMA (i, j) of k4-step s in chunk c contribute?  (warp-uniform)
  auto keep = [&](int i, int j, int s, int64_t c) -> bool {
    if (MODE == 1) return (wn + WN * j) >= 2 * ((wm >> 1) + 2 * i);          // n-block reaches the diagonal of box(i)
    if (MODE == 2) return (int64_t)(wn + WN * j) >= 2 * c + (s >= 2 ? 1 : 0); // rows of B in this step are not all zero
    return true;
  };
  // tile row of accumulator slot i for lane group g
  auto tile_row = [&](int i) -> int {
    if (KC) return (wm * MI + i) * 8 + rowperm(g);
    if (MODE == 1) return 16 * ((wm >> 1) + 2 * i) + colperm(g, wm & 1);
    return 16 * ((wm * MI + i) >> 1) + colperm(g, i & 1);
  };

  const int W = p.tilesM * p.splits;
  uint32_t stage = 0, phase = 0;
  double acc[MI][JW][2];
  double af[2][MI], bf[2][JW];

  for (int w = blockIdx.x; w < W; w += gridDim.x) {
    const int split = w / p.tilesM, tile = w - split * p.tilesM;
    const int64_t c0 = (int64_t)split * p.chunks_per_split;
    const int64_t c1 = min(c0 + p.chunks_per_split, p.chunks_total);
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
      for (int j = 0; j < JW; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

    if (c0 < c1) {
      mbar_wait(sBar + 8 * stage, phase);
      load_frags(af[0], bf[0], smem + stage * kAStageBytes + a_base, smemB + stage * p.b_stage_bytes + b_off, 0);
    }
    for (int64_t c = c0; c < c1; ++c) {
      const unsigned char* a = smem + stage * kAStageBytes + a_base;
      const unsigned char* b = smemB + stage * p.b_stage_bytes + b_off;
      uint32_t nstage = stage + 1, nphase = phase;
      if (nstage == (uint32_t)p.stages) { nstage = 0; nphase ^= 1u; }
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        if (s < 3) {
          load_frags(af[(s + 1) & 1], bf[(s + 1) & 1], a, b, s + 1);
        } else if (c + 1 < c1) {
          // fragments of the next chunk's first step are fetched before this chunk's last DMMAs are issued
          mbar_wait(sBar + 8 * nstage, nphase);
          load_frags(af[0], bf[0], smem + nstage * kAStageBytes + a_base, smemB + nstage * p.b_stage_bytes + b_off, 0);
        }
        if (MODE == 2) {
          // rows 16c + 4s .. of the upper-triangular B are zero for the n-blocks nb = wn + WN*j < 2c + (s >= 2)
          const int kb = (int)(2 * c) + (s >= 2 ? 1 : 0);
          const int j0 = kb > wn ? (kb - wn + WN - 1) / WN : 0;
          dmma_step_skip<MI, JW, FULL>(j0, acc, af[s & 1], bf[s & 1], jn);
        } else {
#pragma unroll
          for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < JW; ++j)
              if ((FULL || j < jn) && (MODE == 0 || keep(i, j, s, c)))
                dmma_m8n8k4(acc[i][j][0], acc[i][j][1], af[s & 1][i], bf[s & 1][j]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(sBar + 8 * (8 + stage));
      stage = nstage; phase = nphase;
    }

    // ------------------------------ epilogue ------------------------------
    const int Lc = p.nblk * 8;
    double ss = 0.0;
    if (p.splits > 1) {
      double* wsp = p.ws + (int64_t)w * kTileM * Lc;
#pragma unroll
      for (int i = 0; i < MI; ++i) {
        const int tr = tile_row(i);
#pragma unroll
        for (int j = 0; j < JW; ++j)
          if (FULL || j < jn) {
            const int col = 8 * (wn + WN * j) + 2 * t;
            *reinterpret_cast<double2*>(wsp + (int64_t)tr * Lc + col) = make_double2(acc[i][j][0], acc[i][j][1]);
          }
      }
    } else {
#pragma unroll
      for (int i = 0; i < MI; ++i) {
        const int tr = tile_row(i);
        const int64_t row = (int64_t)tile * kTileM + tr;
        if (row < p.Mside) {
#pragma unroll
          for (int j = 0; j < JW; ++j)
            if (FULL || j < jn) {
              const int col = 8 * (wn + WN * j) + 2 * t;
              double v0 = acc[i][j][0], v1 = acc[i][j][1];
              if (p.col_bias != nullptr) { v0 -= p.col_bias[col]; v1 -= p.col_bias[col + 1]; }
              v0 *= alpha; v1 *= alpha;
              double* o = p.out + row * p.out_rs + (int64_t)col * p.out_cs;
              if (p.accumulate) {
                if (col < p.ncols_out) v0 += o[0];
                if (col + 1 < p.ncols_out) v1 += o[p.out_cs];
              }
              ss += v0 * v0 + v1 * v1;
              if (p.out_cs == 1 && (p.out_rs & 1) == 0 && col + 1 < p.ncols_out) {
                *reinterpret_cast<double2*>(o) = make_double2(v0, v1);
              } else {
                if (col < p.ncols_out) o[0] = v0;
                if (col + 1 < p.ncols_out) o[p.out_cs] = v1;
              }
            }
        }
      }
      if (p.sumsq_partials != nullptr) {
        // one partial per (work item, warp): no CTA-wide barrier in the epilogue; sumsq_finalize adds them in order
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        if (lane == 0) p.sumsq_partials[(int64_t)w * NCW + warp] = ss;
      }
    }
  }
}

// reduce_outer, narrow sketch (8x1 warp layout) and a short output side: see GemmArgs::kgroups.  Warp w works on box
// column w % a_boxes of sub-chunk w / a_boxes of every stage.  Every group's partial tile (16*a_boxes rows) goes to the
// split-K workspace as if it were one more split.
template <int JW, bool FULL>
__device__ __forceinline__ void consumer_loop_mc_kgroups(const GemmArgs& p, unsigned char* smem, unsigned char* smemB,
                                                         uint32_t sBar, int warp, int lane, int jn) {
  const int g = lane >> 2, t = lane & 3;
  const int kg = p.kgroups, bv = p.a_boxes;
  const int box = warp % bv, kgrp = warp / bv;
  int a_xo[2][4];
#pragma unroll
  for (int par = 0; par < 2; ++par) {
    const int mc = colperm(g, par);
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int kr = 4 * s + t;
      a_xo[par][s] = kr * 128 + ((((mc >> 1) ^ (kr & 7))) << 4) + ((mc & 1) << 3);
    }
  }
  const int a_base = (kgrp * bv + box) * 2048;
  const int b_off = (16 * kgrp + t) * p.ldb * 8 + g * 8;
  const int b_step = 4 * p.ldb * 8;
  const int W = p.tilesM * p.splits;
  const int Lc = p.nblk * 8;
  uint32_t stage = 0, phase = 0;
  double acc[2][JW][2];
  for (int w = blockIdx.x; w < W; w += gridDim.x) {
    const int split = w / p.tilesM;
    const int64_t c0 = (int64_t)split * p.chunks_per_split;
    const int64_t c1 = min(c0 + p.chunks_per_split, p.chunks_total);
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < JW; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
    for (int64_t c = c0; c < c1; ++c) {
      mbar_wait(sBar + 8 * stage, phase);
      // my sub-chunk may lie beyond the end of the reduction axis in the last stage: its rows of B were not loaded
      if ((c * kg + kgrp) * kChunkK < p.k16) {
        const unsigned char* a = smem + stage * kAStageBytes + a_base;
        const unsigned char* b = smemB + stage * p.b_stage_bytes + b_off;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          const double a0 = *reinterpret_cast<const double*>(a + a_xo[0][s]);
          const double a1 = *reinterpret_cast<const double*>(a + a_xo[1][s]);
#pragma unroll
          for (int j = 0; j < JW; ++j)
            if (FULL || j < jn) {
              const double bf = *reinterpret_cast<const double*>(b + s * b_step + j * 64);
              dmma_m8n8k4(acc[0][j][0], acc[0][j][1], a0, bf);
              dmma_m8n8k4(acc[1][j][0], acc[1][j][1], a1, bf);
            }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(sBar + 8 * (8 + stage));
      if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1u; }
    }
    double* wsp = p.ws + ((int64_t)w * kg + kgrp) * (16 * bv) * Lc;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int tr = 16 * box + colperm(g, i);
#pragma unroll
      for (int j = 0; j < JW; ++j)
        if (FULL || j < jn)
          *reinterpret_cast<double2*>(wsp + (int64_t)tr * Lc + 8 * j + 2 * t) = make_double2(acc[i][j][0], acc[i][j][1]);
    }
  }
}

template <bool KC, int WM, int WN, int MI, int JW, int MODE>
__global__ void __launch_bounds__(kThreads, 1)
skinny_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const GemmArgs p) {
  static_assert(WM * MI * 8 == kTileM, "tile rows");
  static_assert(MI % 2 == 0, "MI even (box parity)");
  static_assert(WM * WN == 8, "two consumer warpgroups");
  constexpr int NCW = WM * WN;

  if (p.cond_flag != nullptr && *p.cond_flag == 0) return;

  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t pad = ((raw_u32 + 1023u) & ~1023u) - raw_u32;
  unsigned char* smem = smem_raw + pad;
  const uint32_t sA = raw_u32 + pad;
  const uint32_t sB = sA + p.stages * kAStageBytes;
  const uint32_t sBar = sB + p.stages * p.b_stage_bytes;     // full[stages], empty[stages]
  unsigned char* smemB = smem + p.stages * kAStageBytes;
  double* red = reinterpret_cast<double*>(smemB + p.stages * p.b_stage_bytes + 2 * 8 * 8);  // after 16 barriers

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(sBar + 8 * s, 1);
      mbar_init(sBar + 8 * (8 + s), NCW);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp >= NCW) {
    // ------------------------------ TMA producer warpgroup ------------------------------
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kProducerRegs));
    if (warp != NCW || lane != 0) return;
    tma_prefetch_desc(&tmA);
    const int W = p.tilesM * p.splits;
    uint32_t stage = 0, phase = 0;
    const uint32_t tx = (KC ? kAStageBytes : (uint32_t)(p.kgroups * p.a_boxes) * 2048u) + p.b_stage_bytes;
    for (int w = blockIdx.x; w < W; w += gridDim.x) {
      const int split = w / p.tilesM, tile = w - split * p.tilesM;
      const int64_t c0 = (int64_t)split * p.chunks_per_split;
      const int64_t c1 = min(c0 + p.chunks_per_split, p.chunks_total);
      const int m0 = tile * kTileM;
      for (int64_t c = c0; c < c1; ++c) {
        mbar_wait(sBar + 8 * (8 + stage), phase ^ 1u);
        const uint32_t full = sBar + 8 * stage;
        const int k0 = (int)(c * kChunkK * p.kgroups);
        const uint32_t dstA = sA + stage * kAStageBytes;
        if (KC) {
          mbar_arrive_expect_tx(full, tx);
          tma_load_2d(dstA, &tmA, k0, m0, full);
          bulk_load(sB + stage * p.b_stage_bytes, p.B + (int64_t)k0 * p.ldb, p.b_stage_bytes, full);
        } else if (p.kgroups == 1) {
          mbar_arrive_expect_tx(full, tx);
#pragma unroll
          for (int b = 0; b < 8; ++b)
            if (b < p.a_boxes) tma_load_2d(dstA + b * 2048, &tmA, m0 + 16 * b, k0, full);
          bulk_load(sB + stage * p.b_stage_bytes, p.B + (int64_t)k0 * p.ldb, p.b_stage_bytes, full);
        } else {
          // kgroups sub-chunks of 16 reduction rows: boxes [sub-chunk][box column]; only the rows of B that exist
          const int64_t rows_b = min((int64_t)kChunkK * p.kgroups, p.k16 - (int64_t)k0);
          const uint32_t b_bytes = (uint32_t)(rows_b * p.ldb * 8);
          mbar_arrive_expect_tx(full, (uint32_t)(p.kgroups * p.a_boxes) * 2048u + b_bytes);
          for (int kb = 0; kb < p.kgroups; ++kb)
            for (int cb = 0; cb < p.a_boxes; ++cb)
              tma_load_2d(dstA + (kb * p.a_boxes + cb) * 2048, &tmA, m0 + 16 * cb, k0 + kChunkK * kb, full);
          bulk_load(sB + stage * p.b_stage_bytes, p.B + (int64_t)k0 * p.ldb, b_bytes, full);
        }
        if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1u; }
      }
    }
    return;
  }

  // ------------------------------ DMMA consumer warpgroups ------------------------------
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kConsumerRegs));
  const int jn = min(JW, (p.nblk - (warp / WM) + WN - 1) / WN);   // valid n-block slots of this warp (nb = wn + WN*j)
  if (!KC && MODE == 0 && WN == 1 && p.kgroups > 1) {
    if (jn == JW) consumer_loop_mc_kgroups<JW, true>(p, smem, smemB, sBar, warp, lane, jn);
    else consumer_loop_mc_kgroups<JW, false>(p, smem, smemB, sBar, warp, lane, jn);
    return;
  }
  if (jn == JW) consumer_loop<KC, WM, WN, MI, JW, true, MODE>(p, smem, smemB, sBar, red, warp, lane, jn);
  else consumer_loop<KC, WM, WN, MI, JW, false, MODE>(p, smem, smemB, sBar, red, warp, lane, jn);
}

// out(r, c) = alpha * sum_s ws[s][tile(r)][r % 128][c]; one thread per column pair.
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const double* __restrict__ ws, int splits, int64_t split_stride, int Lc, int64_t Mside,
                     double* __restrict__ out, int64_t out_rs, int64_t out_cs, int ncols_out,
                     const double* alpha_sumsq, const double* col_bias, double* sumsq_partials, const int* cond_flag,
                     int accumulate) {
  if (cond_flag != nullptr && *cond_flag == 0) return;
  __shared__ double red[8];
  const int half = Lc >> 1;
  const int64_t total = Mside * half;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double alpha = 1.0;
  if (alpha_sumsq != nullptr) alpha = rsqrt(*alpha_sumsq);
  double ss = 0.0;
  if (idx < total) {
    const int64_t r = idx / half;
    const int col = (int)(idx - r * half) * 2;
    const double* src = ws + r * Lc + col;                 // (tile*128 + row in tile) * Lc == r * Lc
    double s0 = 0.0, s1 = 0.0;
    for (int s = 0; s < splits; ++s) {
      const double2 v = *reinterpret_cast<const double2*>(src + s * split_stride);
      s0 += v.x; s1 += v.y;
    }
    if (col_bias != nullptr) { s0 -= col_bias[col]; s1 -= col_bias[col + 1]; }
    s0 *= alpha; s1 *= alpha;
    double* o = out + r * out_rs + (int64_t)col * out_cs;
    if (accumulate) {
      if (col < ncols_out) s0 += o[0];
      if (col + 1 < ncols_out) s1 += o[out_cs];
    }
    ss = s0 * s0 + s1 * s1;
    if (col < ncols_out) o[0] = s0;
    if (col + 1 < ncols_out) o[out_cs] = s1;
  }
  if (sumsq_partials != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += red[i];
      sumsq_partials[blockIdx.x] = tot;
    }
  }
}

// The same reduction for MANY splits and a small output (a 64 x 24 product of a million-row contraction leaves ~600
// partial tiles): one warp per output column pair, lane s adds splits s, s + 32, ... and a fixed butterfly adds the
// lanes -- still one summation order, so still bit-reproducible, but ~20 dependent loads deep instead of ~600.
__global__ void __launch_bounds__(256)
splitk_reduce_warp_kernel(const double* __restrict__ ws, int splits, int64_t split_stride, int Lc, int64_t Mside,
                          double* __restrict__ out, int64_t out_rs, int64_t out_cs, int ncols_out,
                          const double* alpha_sumsq, const double* col_bias, double* sumsq_partials, const int* cond_flag,
                          int accumulate) {
  if (cond_flag != nullptr && *cond_flag == 0) return;
  __shared__ double red[8];
  const int half = Lc >> 1;
  const int64_t total = Mside * half;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t idx = (int64_t)blockIdx.x * 8 + wib;
  double alpha = 1.0;
  if (alpha_sumsq != nullptr) alpha = rsqrt(*alpha_sumsq);
  double ss = 0.0;
  if (idx < total) {
    const int64_t r = idx / half;
    const int col = (int)(idx - r * half) * 2;
    const double* src = ws + r * Lc + col;
    double s0 = 0.0, s1 = 0.0;
    for (int s = lane; s < splits; s += 32) {
      const double2 v = *reinterpret_cast<const double2*>(src + s * split_stride);
      s0 += v.x; s1 += v.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o);
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    if (lane == 0) {
      if (col_bias != nullptr) { s0 -= col_bias[col]; s1 -= col_bias[col + 1]; }
      s0 *= alpha; s1 *= alpha;
      double* o = out + r * out_rs + (int64_t)col * out_cs;
      if (accumulate) {
        if (col < ncols_out) s0 += o[0];
        if (col + 1 < ncols_out) s1 += o[out_cs];
      }
      ss = s0 * s0 + s1 * s1;
      if (col < ncols_out) o[0] = s0;
      if (col + 1 < ncols_out) o[out_cs] = s1;
    }
  }
  if (sumsq_partials != nullptr) {
    if (lane == 0) red[wib] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int i = 0; i < 8; ++i) tot += red[i];
      sumsq_partials[blockIdx.x] = tot;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Split-K reduction fused with the cross-GPU all-reduce over NVLink peer memory (one process per GPU).
//   phase 0  local:   mine[r*ld + c] = sum_s ws[s][...]          (ws == nullptr: the GEMM already wrote `mine`)
//                     the x_extra caller values are copied from out's tail into mine's tail
//   publish           last block to finish: st.release.sys of this epoch into every rank's flag array
//   acquire           every block spins (bounded) until all ranks have published this epoch
//   phase 1  global:  out[i] = sum over ranks, in rank order, of peer[rank][i]   (i < x_count) -- identical bits on
//                     every rank, read straight out of the peers' HBM through NVLink (ld.relaxed.sys, not L1-cached)
// The two halves of the symmetric region alternate by epoch parity, which makes a second barrier unnecessary: a rank
// can only overwrite half h two epochs later, after it has seen every peer publish the epoch in between, i.e. after
// every peer finished reading.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double2 ld_relaxed_sys_f64x2(const double* p) {
  double2 v;
  asm volatile("ld.relaxed.sys.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
  return v;
}

// A rank that has seen a timeout publishes its later epochs with the poison bit set: every peer that reads such a flag
// raises its own error flag (and stops waiting), so one lost rank makes the call fail on ALL ranks instead of leaving
// some of them with a silently wrong sum.
constexpr unsigned long long kEpochPoison = 1ull << 62;
__device__ __forceinline__ unsigned long long publish_word(const PeerExchange& px) {
  return px.epoch | ((*reinterpret_cast<volatile int*>(px.err) != 0) ? kEpochPoison : 0ull);
}
// wait (bounded) until rank `src` has published this epoch
__device__ __forceinline__ void wait_for_peer(const PeerExchange& px, int src) {
  const unsigned long long* f = px.my_flags + src;
  const long long t0 = clock64();
  for (;;) {
    const unsigned long long v = ld_acquire_sys_u64(f);
    if (v & kEpochPoison) { *px.err = 1; break; }
    if (v >= px.epoch) break;
    if (clock64() - t0 > px.timeout_cycles) { *px.err = 1; break; }   // default ~10 s: the peer never arrived
    __nanosleep(64);
  }
}

__global__ void __launch_bounds__(256)
reduce_exchange_kernel(const double* __restrict__ ws, int splits, int64_t split_stride, int Lc, int64_t Mside, int64_t ld,
                       double* __restrict__ out, size_t x_count, size_t x_extra, const PeerExchange px,
                       const int* cond_flag) {
  // A conditionally skipped exchange (every rank sees the same flag: it derives from all-reduced data) still takes part
  // in the epoch protocol: block 0 publishes this epoch and waits for the peers', so that the half of the symmetric
  // region used two epochs apart is never overwritten while a slower peer is still summing it.  No data moves.
  const bool skip = (cond_flag != nullptr && *cond_flag == 0);
  if (skip && blockIdx.x != 0) return;
  __shared__ int s_last;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  double* mine = px.mine;
  if (skip) {
    if (threadIdx.x < px.nranks) {
      st_release_sys_u64(px.peer_flags[threadIdx.x] + px.rank, publish_word(px));
      wait_for_peer(px, threadIdx.x);
    }
    return;
  }

  // ---- phase 0: the whole ld-pitched block is (re)written, pads included, so nothing stale is ever summed
  {
    const int64_t ld2 = ld >> 1;
    const int64_t body_pairs = (int64_t)((x_count - x_extra) >> 1);
    for (int64_t idx = tid; idx < body_pairs; idx += nthreads) {
      const int64_t r = idx / ld2;
      const int col = (int)(idx - r * ld2) * 2;
      const bool valid = (r < Mside) && (col < Lc);
      if (ws != nullptr) {
        double s0 = 0.0, s1 = 0.0;
        if (valid) {
          const double* src = ws + r * Lc + col;              // (tile*128 + rt) * Lc == r * Lc
          for (int s = 0; s < splits; ++s) {
            const double2 v = *reinterpret_cast<const double2*>(src + s * split_stride);
            s0 += v.x; s1 += v.y;
          }
        }
        *reinterpret_cast<double2*>(mine + 2 * idx) = make_double2(s0, s1);
      } else if (!valid) {
        *reinterpret_cast<double2*>(mine + 2 * idx) = make_double2(0.0, 0.0);   // the GEMM wrote the valid part itself
      }
    }
  }
  for (size_t i = (size_t)tid; i < x_extra; i += (size_t)nthreads) mine[x_count - x_extra + i] = out[x_count - x_extra + i];

  // ---- publish: every thread's stores are fenced system-wide, the last block raises the flags
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int ticket = atomicAdd(px.block_counter, 1u);
    s_last = (ticket == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (s_last) {
    if (threadIdx.x < px.nranks) st_release_sys_u64(px.peer_flags[threadIdx.x] + px.rank, publish_word(px));
    if (threadIdx.x == 0) *px.block_counter = 0u;               // ready for the next launch (stream ordered)
  }

  // ---- acquire: wait (bounded) for every rank's flag
  if (threadIdx.x < px.nranks) wait_for_peer(px, threadIdx.x);
  __syncthreads();

  // ---- phase 1
  const size_t pairs = x_count >> 1;
  for (size_t i = (size_t)tid; i < pairs; i += (size_t)nthreads) {
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r)
      if (r < px.nranks) {
        const double2 v = (r == px.rank) ? *reinterpret_cast<const double2*>(mine + 2 * i)
                                         : ld_relaxed_sys_f64x2(px.peer[r] + 2 * i);
        s0 += v.x; s1 += v.y;
      }
    *reinterpret_cast<double2*>(out + 2 * i) = make_double2(s0, s1);
  }
  if ((x_count & 1) && tid == 0) {
    double s0 = 0.0;
    for (int r = 0; r < px.nranks; ++r) {
      const double* q = (r == px.rank) ? mine : px.peer[r];
      double v;
      asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(q + x_count - 1) : "memory");
      s0 += v;
    }
    out[x_count - 1] = s0;
  }
}

// *slot = sum(partials[0..n)) in a fixed order (bit-reproducible).
__global__ void __launch_bounds__(1024)
sumsq_finalize_kernel(const double* __restrict__ partials, int64_t n, double* slot, const int* cond_flag) {
  if (cond_flag != nullptr && *cond_flag == 0) return;
  __shared__ double red[32];
  // four independent accumulators per thread: the loads of a trip are in flight together (one CTA sums up to 262 144
  // partials; a single dependent chain of loads made this kernel 40 us at C3).  The order stays fixed.
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  const int64_t bd = blockDim.x;
  int64_t i = threadIdx.x;
  for (; i + 3 * bd < n; i += 4 * bd) {
    s0 += partials[i]; s1 += partials[i + bd]; s2 += partials[i + 2 * bd]; s3 += partials[i + 3 * bd];
  }
  for (; i < n; i += bd) s0 += partials[i];
  double s = (s0 + s1) + (s2 + s3);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int i = 0; i < 32; ++i) tot += red[i];
    *slot = tot;
  }
}

// ----------------------------------- host side -----------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}

bool encode_map(CUtensorMap* tm, const MatView& v, bool kc) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return false;
  cuuint64_t dims[2] = {(cuuint64_t)v.inner, (cuuint64_t)v.outer};
  cuuint64_t strides[1] = {(cuuint64_t)v.ld * 8};
  cuuint32_t box[2] = {16, kc ? (cuuint32_t)kTileM : (cuuint32_t)kChunkK};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(v.p), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "corrla: cuTensorMapEncodeTiled failed (%d) p=%p inner=%lld outer=%lld ld=%lld\n", (int)r,
            (const void*)v.p, (long long)v.inner, (long long)v.outer, (long long)v.ld);
    return false;
  }
  return true;
}

typedef void (*KernelFn)(const CUtensorMap, const GemmArgs);

template <bool KC, int MODE>
KernelFn pick_kernel(int nblk, int* threads) {
  *threads = kThreads;
  switch (nblk) {
    case 8: return skinny_gemm_kernel<KC, 4, 2, 4, 4, MODE>;
    case 9: case 10: return skinny_gemm_kernel<KC, 4, 2, 4, 5, MODE>;
    case 11: case 12: return skinny_gemm_kernel<KC, 4, 2, 4, 6, MODE>;
    case 13: case 14: return skinny_gemm_kernel<KC, 4, 2, 4, 7, MODE>;
    case 15: case 16: return skinny_gemm_kernel<KC, 4, 2, 4, 8, MODE>;
    default: break;
  }
  if (MODE != 0) return nullptr;      // narrow outputs (8x1 warp layout): dense only
  switch (nblk) {
    case 1: return skinny_gemm_kernel<KC, 8, 1, 2, 1, 0>;
    case 2: return skinny_gemm_kernel<KC, 8, 1, 2, 2, 0>;
    case 3: return skinny_gemm_kernel<KC, 8, 1, 2, 3, 0>;
    case 4: return skinny_gemm_kernel<KC, 8, 1, 2, 4, 0>;
    case 5: return skinny_gemm_kernel<KC, 8, 1, 2, 5, 0>;
    case 6: return skinny_gemm_kernel<KC, 8, 1, 2, 6, 0>;
    case 7: return skinny_gemm_kernel<KC, 8, 1, 2, 7, 0>;
    default: return nullptr;
  }
}

constexpr size_t kMaxDynSmem = 227 * 1024;
constexpr size_t kWsCapBytes = (size_t)512 << 20;

}  // namespace

bool tma_compatible(const MatView& v) {
  return (reinterpret_cast<uintptr_t>(v.p) % 16 == 0) && (v.ld % 2 == 0) && v.inner > 0 && v.outer > 0 &&
         v.inner < ((int64_t)1 << 31) && v.outer < ((int64_t)1 << 31) && (v.ld * 8 < ((int64_t)1 << 40));
}

void gemm_plan(int64_t Mside, int64_t K, int nblk, int num_sms, int force_splits, int* tilesM_out, int* splits_out,
               int64_t* cps_out, size_t* ws_bytes, size_t* n_partials) {
  const int Lc = nblk * 8;
  const int64_t tilesM = (Mside + kTileM - 1) / kTileM;
  const int64_t chunks = std::max<int64_t>(1, (K + kChunkK - 1) / kChunkK);
  const double t_chunk = std::max(Lc * 0.016384, 0.37);  // us: DMMA time vs HBM time of one 16 KB A chunk
  const size_t tile_bytes = (size_t)kTileM * Lc * 8;
  int best_s = 1;
  double best_t = 1e300;
  const int64_t smax = force_splits > 0 ? force_splits : std::min<int64_t>(chunks, 1024);
  for (int64_t s = force_splits > 0 ? force_splits : 1; s <= smax; ++s) {
    const int64_t cps = (chunks + s - 1) / s;
    if (s > 1 && force_splits == 0 && cps < 8) break;
    if ((cps * (s - 1)) >= chunks && s > 1) continue;  // last split would be empty
    const int64_t items = tilesM * s;
    if (s > 1 && (size_t)items * tile_bytes > kWsCapBytes && force_splits == 0) break;
    const int64_t per_cta = (items + num_sms - 1) / num_sms;
    double t = per_cta * (cps + 3.0) * t_chunk;
    if (s > 1) t += (double)items * tile_bytes * 2.0 / 4.0e6 + 4.0;
    if (t < best_t * 0.999) { best_t = t; best_s = (int)s; }
  }
  const int64_t cps = (chunks + best_s - 1) / best_s;
  *tilesM_out = (int)tilesM;
  *splits_out = best_s;
  *cps_out = cps;
  *ws_bytes = best_s > 1 ? (size_t)tilesM * best_s * tile_bytes : 0;
  const int64_t reduce_blocks = (Mside * (Lc / 2) + 255) / 256;
  *n_partials = best_s > 1 ? (size_t)reduce_blocks : (size_t)tilesM * 8;   // 8 consumer warps per work item
  if (Mside <= 64 && nblk <= 7) {
    // a reduce_outer product this short runs in k-groups (GemmArgs::kgroups): it always goes through the workspace,
    // and its reduction uses one warp per output pair
    *ws_bytes = std::max(*ws_bytes, (size_t)best_s * tile_bytes);
    *n_partials = std::max(*n_partials, (size_t)((Mside * (Lc / 2) + 7) / 8));
  }
}

// The blocks of reduce_exchange_kernel wait for the flag their own LAST block publishes, so all of them must be
// resident at once: a cooperative launch guarantees that (or fails) even when other streams hold SMs.
static cudaError_t launch_reduce_exchange(int blocks, cudaStream_t stream, const double* ws, int splits, int64_t split_stride,
                                          int Lc, int64_t Mside, int64_t ld, double* out, size_t x_count, size_t x_extra,
                                          const PeerExchange& px, const int* cond_flag) {
  void* args[] = {(void*)&ws, (void*)&splits, (void*)&split_stride, (void*)&Lc, (void*)&Mside, (void*)&ld, (void*)&out,
                  (void*)&x_count, (void*)&x_extra, (void*)&px, (void*)&cond_flag};
  return cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(reduce_exchange_kernel), dim3((unsigned)blocks), dim3(256),
                                     args, 0, stream);
}

cudaError_t sum_array_launch(const double* partials, int64_t n, double* slot, cudaStream_t stream) {
  sumsq_finalize_kernel<<<1, 1024, 0, stream>>>(partials, n, slot, nullptr);
  return cudaGetLastError();
}

cudaError_t reduce_partials_launch(const double* ws, int splits, int rows_pad, int Lc, int64_t Mside, double* out,
                                   int64_t ld, const PeerExchange* px, size_t x_count, const int* cond_flag,
                                   int num_sms, cudaStream_t stream, int* launches) {
  const int tilesM = rows_pad / kTileM;
  if (px != nullptr) {
    if ((ld & 1) || x_count < (size_t)(Mside * ld) || (x_count & 1)) return cudaErrorInvalidValue;
    const int64_t work = std::max<int64_t>(Mside * (Lc / 2), (int64_t)(x_count / 2));
    const int blocks = (int)std::min<int64_t>((work + 255) / 256, num_sms);
    const cudaError_t ce = launch_reduce_exchange(blocks, stream, ws, splits, (int64_t)tilesM * kTileM * Lc, Lc, Mside, ld, out,
                                                  x_count, 0, *px, cond_flag);
    if (ce != cudaSuccess) return ce;
  } else {
    const int64_t total = Mside * (Lc / 2);
    const int blocks = (int)((total + 255) / 256);
    splitk_reduce_kernel<<<blocks, 256, 0, stream>>>(ws, splits, (int64_t)tilesM * kTileM * Lc, Lc, Mside, out, ld, 1, Lc,
                                                     nullptr, nullptr, nullptr, cond_flag, 0);
  }
  if (launches) ++*launches;
  return cudaGetLastError();
}

cudaError_t gemm_launch(const GemmCall& c, const GemmWorkspace& w, cudaStream_t stream, int* launches) {
  GemmArgs a{};
  const bool kc = c.reduce_inner;
  a.Mside = kc ? c.a.outer : c.a.inner;
  a.K = kc ? c.a.inner : c.a.outer;
  a.B = c.B; a.ldb = c.ldb; a.nblk = c.nblk;
  a.out = c.out; a.out_rs = c.out_rs; a.out_cs = c.out_cs; a.ncols_out = c.ncols_out;
  a.alpha_sumsq = c.alpha_sumsq;
  a.col_bias = c.col_bias;
  a.accumulate = c.accumulate ? 1 : 0;
  a.cond_flag = c.cond_flag;
  if (c.nblk < 1 || c.nblk > kMaxNblk || (c.ldb % 8) != 4 || c.ldb < c.nblk * 8) return cudaErrorInvalidValue;
  if (!tma_compatible(c.a)) return cudaErrorInvalidValue;

  size_t ws_bytes, n_partials;
  gemm_plan(a.Mside, a.K, a.nblk, w.num_sms, c.force_splits, &a.tilesM, &a.splits, &a.chunks_per_split, &ws_bytes,
            &n_partials);
  a.chunks_total = std::max<int64_t>(1, (a.K + kChunkK - 1) / kChunkK);
  a.k16 = (a.K + kChunkK - 1) / kChunkK * kChunkK;
  // short output side of a reduce_outer product with the narrow warp layout: k-groups (GemmArgs::kgroups)
  a.kgroups = 1; a.a_boxes = 8;
  if (!kc && a.nblk <= 7 && a.tilesM == 1) {
    a.a_boxes = (int)std::min<int64_t>(8, (a.Mside + 15) / 16);
    if (a.Mside <= 64) {
      a.a_boxes = a.Mside <= 16 ? 1 : (a.Mside <= 32 ? 2 : 4);
      a.kgroups = 8 / a.a_boxes;
      // a stage now spans kgroups sub-chunks: re-cut the reduction axis into stages, keep the planned number of splits
      a.chunks_total = std::max<int64_t>(1, (a.K + (int64_t)kChunkK * a.kgroups - 1) / ((int64_t)kChunkK * a.kgroups));
      a.splits = (int)std::min<int64_t>(a.splits, a.chunks_total);
      a.chunks_per_split = (a.chunks_total + a.splits - 1) / a.splits;
      a.splits = (int)((a.chunks_total + a.chunks_per_split - 1) / a.chunks_per_split);
      ws_bytes = (size_t)a.splits * kTileM * a.nblk * 8 * sizeof(double);      // splits * kgroups * (16 * a_boxes) rows
      n_partials = (size_t)((a.Mside * (a.nblk * 4) + 7) / 8);                 // blocks of the warp-parallel reduction
    }
  }
  const int splits_eff = a.splits * a.kgroups;                                 // partial tiles per output tile
  const int64_t split_stride = (a.kgroups > 1) ? (int64_t)16 * a.a_boxes * a.nblk * 8 : (int64_t)a.tilesM * kTileM * a.nblk * 8;
  if (ws_bytes > w.ws_bytes) return cudaErrorMemoryAllocation;
  if (c.sumsq_slot != nullptr && n_partials > w.n_partials) return cudaErrorMemoryAllocation;
  a.ws = w.ws;
  a.sumsq_partials = (c.sumsq_slot != nullptr && splits_eff == 1) ? w.sumsq_partials : nullptr;

  const bool fused_x = (c.px != nullptr);
  if (fused_x) {
    if (c.out_cs != 1 || c.alpha_sumsq != nullptr || c.col_bias != nullptr || c.sumsq_slot != nullptr || c.accumulate ||
        c.x_count < (size_t)(a.Mside * c.out_rs) + c.x_extra || ((c.x_count - c.x_extra) & 1) || (c.out_rs & 1) ||
        c.ncols_out != c.nblk * 8)
      return cudaErrorInvalidValue;
    if (splits_eff == 1) a.out = c.px->mine;      // the GEMM epilogue writes my half of the symmetric region directly
  }
  a.b_stage_bytes = (uint32_t)(kChunkK * a.kgroups * c.ldb * 8);
  const size_t per_stage = kAStageBytes + a.b_stage_bytes;
  const size_t fixed = 1024 + 16 * 8 + 16 * 8;  // alignment slack + barriers + reduction scratch
  a.stages = (int)std::min<size_t>(8, (kMaxDynSmem - fixed) / per_stage);
  if (a.stages < 2) return cudaErrorInvalidValue;
  const size_t smem = fixed + a.stages * per_stage;

  CUtensorMap tm;
  if (!encode_map(&tm, c.a, kc)) return cudaErrorInvalidValue;

  int threads = 0;
  KernelFn fn = nullptr;
  if (c.mode == 1 && !kc) fn = pick_kernel<false, 1>(a.nblk, &threads);
  else if (c.mode == 2 && kc) fn = pick_kernel<true, 2>(a.nblk, &threads);
  if (fn == nullptr) fn = kc ? pick_kernel<true, 0>(a.nblk, &threads) : pick_kernel<false, 0>(a.nblk, &threads);
  if (fn == nullptr) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(reinterpret_cast<const void*>(fn), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)kMaxDynSmem);
  if (e != cudaSuccess) return e;

  const int W = a.tilesM * a.splits;
  const int grid = std::min(W, w.num_sms);
  if (c.ev_begin) cudaEventRecord(c.ev_begin, stream);
  fn<<<grid, threads, smem, stream>>>(tm, a);
  if (c.ev_end) cudaEventRecord(c.ev_end, stream);
  if (launches) ++*launches;
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;

  int64_t n_part_used = (int64_t)a.tilesM * 8;
  if (fused_x) {
    const int Lc = a.nblk * 8;
    const int64_t work = std::max<int64_t>(a.Mside * (Lc / 2), (int64_t)(c.x_count / 2));
    const int blocks = (int)std::min<int64_t>((work + 255) / 256, w.num_sms);     // all co-resident: blocks spin
    const cudaError_t ce = launch_reduce_exchange(blocks, stream, splits_eff > 1 ? w.ws : nullptr, splits_eff, split_stride, Lc,
                                                  a.Mside, c.out_rs, c.out, c.x_count, c.x_extra, *c.px, a.cond_flag);
    if (launches) ++*launches;
    return ce != cudaSuccess ? ce : cudaGetLastError();
  }
  if (splits_eff > 1) {
    const int Lc = a.nblk * 8;
    const int64_t total = a.Mside * (Lc / 2);
    int blocks;
    if (splits_eff >= 64 && total <= 65536) {
      blocks = (int)((total + 7) / 8);                                          // one warp per output pair
      splitk_reduce_warp_kernel<<<blocks, 256, 0, stream>>>(w.ws, splits_eff, split_stride, Lc, a.Mside, a.out, a.out_rs,
                                                            a.out_cs, a.ncols_out, a.alpha_sumsq, a.col_bias,
                                                            c.sumsq_slot ? w.sumsq_partials : nullptr, a.cond_flag,
                                                            a.accumulate);
    } else {
      blocks = (int)((total + 255) / 256);
      splitk_reduce_kernel<<<blocks, 256, 0, stream>>>(w.ws, splits_eff, split_stride, Lc, a.Mside, a.out, a.out_rs, a.out_cs,
                                                       a.ncols_out, a.alpha_sumsq, a.col_bias,
                                                       c.sumsq_slot ? w.sumsq_partials : nullptr, a.cond_flag, a.accumulate);
    }
    if (launches) ++*launches;
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    n_part_used = blocks;
  }
  if (c.sumsq_slot != nullptr) {
    sumsq_finalize_kernel<<<1, 1024, 0, stream>>>(w.sumsq_partials, n_part_used, c.sumsq_slot, a.cond_flag);
    if (launches) ++*launches;
    e = cudaGetLastError();
  }
  return e;
}

}  // namespace corrla
