// Skinny f64 GEMM on the FP64 tensor pipe (DMMA.8x8x4) fed by TMA, for the RSVD passes.
//
//   D[Mside x Lc] (+)= op(A)[Mside x K] * B[K x Lc],   Lc = 8*nblk <= 128
//
// A is any 2-D f64 array in global memory seen as `outer` rows of `inner` contiguous elements
// (row pitch ld).  Two ways to contract it:
//   reduce_inner : Mside = outer, K = inner  (e.g. Y = A*Omega for row-major A;  A^T*Y for column-major A)
//   reduce_outer : Mside = inner, K = outer  (e.g. Z = A^T*Y for row-major A;  Gram Y^T*Y)
// B, and every matrix the engine owns, is row-major with pitch ldb = Lc + 4 (== 4 mod 8), zero padded,
// which makes the DMMA B-fragment loads bank-conflict free and lets a whole 16-row slab travel as
// one cp.async.bulk.
//
// This replaces faer's matmul behind par_matmul_helper (reference src/lib_math_utils/mat_utils.rs:20-33)
// at the call sites random_svd.rs:31, :42-51, :80, :92.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#include "comm.cuh"

namespace corrla {

constexpr int kTileM = 128;   // output rows per work item
constexpr int kChunkK = 16;   // reduction elements per pipeline stage
constexpr int kMaxNblk = 16;  // Lc <= 128

struct MatView {            // element (o, i) at p[o*ld + i]
  const double* p;
  int64_t inner, outer, ld;
};

struct GemmArgs {
  int64_t Mside, K;
  const double* B; int ldb; int nblk;
  // direct output (used when splits == 1) or final destination of the split-K reduction
  double* out; int64_t out_rs, out_cs; int ncols_out;
  // split-K
  double* ws; int tilesM, splits; int64_t chunks_total, chunks_per_split;
  // alpha = rsqrt(*alpha_sumsq) when non-null, else 1
  const double* alpha_sumsq;
  // out = alpha * (acc - col_bias[col]) when non-null (rank-1 centring correction: 1 * (mu^T X))
  const double* col_bias;
  int accumulate;          // out += alpha * (acc - bias) instead of out = ...
  // per-work-item (splits==1) sum of squares of the scaled output, or null
  double* sumsq_partials;
  // when non-null and *cond_flag == 0 the kernel does nothing
  const int* cond_flag;
  int stages; uint32_t b_stage_bytes;
  // reduce_outer with a SHORT output side (Mside <= 64, narrow sketches): only `a_boxes` of the eight 16-column TMA
  // boxes of a stage would hold rows of the output.  The other box slots then carry MORE of the reduction axis: a
  // stage is kgroups = 8 / a_boxes sub-chunks of 16 reduction rows (a_boxes boxes each, same 16 KB), the 8 consumer
  // warps form kgroups groups, group g contracts sub-chunk g of every stage, and each group writes its own partial
  // tile (summed by the split-K reduction).  No DMMA is spent on zero padding, all four FP64 pipes of the SM stay
  // busy, and the per-stage barrier traffic is amortised over kgroups times more data.  kgroups = 1, a_boxes = 8 otherwise.
  int kgroups, a_boxes;
  int64_t k16;             // reduction length rounded up to 16 (rows of B that exist)
};

// Host-side description of one product; see gemm_launch().
struct GemmCall {
  MatView a; bool reduce_inner;
  const double* B; int ldb; int nblk;
  double* out; int64_t out_rs, out_cs; int ncols_out;
  const double* alpha_sumsq = nullptr;
  const double* col_bias = nullptr;
  int mode = 0;                     // 1: symmetric output, only the upper triangle is needed (Gram; reduce_outer);
                                    // 2: B is upper triangular (reduce_inner).  Work below the diagonal is skipped.
  bool accumulate = false;          // out += product (used to refill deflated columns with fresh A*omega vectors)
  double* sumsq_slot = nullptr;     // if set: *sumsq_slot = sum of squares of the (scaled) output
  const int* cond_flag = nullptr;
  int force_splits = 0;             // testing hook: 0 = choose
  cudaEvent_t ev_begin = nullptr, ev_end = nullptr;   // optional: recorded around the DMMA kernel only
  // Fused all-reduce over peer memory (row-sharded runs): when px != nullptr the product is summed over the ranks
  // by the reduction kernel itself.  `out` must then be a contiguous row-major block (out_cs == 1) of x_count
  // doubles in total, of which the last x_extra are caller-provided local values already sitting at out + x_count -
  // x_extra (e.g. a squared norm) that ride along with the matrix.
  const PeerExchange* px = nullptr;
  size_t x_count = 0, x_extra = 0;
};

struct GemmWorkspace {
  double* ws = nullptr; size_t ws_bytes = 0;           // split-K partial tiles
  double* sumsq_partials = nullptr; size_t n_partials = 0;
  int num_sms = 148;
};

// Returns cudaSuccess or the first CUDA error; throws nothing.  All work is enqueued on `stream`.
cudaError_t gemm_launch(const GemmCall& c, const GemmWorkspace& w, cudaStream_t stream, int* launches = nullptr);

// Bytes of split-K workspace / partial slots a call will need (so plans can size buffers up front).
void gemm_plan(int64_t Mside, int64_t K, int nblk, int num_sms, int force_splits,
               int* tilesM, int* splits, int64_t* chunks_per_split, size_t* ws_bytes, size_t* n_partials);

bool tma_compatible(const MatView& v);

// *slot = partials[0] + ... + partials[n-1] in index order (one CTA).
cudaError_t sum_array_launch(const double* partials, int64_t n, double* slot, cudaStream_t stream);

// Sum `splits` partial blocks laid out like the split-K workspace ([split][rows_pad][Lc], rows_pad a multiple of 128)
// into out (row-major, pitch ld, Mside valid rows); with px != nullptr the sum also runs over the ranks (x_count doubles).
cudaError_t reduce_partials_launch(const double* ws, int splits, int rows_pad, int Lc, int64_t Mside, double* out,
                                   int64_t ld, const PeerExchange* px, size_t x_count, const int* cond_flag,
                                   int num_sms, cudaStream_t stream, int* launches = nullptr);

}  // namespace corrla
