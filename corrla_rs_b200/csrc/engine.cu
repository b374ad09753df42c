// Host-side engine and C ABI of libcorrla_b200.so: orchestrates the RSVD of the reference
// (random_svd.rs:15-110) as a sequence of skinny DMMA GEMMs, CholeskyQR and a Jacobi SVD, all on one
// CUDA stream, with NCCL all-reduces of the small replicated factors when the rows are sharded.
#include <sys/mman.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/corrla_b200.h"
#include "comm.cuh"
#include "hostcopy.cuh"
#include "skinny_gemm.cuh"
#include "small_kernels.cuh"

using namespace corrla;

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
struct corrla_ctx {
  int device = 0;
  int num_sms = 148;
  cudaStream_t own_stream = nullptr;
  cudaStream_t copy_stream = nullptr;   // host->device chunks of A while the main stream computes on earlier chunks
  std::mutex mu;
  struct Buf { void* p = nullptr; size_t bytes = 0; };
  std::map<std::string, Buf> pool;
  void* pinned = nullptr; size_t pinned_bytes = 0;
  BounceBuffers bounce;
  int* hflag = nullptr;              // pinned: device-side decisions read back by the host (two ints)
  std::vector<cudaEvent_t> events;   // reusable timing events
  cudaEvent_t event(size_t i) {
    while (events.size() <= i) {
      cudaEvent_t e = nullptr;
      if (cudaEventCreate(&e) != cudaSuccess) { cudaGetLastError(); return nullptr; }
      events.push_back(e);
    }
    return events[i];
  }

  // grow-only device buffer
  void* get(const char* name, size_t bytes) {
    Buf& b = pool[name];
    if (b.bytes >= bytes && b.p != nullptr) return b.p;
    if (b.p != nullptr) { cudaFree(b.p); b.p = nullptr; b.bytes = 0; }
    void* p = nullptr;
    const size_t want = std::max<size_t>(bytes, 256);
    if (cudaMalloc(&p, want) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    b.p = p; b.bytes = want;
    return p;
  }
  void* get_pinned(size_t bytes) {
    if (pinned_bytes >= bytes) return pinned;
    if (pinned) cudaFreeHost(pinned);
    pinned = nullptr; pinned_bytes = 0;
    if (cudaMallocHost(&pinned, bytes) != cudaSuccess) { cudaGetLastError(); pinned = nullptr; return nullptr; }
    pinned_bytes = bytes;
    return pinned;
  }
  ~corrla_ctx() {
    for (auto& kv : pool) if (kv.second.p) cudaFree(kv.second.p);
    for (auto e : events) cudaEventDestroy(e);
    if (pinned) cudaFreeHost(pinned);
    bounce.release();
    if (hflag) cudaFreeHost(hflag);
    if (own_stream) cudaStreamDestroy(own_stream);
    if (copy_stream) cudaStreamDestroy(copy_stream);
  }
};

namespace {

#define CU_TRY(expr)                                                                               \
  do {                                                                                             \
    cudaError_t e__ = (expr);                                                                      \
    if (e__ != cudaSuccess) {                                                                      \
      set_last_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
      cudaGetLastError();                                                                          \
      return CORRLA_ERR_CUDA;                                                                      \
    }                                                                                              \
  } while (0)

#define ST_TRY(expr)               \
  do {                             \
    int s__ = (expr);              \
    if (s__ != CORRLA_OK) return s__; \
  } while (0)

inline int64_t round_up(int64_t x, int64_t q) { return (x + q - 1) / q * q; }

int ensure_device(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    set_last_error("no CUDA device available (%s); this library has no CPU fallback",
                   e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    return CORRLA_ERR_NO_DEVICE;
  }
  if (device >= 0) {
    if (device >= count) { set_last_error("device %d out of range (%d devices)", device, count); return CORRLA_ERR_INVALID; }
    CU_TRY(cudaSetDevice(device));
  }
  return CORRLA_OK;
}

int ctx_create(int device, corrla_ctx** out) {
  ST_TRY(ensure_device(device));
  corrla_ctx* c = new corrla_ctx();
  if (device < 0) cudaGetDevice(&device);
  c->device = device;
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, device) == cudaSuccess) c->num_sms = p.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaMallocHost(reinterpret_cast<void**>(&c->hflag), 64) != cudaSuccess) {
    set_last_error("cudaStreamCreate / cudaMallocHost failed");
    delete c;
    return CORRLA_ERR_CUDA;
  }
  *out = c;
  return CORRLA_OK;
}

// ---------------------------------------------------------------------------------------------
// one RSVD / power_iter / QR run on a resident thin matrix
// ---------------------------------------------------------------------------------------------
struct Core {
  corrla_ctx* ctx = nullptr;
  cudaStream_t st = nullptr;
  corrla_comm* comm = nullptr;
  GemmWorkspace gw;
  // thin A: m local rows, n columns
  bool a_rowmajor = true;   // true: view(inner = n, outer = m); false: view(inner = m, outer = n)
  MatView av{};
  int64_t m = 0, n = 0, m16 = 0, n16 = 0;
  double grows = 0;         // global row count (for the CholeskyQR shift)
  int l = 0, nblk = 0, Lc = 0, ld = 0, L16 = 0;
  double *Y = nullptr, *Za = nullptr, *Zb = nullptr, *Qz = nullptr;
  double *SK = nullptr;      // sketch of the matrix being orthonormalised: sketch_rows(Lc) x ld
  double *G = nullptr, *T1 = nullptr, *Tf = nullptr, *Tzf = nullptr, *Wm = nullptr, *Vr = nullptr, *Ur = nullptr,
         *M1 = nullptr, *sig = nullptr, *scal = nullptr, *jscratch = nullptr;
  // flags: [0] flag3 of the current QR, [1..2] chol info (live, shifted), [3] dead-column flag of the current QR,
  //        [4..5] jacobi info, [8] third-pass counter, [9] refill counter, [12..15] scratch info of refill-phase chol
  int* flags = nullptr;
  int* deadmask = nullptr;  // l ints
  // on-the-fly centring (thin matrix is C = A - 1*mu^T): mu (n), bvec = X^T mu (Lc), column sums of Y live right
  // behind the Z buffers (so one all-reduce carries both), sum_partials is scratch of the column-sum kernels
  int center = 0;
  double *mu = nullptr, *bvec = nullptr, *sum_partials = nullptr;
  uint64_t refill_seed = 0x5eedu; uint64_t refill_stream = 0; int qr_calls = 0;
  int launches = 0;
  int64_t chunk_rows = 0;   // > 0: the first two passes run per row chunk while A is still arriving from the host

  size_t small_elems() const { return (size_t)L16 * ld; }

  int setup_dims(int64_t m_, int64_t n_, int l_) {
    m = m_; n = n_; l = l_;
    nblk = (l + 7) / 8; Lc = nblk * 8; ld = Lc + 4; L16 = (int)round_up(Lc, 16);
    m16 = round_up(m, 16); n16 = round_up(n, 16);
    if (nblk > kMaxNblk) {
      set_last_error("n_rank + n_oversamples = %d exceeds the 128-column limit of the register-tiled kernels", l);
      return CORRLA_ERR_UNSUPPORTED;
    }
    return CORRLA_OK;
  }

  // workspace big enough for every product this run can issue
  int alloc_workspace(bool need_z) {
    size_t ws = 0, np = 0;
    auto need = [&](int64_t Mside, int64_t K) {
      int t, s; int64_t cps; size_t w, p;
      gemm_plan(Mside, K, nblk, ctx->num_sms, 0, &t, &s, &cps, &w, &p);
      ws = std::max(ws, w); np = std::max(np, p);
      np = std::max(np, (size_t)t * 8);
    };
    need(m, n); need(n, m); need(Lc, m); need(m, Lc);
    if (chunk_rows > 0) { need(chunk_rows, n); need(n, chunk_rows); need(m % chunk_rows ? m % chunk_rows : chunk_rows, n); need(n, m % chunk_rows ? m % chunk_rows : chunk_rows); }
    ws = std::max(ws, sketch_ws_bytes(Lc, ctx->num_sms));
    if (need_z) { need(Lc, n); need(n, Lc); need(Lc, Lc); }
    gw.num_sms = ctx->num_sms;
    gw.ws_bytes = ws;
    gw.ws = ws ? static_cast<double*>(ctx->get("ws", ws)) : nullptr;
    gw.n_partials = np + 8;
    gw.sumsq_partials = static_cast<double*>(ctx->get("partials", gw.n_partials * 8));
    if ((ws && !gw.ws) || !gw.sumsq_partials) { set_last_error("device allocation of the split-K workspace failed"); return CORRLA_ERR_ALLOC; }
    return CORRLA_OK;
  }

  int alloc_buffers(bool need_z) {
    auto getz = [&](const char* name, size_t elems) -> double* {
      double* p = static_cast<double*>(ctx->get(name, elems * 8));
      if (p != nullptr && cudaMemsetAsync(p, 0, elems * 8, st) != cudaSuccess) { cudaGetLastError(); return nullptr; }
      return p;
    };
    Y = getz("Y", (size_t)m16 * ld);
    G = getz("G", small_elems()); T1 = getz("T1", small_elems()); Tf = getz("Tf", small_elems());
    SK = getz("SK", (size_t)256 * ld + 128);
    scal = getz("scal", 16);
    flags = reinterpret_cast<int*>(getz("flags", 16));
    deadmask = reinterpret_cast<int*>(getz("deadmask", (size_t)L16));
    bool ok = Y && G && T1 && Tf && scal && flags && deadmask && SK;
    if (need_z) {
      Za = getz("Za", (size_t)n16 * ld + 256); Zb = getz("Zb", (size_t)n16 * ld + 256); Qz = getz("Qz", (size_t)n16 * ld);
      if (center) {
        mu = getz("mu", (size_t)n16 + 128); bvec = getz("bvec", 256);
        sum_partials = getz("sum_partials", (size_t)sum_blocks(m) * (size_t)std::max<int64_t>(n, Lc) + 128);
        if (!mu || !bvec || !sum_partials) { set_last_error("device allocation failed (centring buffers)"); return CORRLA_ERR_ALLOC; }
      }
      Tzf = getz("Tzf", small_elems()); Wm = getz("Wm", small_elems()); Vr = getz("Vr", small_elems());
      Ur = getz("Ur", small_elems()); M1 = getz("M1", small_elems()); sig = getz("sig", (size_t)L16);
      jscratch = getz("jscratch", 2 * (size_t)l * (l + 2) + 8);
      ok = ok && Za && Zb && Qz && Tzf && Wm && Vr && Ur && M1 && sig && jscratch;
    }
    if (!ok) { set_last_error("device allocation failed (m=%lld n=%lld l=%d)", (long long)m, (long long)n, l); return CORRLA_ERR_ALLOC; }
    return CORRLA_OK;
  }

  bool profile_passes = false;
  int p2p_exchanges = 0;
  int n_pass_events = 0;    // pairs recorded so far: events 2i, 2i+1 (offset by 2 for the whole-call pair)

  int mm(const MatView& a, bool reduce_inner, const double* B, double* out, int64_t ors, int64_t ocs, int ncols_out,
         const double* alpha = nullptr, double* sumsq = nullptr, const int* cond = nullptr, int force_splits = 0,
         bool is_pass = false, const double* col_bias = nullptr, size_t x_count = 0, size_t x_extra = 0,
         bool accumulate = false, int mode = 0) {
    GemmCall c{};
    c.col_bias = col_bias;
    c.accumulate = accumulate;
    c.mode = mode;
    // cross-rank sum of the product: fused into the reduction kernel over peer memory when possible, else NCCL
    PeerExchange px;
    bool nccl_after = false;
    if (x_count > 0 && comm != nullptr && comm->nranks > 1) {
      if (comm->next_exchange(x_count, &px)) { c.px = &px; c.x_count = x_count; c.x_extra = x_extra; ++p2p_exchanges; }
      else nccl_after = true;
    }
    if (is_pass && profile_passes) {
      c.ev_begin = ctx->event(2 + 2 * (size_t)n_pass_events);
      c.ev_end = ctx->event(3 + 2 * (size_t)n_pass_events);
      if (c.ev_begin && c.ev_end) ++n_pass_events;
    }
    c.a = a; c.reduce_inner = reduce_inner; c.B = B; c.ldb = ld; c.nblk = nblk;
    c.out = out; c.out_rs = ors; c.out_cs = ocs; c.ncols_out = ncols_out;
    c.alpha_sumsq = alpha; c.sumsq_slot = sumsq; c.cond_flag = cond; c.force_splits = force_splits;
    cudaError_t e = gemm_launch(c, gw, st, &launches);
    if (e != cudaSuccess) {
      set_last_error("skinny GEMM launch failed: %s (Mside/K view inner=%lld outer=%lld ld=%lld reduce_inner=%d)",
                     cudaGetErrorString(e), (long long)a.inner, (long long)a.outer, (long long)a.ld, (int)reduce_inner);
      cudaGetLastError();
      return CORRLA_ERR_CUDA;
    }
    if (nccl_after) ST_TRY(allreduce(out, x_count));
    return CORRLA_OK;
  }

  MatView view_rows(const double* p, int64_t rows) const { return MatView{p, (int64_t)Lc, rows, (int64_t)ld}; }

  // G = X^T X (upper triangle only: mode 1), summed over the ranks when gx > 0
  int gram(const MatView& vx, const double* X, const int* cond, size_t gx) {
    return mm(vx, false, X, G, ld, 1, Lc, nullptr, nullptr, cond, 0, false, nullptr, gx, 0, false, 1);
  }
  // X <- X * T in place, T upper triangular (mode 2)
  int apply_tri(const MatView& vx, const double* T, double* X, const int* cond) {
    return mm(vx, true, T, X, ld, 1, Lc, nullptr, nullptr, cond, 1, false, nullptr, 0, 0, false, 2);
  }

  // Y[m x Lc] = alpha * C * X, C = A or A - 1*mu^T        (X: n16 x ld)
  int mm_AX(const double* X, double* Yout, const double* alpha, double* sumsq) {
    const double* bias = nullptr;
    if (center) {
      cudaError_t e = gemv_t_launch(X, n, Lc, ld, mu, bvec, st);     // b = X^T mu: (1 mu^T) X = 1 b^T
      ++launches;
      if (e != cudaSuccess) { set_last_error("gemv launch failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return CORRLA_ERR_CUDA; }
      bias = bvec;
    }
    return mm(av, a_rowmajor, X, Yout, ld, 1, Lc, alpha, sumsq, nullptr, 0, true, bias);
  }
  // Z[n x Lc] = C^T * Yin, summed over the ranks    (Yin: m16 x ld).  With centring: A^T Y - mu * (1^T Y).
  int mm_AtY(const double* Yin, double* Zout) {
    // tail of the Z buffer, summed over the ranks together with Z: [column sums of Y (Lc) | ||Y||_F^2 (1)]
    double* colsum = Zout + (size_t)n16 * ld;
    if (center) {
      cudaError_t e = sum_over_outer_launch(Yin, Lc, m, ld, sum_partials, colsum, st);
      launches += 2;
      if (e != cudaSuccess) { set_last_error("column-sum launch failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return CORRLA_ERR_CUDA; }
    }
    ST_TRY(mm(av, !a_rowmajor, Yin, Zout, ld, 1, Lc, nullptr, nullptr, nullptr, 0, true, nullptr,
              (size_t)n16 * ld + (size_t)Lc + 1, (size_t)Lc + 1));
    if (center) {
      cudaError_t e = rank1_sub_launch(Zout, n, Lc, ld, mu, colsum, st);
      ++launches;
      if (e != cudaSuccess) { set_last_error("rank-1 launch failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return CORRLA_ERR_CUDA; }
    }
    return CORRLA_OK;
  }

  // mu = column means of the thin matrix (all ranks), for the fused centring
  int compute_means_thin_cols() {
    cudaError_t e = a_rowmajor ? sum_over_outer_launch(av.p, n, m, av.ld, sum_partials, mu, st)
                               : sum_over_inner_launch(av.p, m, n, av.ld, mu, st);
    launches += 2;
    if (e != cudaSuccess) { set_last_error("mean launch failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return CORRLA_ERR_CUDA; }
    ST_TRY(allreduce(mu, (size_t)n));
    e = scale_vec_launch(mu, n, 1.0 / grows, st);
    ++launches;
    if (e != cudaSuccess) { set_last_error("scale launch failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return CORRLA_ERR_CUDA; }
    return CORRLA_OK;
  }

  int allreduce(double* buf, size_t count) {
    if (comm == nullptr || comm->nranks <= 1) return CORRLA_OK;
    return comm->allreduce_f64(buf, count, st);
  }

  // Cholesky + inverse of G -> T.  quiet: do not publish liveness (deadmask / flags[1..3]); f2: where kCholCheck
  // writes its "one more pass" flag.
  int chol(int mode, double rows_for_shift, double* T, const int* cond, bool quiet = false, int* f2 = nullptr) {
    cudaError_t e = quiet
        ? chol_inv_launch(G, ld, l, T, L16, ld, mode, rows_for_shift, f2, flags + 12, scal + 9, nullptr, flags + 14, cond, st)
        : chol_inv_launch(G, ld, l, T, L16, ld, mode, rows_for_shift, f2 ? f2 : flags + 0, flags + 1, scal + 8, deadmask, flags + 3, cond, st);
    ++launches;
    if (e != cudaSuccess) { set_last_error("chol_inv launch failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return CORRLA_ERR_CUDA; }
    return CORRLA_OK;
  }

  // One sketch-preconditioned CholeskyQR stage on X (rows x Lc, pitch ld), in place; every kernel of the stage is
  // skipped on the device when cond != nullptr and *cond == 0.
  //   SK = S*X (sparse sign sketch)  ->  Householder QR of SK in one CTA  ->  T1 = R^-1 (deflated)  ->  X <- X*T1
  //   (cond ~ 5 whatever cond(X) was, up to ~1e15)  ->  Gram  ->  Cholesky  ->  Tfold;  Q = X*Tfold is never formed.
  // A second CholeskyQR pass runs only if the pivot ratio says the embedding was unlucky (device flag f2).
  int qr_stage(double* X, int64_t rows, bool distributed, double rows_for_shift, double* Tfold, const int* cond,
               int* f2, bool refill_phase) {
    const MatView vx = view_rows(X, rows);
    const size_t gx = distributed ? (size_t)Lc * ld : 0;   // Gram matrices are summed over the ranks in the reduction kernel
    const int s_rows = sketch_rows(Lc);
    const int s_pad = (s_rows + 127) / 128 * 128;
    int grid = 0;
    cudaError_t e = sketch_launch(X, rows, Lc, ld, refill_seed + 0x5ce7c4ull * (uint64_t)(2 * qr_calls + (refill_phase ? 2 : 1)),
                                  refill_stream, gw.ws, gw.num_sms, &grid, cond, st);
    ++launches;
    if (e != cudaSuccess) { set_last_error("sketch launch failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return CORRLA_ERR_CUDA; }
    const size_t xs = (size_t)s_rows * ld;
    PeerExchange px;
    const bool multi = distributed && comm != nullptr && comm->nranks > 1;
    const bool fused = multi && comm->next_exchange(xs, &px);
    if (fused) ++p2p_exchanges;
    e = reduce_partials_launch(gw.ws, grid, s_pad, Lc, s_rows, SK, ld, fused ? &px : nullptr, xs, cond, gw.num_sms, st, &launches);
    if (e != cudaSuccess) { set_last_error("sketch reduction failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return CORRLA_ERR_CUDA; }
    if (multi && !fused) ST_TRY(allreduce(SK, xs));
    // main phase publishes liveness (flags[1..3], deadmask); the refill phase must not clobber it
    e = refill_phase ? hqr_inv_launch(SK, ld, s_rows, l, T1, L16, ld, flags + 12, nullptr, flags + 14, cond, st)
                     : hqr_inv_launch(SK, ld, s_rows, l, T1, L16, ld, flags + 1, deadmask, flags + 3, cond, st);
    ++launches;
    if (e != cudaSuccess) { set_last_error("hqr_inv launch failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return CORRLA_ERR_CUDA; }
    ST_TRY(apply_tri(vx, T1, X, cond));
    ST_TRY(gram(vx, X, cond, gx));
    ST_TRY(chol(kCholCheck, rows_for_shift, Tfold, cond, true, f2));
    ST_TRY(apply_tri(vx, Tfold, X, f2));
    ST_TRY(gram(vx, X, f2, gx));
    ST_TRY(chol(kCholPlain, rows_for_shift, Tfold, f2, true, nullptr));
    return CORRLA_OK;
  }

  // Thin-Q factor of X in place: on return the orthonormal factor is X * Tfold (never formed here).
  //   1. Gram + Cholesky probe.  If every pivot ratio is >= 1e-8 (cond(X) below ~1e4) plain CholeskyQR2 finishes:
  //      apply, Gram, Cholesky -- the cheapest path, taken by the benchmark matrices.
  //   2. Otherwise (device flag, no host round trip) the sketch-preconditioned stage runs instead: Householder-grade
  //      stability up to cond ~ 1e15, rank decisions made column-relative on the sketch.
  //   3. Columns found numerically dependent are refilled with fresh vectors and orthonormalised again.
  // distributed: rows are sharded over comm.  refill_from_a: X is A times something, so directions lost to numerical
  // rank deficiency are replaced by fresh vectors from range(A) instead of arbitrary ones.
  // Read two device ints back (pinned buffer + stream sync).  The three decisions per QR that the host takes this way
  // cost ~10 us each; enqueueing every alternative behind device-side flags cost more in empty launches.
  int read_flags(const int* dev, int* out0, int* out1) {
    CU_TRY(cudaMemcpyAsync(ctx->hflag, dev, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    *out0 = ctx->hflag[0];
    if (out1) *out1 = ctx->hflag[1];
    return CORRLA_OK;
  }

  int qr_inplace(double* X, int64_t rows, bool distributed, double rows_for_shift, double* Tfold,
                 bool refill_from_a = false) {
    const MatView vx = view_rows(X, rows);
    const size_t gx = distributed ? (size_t)Lc * ld : 0;
    CU_TRY(cudaMemsetAsync(flags, 0, 8 * sizeof(int), st));      // f2 (0), liveness (1..3), refill-stage f2 (6) start clear
    int* fs = flags + 16;                                        // [0] robust stage needed, [1] fast path ok
    ST_TRY(gram(vx, X, nullptr, gx));
    ST_TRY(chol(kCholProbe, rows_for_shift, T1, nullptr, true, fs));
    int robust = 0;
    ST_TRY(read_flags(fs, &robust, nullptr));                    // identical on every rank: G is the all-reduced Gram
    if (!robust) {
      // fast path: CholeskyQR2
      ST_TRY(apply_tri(vx, T1, X, nullptr));
      ST_TRY(gram(vx, X, nullptr, gx));
      ST_TRY(chol(kCholPlain, rows_for_shift, Tfold, nullptr, true, nullptr));
      ++qr_calls;
      return CORRLA_OK;
    }
    ++n_robust;
    ST_TRY(qr_stage(X, rows, distributed, rows_for_shift, Tfold, nullptr, flags + 0, false));
    // Refill, only when columns were deflated as numerically dependent: form Q (zero columns where dead), put fresh
    // vectors into those columns and orthonormalise again -- the completion a Householder QR would return
    // (random_svd.rs:38 keeps l orthonormal columns even for rank-deficient Y).
    int any_dead = 0;
    ST_TRY(read_flags(flags + 3, &any_dead, nullptr));
    if (any_dead) {
      ++n_refill;
      ST_TRY(apply_tri(vx, Tfold, X, nullptr));
      if (refill_from_a && Za != nullptr) {
        // X[:, dead] += A * Omega', Omega' Gaussian in the dead columns and zero elsewhere (Za is free while Y is being
        // orthonormalised).  Householder's completion of a numerically rank-deficient Y is rounding noise of A*(...),
        // which lies in range(A) too; vectors from outside it would waste the slots.
        CU_TRY(cudaMemsetAsync(Za, 0, (size_t)n16 * ld * 8, st));
        cudaError_t e = refill_dead_launch(Za, n, l, ld, deadmask, refill_seed + (uint64_t)qr_calls, 0, nullptr, st);
        ++launches;
        if (e != cudaSuccess) { set_last_error("refill launch failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return CORRLA_ERR_CUDA; }
        ST_TRY(mm(av, a_rowmajor, Za, X, ld, 1, Lc, nullptr, nullptr, nullptr, 0, false, nullptr, 0, 0, true));
      } else {
        cudaError_t e = refill_dead_launch(X, rows, l, ld, deadmask, refill_seed + (uint64_t)qr_calls, refill_stream, nullptr, st);
        ++launches;
        if (e != cudaSuccess) { set_last_error("refill launch failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return CORRLA_ERR_CUDA; }
      }
      ST_TRY(qr_stage(X, rows, distributed, rows_for_shift, Tfold, nullptr, flags + 6, true));
    }
    ++qr_calls;
    return CORRLA_OK;
  }

  int n_robust = 0, n_refill = 0;


  // Omega (n x l, standard normal, Philox counter = element index) into Za     random_svd.rs:27
  int draw_omega(uint64_t seed) {
    cudaError_t e = philox_normal_launch(Za, n, l, ld, seed, st);
    ++launches;
    if (e != cudaSuccess) { set_last_error("philox launch failed: %s", cudaGetErrorString(e)); return CORRLA_ERR_CUDA; }
    return CORRLA_OK;
  }

  // Host-resident A: copy it in row chunks on the copy stream and run the first product Y = A*Omega -- and, when
  // `with_second` is set, the first Z = A^T*Y as a running sum -- chunk by chunk behind the copies, so that the two
  // passes cost no time on top of the transfer.  Leaves `av` describing the full resident matrix.
  // `a` is the thin matrix on the host with strides (rs, cs); rowmajor_like says which stride is 1.
  int stream_in(const double* a, int64_t rs, int64_t cs, bool rowmajor_like, bool with_second, int* n_chunks) {
    const int64_t inner = rowmajor_like ? n : m, outer = rowmajor_like ? m : n;
    const int64_t src_ld = rowmajor_like ? rs : cs;
    const int64_t ldd = round_up(inner, 2);
    double* buf = static_cast<double*>(ctx->get("A", (size_t)outer * ldd * 8));
    if (!buf) { set_last_error("device allocation for A failed (%lld x %lld)", (long long)m, (long long)n); return CORRLA_ERR_ALLOC; }
    av = MatView{buf, inner, outer, ldd};
    a_rowmajor = rowmajor_like;
    const int nch = (int)((m + chunk_rows - 1) / chunk_rows);
    double* slots = static_cast<double*>(ctx->get("chunk_nu2", (size_t)nch * 8));
    cudaEvent_t ev = ctx->event(0);           // re-recorded per chunk: a stream wait binds to the record before it
    if (!slots || !ev) { set_last_error("allocation failed (streamed input)"); return CORRLA_ERR_ALLOC; }
    cudaStream_t cst = ctx->copy_stream;
    // earlier work on the compute stream may still read the A buffer (context reuse without a sync in between)
    CU_TRY(cudaEventRecord(ev, st));
    CU_TRY(cudaStreamWaitEvent(cst, ev, 0));
    double* nu2 = Zb + (size_t)n16 * ld + Lc;
    int status = CORRLA_OK;
    for (int ci = 0; ci < nch && status == CORRLA_OK; ++ci) {
      const int64_t r0 = (int64_t)ci * chunk_rows, rc = std::min<int64_t>(chunk_rows, m - r0);
      cudaError_t e;
      MatView cv;
      if (rowmajor_like) {
        e = copy_h2d_2d(ctx->bounce, cst, buf + r0 * ldd, ldd * 8, a + r0 * src_ld, src_ld * 8, n * 8, rc, false);
        cv = MatView{buf + r0 * ldd, n, rc, ldd};
      } else {
        e = copy_h2d_2d(ctx->bounce, cst, buf + r0, ldd * 8, a + r0, src_ld * 8, rc * 8, n, false);
        cv = MatView{buf + r0, rc, n, ldd};
      }
      if (e == cudaSuccess) e = cudaEventRecord(ev, cst);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(st, ev, 0);
      if (e != cudaSuccess) { set_last_error("host->device copy of A failed: %s", cudaGetErrorString(e)); cudaGetLastError(); status = CORRLA_ERR_CUDA; break; }
      double* Yc = Y + (size_t)r0 * ld;
      status = mm(cv, a_rowmajor, Za, Yc, ld, 1, Lc, nullptr, slots + ci);
      if (status == CORRLA_OK && with_second)
        status = mm(cv, !a_rowmajor, Yc, Zb, ld, 1, Lc, nullptr, nullptr, nullptr, 0, false, nullptr, 0, 0, ci > 0);
    }
    // the caller may free or overwrite its matrix once we return: wait for the last DMA (the products stay queued)
    cudaError_t e = cudaStreamSynchronize(cst);
    ctx->bounce.in_flight[0] = ctx->bounce.in_flight[1] = false;
    if (status != CORRLA_OK) return status;
    if (e != cudaSuccess) { set_last_error("host->device copy of A failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return CORRLA_ERR_CUDA; }
    e = sum_array_launch(slots, nch, nu2, st);
    ++launches;
    if (e != cudaSuccess) { set_last_error("norm reduction launch failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return CORRLA_ERR_CUDA; }
    *n_chunks = nch;
    return CORRLA_OK;
  }

  // power iteration with the reference schedule; leaves Y and Tf such that Q = Y * Tf
  // resume: 0 = from scratch; 1 = Omega, Y = A*Omega and its norm are already there (streamed host path);
  //         2 = additionally Zb = A^T*Y of the first iteration is there, summed over the ranks
  int power_iter(const double* omega_dev_packed, uint64_t seed, int n_iter, int schedule, int resume = 0) {
    if (omega_dev_packed == nullptr && resume == 0) ST_TRY(draw_omega(seed));
    // ||Y||_F^2 lives in the tail of Zb: the next A^T*Y sums it over the ranks together with Z
    double* nu2 = Zb + (size_t)n16 * ld + Lc;
    if (resume == 0) ST_TRY(mm_AX(Za, Y, nullptr, nu2));      // random_svd.rs:31
    for (int i = 0; i < n_iter; ++i) {                        // :35
      const bool do_qr = (schedule == 1) || (i > 2);          // :37
      if (do_qr) {
        ST_TRY(qr_inplace(Y, m, true, grows, Tf, true));      // :38
        ST_TRY(mm_AtY(Y, Zb));                                // :42-46 (on the pre-fold iterate), all-reduced
        ST_TRY(mm(view_rows(Zb, n), true, Tf, Za, ld, 1, Lc)); // fold R^-1 into the small side
        ST_TRY(mm_AX(Za, Y, nullptr, nu2));                   // :47-51
      } else {
        if (!(i == 0 && resume == 2)) ST_TRY(mm_AtY(Y, Zb));
        ST_TRY(mm_AX(Zb, Y, nu2, nu2));                       // :47-51 with the deferred :53-55 scaling
      }
    }
    ST_TRY(qr_inplace(Y, m, true, grows, Tf, true));          // :57
    return CORRLA_OK;
  }
};


struct Timer {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  double ms() const { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};

// Describe a strided matrix as a TMA-friendly view if possible.
// rows x cols with element strides (rs, cs).  Returns 1 = row-major-like, 2 = column-major-like, 0 = needs repack.
int classify(const double* p, int64_t rows, int64_t cols, int64_t rs, int64_t cs, MatView* v) {
  if (cs == 1 && rs >= cols) {
    *v = MatView{p, cols, rows, rs};
    if (tma_compatible(*v)) return 1;
  }
  if (rs == 1 && cs >= rows) {
    *v = MatView{p, rows, cols, cs};
    if (tma_compatible(*v)) return 2;
  }
  return 0;
}

// Bring a (possibly host, possibly oddly strided) matrix into device memory as a TMA-compatible view.
// *rowmajor tells which of the two contractions is "reduce_inner".
int stage_matrix(corrla_ctx* ctx, cudaStream_t st, const char* bufname, const double* a, int64_t rows, int64_t cols,
                 int64_t rs, int64_t cs, bool on_device, MatView* view, bool* rowmajor, double* h2d_ms, int* launches) {
  if (on_device) {
    const int kind = classify(a, rows, cols, rs, cs, view);
    if (kind != 0) { *rowmajor = (kind == 1); return CORRLA_OK; }
    // unaligned or doubly-strided device view: one repack pass into an aligned copy, keeping its orientation
    const bool colmajor_like = (llabs(rs) < llabs(cs));
    const int64_t outer = colmajor_like ? cols : rows, inner = colmajor_like ? rows : cols;
    const int64_t ldd = round_up(inner, 2);
    double* buf = static_cast<double*>(ctx->get(bufname, (size_t)outer * ldd * 8));
    if (!buf) { set_last_error("device allocation for the repacked matrix failed"); return CORRLA_ERR_ALLOC; }
    cudaError_t e = colmajor_like ? repack_launch(a, cols, rows, cs, rs, buf, ldd, st)
                                  : repack_launch(a, rows, cols, rs, cs, buf, ldd, st);
    if (launches) ++*launches;
    if (e != cudaSuccess) { set_last_error("repack failed: %s", cudaGetErrorString(e)); return CORRLA_ERR_CUDA; }
    *view = MatView{buf, inner, outer, ldd};
    *rowmajor = !colmajor_like;
    return CORRLA_OK;
  }
  // host source
  Timer t;
  const bool rowmajor_like = (cs == 1 && rs >= cols);
  const bool colmajor_like = !rowmajor_like && (rs == 1 && cs >= rows);
  if (rowmajor_like || colmajor_like) {
    const int64_t outer = rowmajor_like ? rows : cols, inner = rowmajor_like ? cols : rows;
    const int64_t src_ld = rowmajor_like ? rs : cs;
    const int64_t ldd = round_up(inner, 2);
    double* buf = static_cast<double*>(ctx->get(bufname, (size_t)outer * ldd * 8));
    if (!buf) { set_last_error("device allocation for A failed (%lld x %lld)", (long long)rows, (long long)cols); return CORRLA_ERR_ALLOC; }
    CU_TRY(copy_h2d_2d(ctx->bounce, st, buf, ldd * 8, a, src_ld * 8, inner * 8, outer));
    *view = MatView{buf, inner, outer, ldd};
    *rowmajor = rowmajor_like;
  } else {
    // arbitrary host strides: pack on the host (rare: sliced numpy views)
    const int64_t ldd = round_up(cols, 2);
    std::vector<double> tmp;
    try { tmp.assign((size_t)rows * ldd, 0.0); } catch (...) { set_last_error("host allocation failed"); return CORRLA_ERR_ALLOC; }
    for (int64_t i = 0; i < rows; ++i)
      for (int64_t j = 0; j < cols; ++j) tmp[i * ldd + j] = a[i * rs + j * cs];
    double* buf = static_cast<double*>(ctx->get(bufname, (size_t)rows * ldd * 8));
    if (!buf) { set_last_error("device allocation for A failed"); return CORRLA_ERR_ALLOC; }
    CU_TRY(cudaMemcpyAsync(buf, tmp.data(), tmp.size() * 8, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaStreamSynchronize(st));
    *view = MatView{buf, cols, rows, ldd};
    *rowmajor = true;
  }
  if (h2d_ms) *h2d_ms += t.ms();
  return CORRLA_OK;
}

// Pack a small strided matrix (rows x cols) into the engine's padded row-major layout dst (pitch ld).
int pack_small(corrla_ctx* ctx, cudaStream_t st, const double* src, int64_t rows, int64_t cols, int64_t rs, int64_t cs,
               bool on_device, double* dst, int64_t ld, double scale, int* launches) {
  if (on_device) {
    cudaError_t e = repack_launch(src, rows, cols, rs, cs, dst, ld, st, scale);
    if (launches) ++*launches;
    if (e != cudaSuccess) { set_last_error("repack failed: %s", cudaGetErrorString(e)); return CORRLA_ERR_CUDA; }
    return CORRLA_OK;
  }
  const size_t bytes = (size_t)rows * cols * 8;
  double* stage = static_cast<double*>(ctx->get_pinned(bytes));
  if (!stage) { set_last_error("pinned allocation failed"); return CORRLA_ERR_ALLOC; }
  for (int64_t i = 0; i < rows; ++i)
    for (int64_t j = 0; j < cols; ++j) stage[i * cols + j] = scale * src[i * rs + j * cs];
  CU_TRY(cudaMemcpy2DAsync(dst, ld * 8, stage, cols * 8, cols * 8, rows, cudaMemcpyHostToDevice, st));
  CU_TRY(cudaStreamSynchronize(st));   // the pinned stage is reused
  return CORRLA_OK;
}

struct Scope {
  corrla_ctx* ctx = nullptr; bool owned = false; cudaStream_t st = nullptr;
  std::unique_lock<std::mutex> lock;
  ~Scope() { if (lock.owns_lock()) lock.unlock(); if (owned) delete ctx; }
};

int open_scope(const corrla_rsvd_opts* o, Scope* s) {
  const int device = o ? o->device : -1;
  ST_TRY(ensure_device(device));
  if (o && o->ctx) { s->ctx = o->ctx; s->owned = false; CU_TRY(cudaSetDevice(s->ctx->device)); }
  else { ST_TRY(ctx_create(device, &s->ctx)); s->owned = true; }
  s->lock = std::unique_lock<std::mutex>(s->ctx->mu);
  s->st = (o && o->stream) ? static_cast<cudaStream_t>(o->stream) : s->ctx->own_stream;
  return CORRLA_OK;
}

corrla_rsvd_opts default_opts() {
  corrla_rsvd_opts o;
  memset(&o, 0, sizeof(o));
  o.device = -1;
  return o;
}

int rsvd_impl(const double* a, int64_t nrows, int64_t ncols, int64_t rs, int64_t cs, size_t n_rank, size_t n_iter,
              size_t n_oversamples, const corrla_rsvd_opts* opts_in, double* u, double* s, double* vt,
              corrla_timings* tm, bool power_only, double* q_out, bool u_optional = false, double* means_out = nullptr) {
  Timer total;
  corrla_rsvd_opts o = opts_in ? *opts_in : default_opts();
  if (tm) memset(tm, 0, sizeof(*tm));
  if (a == nullptr || nrows <= 0 || ncols <= 0) { set_last_error("empty or null input matrix"); return CORRLA_ERR_INVALID; }
  if (power_only) { if (q_out == nullptr) { set_last_error("null output"); return CORRLA_ERR_INVALID; } }
  else if ((u == nullptr && !u_optional) || s == nullptr || vt == nullptr) { set_last_error("null output"); return CORRLA_ERR_INVALID; }

  // thin orientation (random_svd.rs:69-74).  With a communicator the caller passes thin row blocks already.
  bool fat = false;
  int64_t m = nrows, n = ncols, trs = rs, tcs = cs;
  if (!power_only && o.comm == nullptr && nrows < ncols) { fat = true; m = ncols; n = nrows; trs = cs; tcs = rs; }
  int l;
  size_t k = n_rank;
  if (power_only) {
    l = (int)std::min<size_t>(n_rank, (size_t)1 << 20);   // omega_rank is used as given (random_svd.rs:24)
  } else {
    const size_t want = n_rank + n_oversamples;
    l = (int)std::min<size_t>(want, (size_t)n);              // :77
    if (k > (size_t)l) {
      set_last_error("n_rank=%zu exceeds l=min(n_rank+n_oversamples, ncols)=%d (the reference panics here)", n_rank, l);
      return CORRLA_ERR_RANK;
    }
  }
  if (l <= 0 || (!power_only && k == 0)) { set_last_error("rank must be positive"); return CORRLA_ERR_INVALID; }

  Scope sc;
  ST_TRY(open_scope(&o, &sc));
  Core c;
  c.ctx = sc.ctx; c.st = sc.st; c.comm = o.comm;
  c.refill_seed = o.seed ^ 0x9e3779b97f4a7c15ull;
  c.refill_stream = o.comm ? (uint64_t)o.comm->rank : 0;
  ST_TRY(c.setup_dims(m, n, l));

  const bool want_center = (o.center != 0) && !power_only;
  if (want_center && fat && o.comm != nullptr) { set_last_error("centring of a fat matrix is not supported with a communicator"); return CORRLA_ERR_UNSUPPORTED; }
  c.center = (want_center && !fat) ? 1 : 0;       // tall: rank-1 corrections inside the passes

  // Large host-resident inputs are streamed: the first product (and the first transposed product when the schedule
  // starts without a QR and no other rank is involved) runs chunk by chunk behind the host->device copies.
  const bool host_rowmajor = (tcs == 1 && trs >= n), host_colmajor = !host_rowmajor && (trs == 1 && tcs >= m);
  int64_t stream_rows = 0;
  if (!o.a_on_device && !want_center && (host_rowmajor || host_colmajor)) {
    int64_t rows = ((int64_t)512 << 20) / (n * 8) / 128 * 128;
    if (const char* env = getenv("CORRLA_B200_STREAM_ROWS")) rows = atoll(env) / 128 * 128;   // 0 disables
    if (rows >= 128 && m > rows) stream_rows = rows;
  }
  c.chunk_rows = stream_rows;

  double h2d_ms = 0.0;
  if (stream_rows == 0)
    ST_TRY(stage_matrix(sc.ctx, sc.st, "A", a, m, n, trs, tcs, o.a_on_device != 0, &c.av, &c.a_rowmajor, &h2d_ms, &c.launches));
  ST_TRY(c.alloc_workspace(true));
  ST_TRY(c.alloc_buffers(true));

  c.grows = (double)m;
  if (o.comm != nullptr && o.comm->nranks > 1) {
    if (o.global_rows > 0) c.grows = (double)o.global_rows;
    else {
      double hm = (double)m;
      CU_TRY(cudaMemcpyAsync(c.scal + 4, &hm, 8, cudaMemcpyHostToDevice, sc.st));
      ST_TRY(c.allreduce(c.scal + 4, 1));
      CU_TRY(cudaMemcpyAsync(&hm, c.scal + 4, 8, cudaMemcpyDeviceToHost, sc.st));
      CU_TRY(cudaStreamSynchronize(sc.st));
      c.grows = hm;
    }
  }

  if (c.center) ST_TRY(c.compute_means_thin_cols());
  if (want_center && fat) {
    // thin = a^T: the column means of `a` are per-ROW constants of the thin matrix.  Fat matrices are small on their
    // long side only (n <= m_thin rows of means): take the explicit centred copy the reference takes (center_mat_col).
    double* mu_rows = static_cast<double*>(sc.ctx->get("mu", ((size_t)m + 128) * 8));
    double* part = static_cast<double*>(sc.ctx->get("sum_partials", ((size_t)sum_blocks(n) * (size_t)m + 128) * 8));
    double* cent = static_cast<double*>(sc.ctx->get("Acent", (size_t)c.av.outer * c.av.ld * 8));
    if (!mu_rows || !part || !cent) { set_last_error("device allocation failed (centred copy)"); return CORRLA_ERR_ALLOC; }
    // row sums of the thin matrix: a_rowmajor view = (inner n, outer m) -> sum over inner; else (inner m, outer n) -> over outer
    cudaError_t e = c.a_rowmajor ? sum_over_inner_launch(c.av.p, n, m, c.av.ld, mu_rows, sc.st)
                                 : sum_over_outer_launch(c.av.p, m, n, c.av.ld, part, mu_rows, sc.st);
    if (e == cudaSuccess) e = scale_vec_launch(mu_rows, m, 1.0 / (double)n, sc.st);
    if (e == cudaSuccess) e = center_copy_launch(c.av.p, c.av.inner, c.av.outer, c.av.ld, cent, c.av.ld, mu_rows,
                                                 c.a_rowmajor ? 0 : 1, sc.st);
    c.launches += 4;
    if (e != cudaSuccess) { set_last_error("centring launch failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return CORRLA_ERR_CUDA; }
    c.av.p = cent;
    c.mu = mu_rows;
  }
  if (means_out != nullptr && want_center) {
    const int64_t cnt = fat ? m : n;                 // == ncols of the matrix as passed
    CU_TRY(cudaMemcpyAsync(means_out, c.mu, (size_t)cnt * 8, o.out_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, sc.st));
  }

  const double* omega_packed = nullptr;
  if (o.omega != nullptr) {
    ST_TRY(pack_small(sc.ctx, sc.st, o.omega, n, l, o.omega_rs, o.omega_cs, o.omega_on_device != 0, c.Za, c.ld, 1.0, &c.launches));
    omega_packed = c.Za;
  }
  int resume = 0, n_chunks = 0;
  if (stream_rows > 0) {
    Timer t;
    if (omega_packed == nullptr) ST_TRY(c.draw_omega(o.seed));
    const bool single = (o.comm == nullptr || o.comm->nranks <= 1);
    const bool with_second = single && n_iter >= 1 && o.schedule == 0;     // ranks could disagree on streaming: keep collectives out of it
    ST_TRY(c.stream_in(a, trs, tcs, host_rowmajor, with_second, &n_chunks));
    resume = with_second ? 2 : 1;
    h2d_ms = t.ms();
  }
  if (tm) tm->h2d_ms = h2d_ms;

  // host outputs: fault their pages in while the device computes (after the input copy, which wants the host cores)
  PreTouch pretouch;
  if (o.out_on_device == 0) {
    if (power_only) pretouch.add(q_out, (size_t)m * l * 8);
    else { pretouch.add(u, (size_t)nrows * k * 8); pretouch.add(vt, (size_t)ncols * k * 8); }
  }

  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (tm) {
    ev0 = sc.ctx->event(0); ev1 = sc.ctx->event(1);
    if (!ev0 || !ev1) { set_last_error("cudaEventCreate failed"); return CORRLA_ERR_CUDA; }
    c.profile_passes = true;
    CU_TRY(cudaEventRecord(ev0, sc.st));
  }
  ST_TRY(c.power_iter(omega_packed, o.seed, (int)n_iter, o.schedule, resume));

  const bool out_dev = o.out_on_device != 0;
  double d2h_ms = 0.0;
  if (power_only) {
    // Q = Y * Tf, column-major m x l
    double* qd = out_dev ? q_out : static_cast<double*>(sc.ctx->get("Uout", (size_t)m * l * 8));
    if (!qd) { set_last_error("device allocation failed"); return CORRLA_ERR_ALLOC; }
    ST_TRY(c.mm(c.view_rows(c.Y, m), true, c.Tf, qd, 1, m, l));
    if (tm) { CU_TRY(cudaEventRecord(ev1, sc.st)); }
    if (!out_dev) {
      CU_TRY(cudaStreamSynchronize(sc.st));
      pretouch.join();
      Timer t; CU_TRY(copy_d2h_2d(sc.ctx->bounce, sc.st, q_out, (size_t)m * l * 8, qd, (size_t)m * l * 8, (size_t)m * l * 8, 1)); d2h_ms = t.ms();
    }
  } else {
    // B^T = Z_B = (A^T Y) Tf                                       random_svd.rs:80
    ST_TRY(c.mm_AtY(c.Y, c.Zb));
    ST_TRY(c.mm(c.view_rows(c.Zb, n), true, c.Tf, c.Za, c.ld, 1, c.Lc));
    // SVD of B (:89): QR-precondition Z_B, Jacobi on the l x l core W = Qz^T Z_B
    CU_TRY(cudaMemcpyAsync(c.Qz, c.Za, (size_t)c.n16 * c.ld * 8, cudaMemcpyDeviceToDevice, sc.st));
    ST_TRY(c.qr_inplace(c.Qz, n, false, (double)n, c.Tzf));
    ST_TRY(c.mm(c.view_rows(c.Qz, n), true, c.Tzf, c.Qz, c.ld, 1, c.Lc, nullptr, nullptr, nullptr, 1));
    ST_TRY(c.mm(c.view_rows(c.Qz, n), false, c.Za, c.Wm, c.ld, 1, c.Lc));
    {
      cudaError_t e = jacobi_svd_launch(c.Wm, c.ld, l, c.sig, c.Vr, c.Ur, c.L16, c.ld, c.jscratch, c.flags + 4, sc.st);
      ++c.launches;
      if (e != cudaSuccess) { set_last_error("jacobi launch failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return CORRLA_ERR_CUDA; }
    }
    // W = Ur S Vr^T  =>  B = Vr S (Qz Ur)^T : U_thin = Q Vr[:, :k] = Y (Tf Vr)   (:92),  V_thin = Qz Ur[:, :k]
    ST_TRY(c.mm(MatView{c.Tf, (int64_t)c.Lc, (int64_t)c.Lc, (int64_t)c.ld}, true, c.Vr, c.M1, c.ld, 1, c.Lc));
    const int kk = (int)k;
    // output placement (:96-109): thin-U is m x k, thin-V is n x k
    //   tall input : u <- U (col-major m x k),  vt <- V^T (col-major k x n  == V row-major n x k)
    //   fat input  : u <- V (col-major n x k),  vt <- U^T (col-major k x m  == U row-major m x k)
    const bool want_u = (u != nullptr);
    double* ud = !want_u ? nullptr : (out_dev ? u : static_cast<double*>(sc.ctx->get("Uout", (size_t)nrows * kk * 8)));
    double* vd = out_dev ? vt : static_cast<double*>(sc.ctx->get("Vout", (size_t)ncols * kk * 8));
    double* sd = out_dev ? s : c.sig;
    if ((want_u && !ud) || !vd) { set_last_error("device allocation of outputs failed"); return CORRLA_ERR_ALLOC; }
    double* Uthin_dst = fat ? vd : ud;
    double* Vthin_dst = fat ? ud : vd;
    const int64_t u_rs = fat ? kk : 1, u_cs = fat ? 1 : m;
    const int64_t v_rs = fat ? 1 : kk, v_cs = fat ? n : 1;
    if (Uthin_dst != nullptr) ST_TRY(c.mm(c.view_rows(c.Y, m), true, c.M1, Uthin_dst, u_rs, u_cs, kk));
    if (Vthin_dst != nullptr) ST_TRY(c.mm(c.view_rows(c.Qz, n), true, c.Ur, Vthin_dst, v_rs, v_cs, kk));
    if (out_dev) CU_TRY(cudaMemcpyAsync(s, c.sig, (size_t)kk * 8, cudaMemcpyDeviceToDevice, sc.st));
    if (tm) { CU_TRY(cudaEventRecord(ev1, sc.st)); }
    if (!out_dev) {
      CU_TRY(cudaStreamSynchronize(sc.st));
      pretouch.join();
      Timer t;
      if (want_u) CU_TRY(copy_d2h_2d(sc.ctx->bounce, sc.st, u, (size_t)nrows * kk * 8, ud, (size_t)nrows * kk * 8, (size_t)nrows * kk * 8, 1));
      CU_TRY(copy_d2h_2d(sc.ctx->bounce, sc.st, vt, (size_t)ncols * kk * 8, vd, (size_t)ncols * kk * 8, (size_t)ncols * kk * 8, 1));
      CU_TRY(cudaMemcpy(s, sd, (size_t)kk * 8, cudaMemcpyDeviceToHost));
      d2h_ms = t.ms();
    }
  }
  if (tm) {
    int hflags[16];
    CU_TRY(cudaMemcpyAsync(hflags, c.flags, sizeof(hflags), cudaMemcpyDeviceToHost, sc.st));
    CU_TRY(cudaStreamSynchronize(sc.st));
    float ms = 0.f;
    CU_TRY(cudaEventElapsedTime(&ms, ev0, ev1));
    double pass_ms = 0.0;
    for (int i = 0; i < c.n_pass_events; ++i) {
      float pm = 0.f;
      if (cudaEventElapsedTime(&pm, sc.ctx->event(2 + 2 * (size_t)i), sc.ctx->event(3 + 2 * (size_t)i)) == cudaSuccess) pass_ms += pm;
      else cudaGetLastError();
    }
    tm->pass_launches = c.n_pass_events; tm->pass_ms = pass_ms; tm->pass_flops = 2.0 * (double)m * (double)n * (double)l;
    tm->p2p_exchanges = c.p2p_exchanges;
    tm->streamed_chunks = n_chunks;
    if (o.comm != nullptr && o.comm->p2p) {
      int herr = 0;
      CU_TRY(cudaMemcpy(&herr, o.comm->err_flag, sizeof(int), cudaMemcpyDeviceToHost));
      if (herr != 0) { set_last_error("peer-memory exchange timed out: a rank never published its epoch"); return CORRLA_ERR_COMM; }
    }
    tm->device_ms = ms; tm->d2h_ms = d2h_ms; tm->gpu_launches = c.launches;
    tm->passes_over_a = 2 + 2 * (int)n_iter - (power_only ? 1 : 0);
    tm->qr_third_passes = c.n_robust; tm->qr_refills = c.n_refill; tm->jacobi_sweeps = hflags[4]; tm->live_columns = hflags[1] ? hflags[1] : l;
    tm->total_ms = total.ms();
  } else if (out_dev) {
    // nothing to wait for: results are ordered on the caller's stream
  }
  return CORRLA_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

void corrla_rsvd_opts_default(corrla_rsvd_opts* opts) { if (opts) *opts = default_opts(); }

int corrla_rsvd_f64(const double* a, int64_t nrows, int64_t ncols, int64_t row_stride, int64_t col_stride,
                    size_t n_rank, size_t n_iter, size_t n_oversamples, const corrla_rsvd_opts* opts, double* u,
                    double* s, double* vt, corrla_timings* timings) {
  try {
    return rsvd_impl(a, nrows, ncols, row_stride, col_stride, n_rank, n_iter, n_oversamples, opts, u, s, vt, timings,
                     false, nullptr);
  } catch (const std::exception& e) { set_last_error("exception: %s", e.what()); return CORRLA_ERR_ALLOC; }
  catch (...) { set_last_error("unknown exception"); return CORRLA_ERR_INVALID; }
}

int corrla_power_iter_f64(const double* a, int64_t nrows, int64_t ncols, int64_t row_stride, int64_t col_stride,
                          size_t omega_rank, size_t n_iter, const corrla_rsvd_opts* opts, double* q,
                          corrla_timings* timings) {
  try {
    return rsvd_impl(a, nrows, ncols, row_stride, col_stride, omega_rank, n_iter, 0, opts, nullptr, nullptr, nullptr,
                     timings, true, q);
  } catch (const std::exception& e) { set_last_error("exception: %s", e.what()); return CORRLA_ERR_ALLOC; }
  catch (...) { set_last_error("unknown exception"); return CORRLA_ERR_INVALID; }
}

int corrla_rpca_f64(const double* a, int64_t nrows, int64_t ncols, int64_t row_stride, int64_t col_stride, size_t n_rank,
                    const corrla_rsvd_opts* opts, double* s, double* components, double* means, corrla_timings* timings) {
  try {
    corrla_rsvd_opts o = opts ? *opts : default_opts();
    o.center = 1;
    // PcaRsvd::new (pca_rsvd.rs:65-66): 20 power iterations, min(n_dim, 10) oversamples, U discarded
    const size_t n_over = (size_t)std::min<int64_t>(ncols, 10);
    return rsvd_impl(a, nrows, ncols, row_stride, col_stride, n_rank, 20, n_over, &o, nullptr, s, components, timings,
                     false, nullptr, true, means);
  } catch (const std::exception& e) { set_last_error("exception: %s", e.what()); return CORRLA_ERR_ALLOC; }
  catch (...) { set_last_error("unknown exception"); return CORRLA_ERR_INVALID; }
}

int corrla_par_matmul_f64(double* res, int64_t res_rs, int64_t res_cs, const double* lhs, int64_t lhs_rows,
                          int64_t lhs_cols, int64_t lhs_rs, int64_t lhs_cs, const double* rhs, int64_t rhs_cols,
                          int64_t rhs_rs, int64_t rhs_cs, double beta, int on_device, const corrla_rsvd_opts* opts) {
  try {
    if (!res || !lhs || !rhs || lhs_rows <= 0 || lhs_cols <= 0 || rhs_cols <= 0) { set_last_error("bad argument"); return CORRLA_ERR_INVALID; }
    Scope sc;
    ST_TRY(open_scope(opts, &sc));
    Core c; c.ctx = sc.ctx; c.st = sc.st;
    ST_TRY(c.setup_dims(lhs_rows, lhs_cols, (int)std::min<int64_t>(rhs_cols, 1 << 20)));
    ST_TRY(stage_matrix(sc.ctx, sc.st, "A", lhs, lhs_rows, lhs_cols, lhs_rs, lhs_cs, on_device != 0, &c.av, &c.a_rowmajor, nullptr, &c.launches));
    ST_TRY(c.alloc_workspace(false));
    double* X = static_cast<double*>(sc.ctx->get("Za", (size_t)c.n16 * c.ld * 8));
    if (!X) { set_last_error("device allocation failed"); return CORRLA_ERR_ALLOC; }
    CU_TRY(cudaMemsetAsync(X, 0, (size_t)c.n16 * c.ld * 8, sc.st));
    ST_TRY(pack_small(sc.ctx, sc.st, rhs, lhs_cols, rhs_cols, rhs_rs, rhs_cs, on_device != 0, X, c.ld, beta, &c.launches));
    const int nc = (int)rhs_cols;
    if (on_device) {
      ST_TRY(c.mm(c.av, c.a_rowmajor, X, res, res_rs, res_cs, nc));
      CU_TRY(cudaStreamSynchronize(sc.st));
    } else {
      double* out = static_cast<double*>(sc.ctx->get("Uout", (size_t)lhs_rows * nc * 8));
      if (!out) { set_last_error("device allocation failed"); return CORRLA_ERR_ALLOC; }
      ST_TRY(c.mm(c.av, c.a_rowmajor, X, out, nc, 1, nc));
      std::vector<double> tmp((size_t)lhs_rows * nc);
      CU_TRY(cudaMemcpyAsync(tmp.data(), out, tmp.size() * 8, cudaMemcpyDeviceToHost, sc.st));
      CU_TRY(cudaStreamSynchronize(sc.st));
      for (int64_t i = 0; i < lhs_rows; ++i)
        for (int j = 0; j < nc; ++j) res[i * res_rs + j * res_cs] = tmp[(size_t)i * nc + j];
    }
    return CORRLA_OK;
  } catch (const std::exception& e) { set_last_error("exception: %s", e.what()); return CORRLA_ERR_ALLOC; }
  catch (...) { set_last_error("unknown exception"); return CORRLA_ERR_INVALID; }
}

int corrla_random_mat_normal_f64(uint64_t seed, int64_t n_rows, int64_t n_cols, double* out, int out_on_device,
                                 const corrla_rsvd_opts* opts) {
  try {
    if (!out || n_rows <= 0 || n_cols <= 0 || n_cols > (1 << 30)) { set_last_error("bad argument"); return CORRLA_ERR_INVALID; }
    Scope sc;
    ST_TRY(open_scope(opts, &sc));
    const size_t elems = (size_t)n_rows * n_cols;
    double* rm = static_cast<double*>(sc.ctx->get("omega_rm", elems * 8));
    if (!rm) { set_last_error("device allocation failed"); return CORRLA_ERR_ALLOC; }
    CU_TRY(philox_normal_launch(rm, n_rows, (int)n_cols, n_cols, seed, sc.st));
    // column-major output, as faer's Mat::from_fn stores it
    double* cm = out_on_device ? out : static_cast<double*>(sc.ctx->get("omega_cm", elems * 8));
    if (!cm) { set_last_error("device allocation failed"); return CORRLA_ERR_ALLOC; }
    CU_TRY(scatter_launch(rm, n_rows, n_cols, n_cols, cm, 1, n_rows, sc.st));
    if (!out_on_device) CU_TRY(cudaMemcpyAsync(out, cm, elems * 8, cudaMemcpyDeviceToHost, sc.st));
    CU_TRY(cudaStreamSynchronize(sc.st));
    return CORRLA_OK;
  } catch (...) { set_last_error("exception"); return CORRLA_ERR_ALLOC; }
}

int corrla_thin_q_f64(const double* a, int64_t nrows, int64_t ncols, int64_t row_stride, int64_t col_stride,
                      int on_device, const corrla_rsvd_opts* opts, double* q, int* rank_out) {
  try {
    if (!a || !q || nrows <= 0 || ncols <= 0) { set_last_error("bad argument"); return CORRLA_ERR_INVALID; }
    Scope sc;
    ST_TRY(open_scope(opts, &sc));
    Core c; c.ctx = sc.ctx; c.st = sc.st; c.comm = opts ? opts->comm : nullptr;
    ST_TRY(c.setup_dims(nrows, ncols, (int)std::min<int64_t>(ncols, 1 << 20)));
    c.grows = (opts && opts->global_rows > 0) ? (double)opts->global_rows : (double)nrows;
    ST_TRY(c.alloc_workspace(false));
    ST_TRY(c.alloc_buffers(false));
    ST_TRY(pack_small(sc.ctx, sc.st, a, nrows, ncols, row_stride, col_stride, on_device != 0, c.Y, c.ld, 1.0, &c.launches));
    ST_TRY(c.qr_inplace(c.Y, nrows, c.comm != nullptr, c.grows, c.Tf));
    double* qd = on_device ? q : static_cast<double*>(sc.ctx->get("Uout", (size_t)nrows * ncols * 8));
    if (!qd) { set_last_error("device allocation failed"); return CORRLA_ERR_ALLOC; }
    ST_TRY(c.mm(c.view_rows(c.Y, nrows), true, c.Tf, qd, 1, nrows, (int)ncols));
    int hinfo[4] = {0, 0, 0, 0};
    CU_TRY(cudaMemcpyAsync(hinfo, c.flags + 1, 8, cudaMemcpyDeviceToHost, sc.st));
    if (!on_device) CU_TRY(cudaMemcpyAsync(q, qd, (size_t)nrows * ncols * 8, cudaMemcpyDeviceToHost, sc.st));
    CU_TRY(cudaStreamSynchronize(sc.st));
    if (rank_out) *rank_out = hinfo[0] ? hinfo[0] : (int)ncols;   // 0: the probe passed, plain CholeskyQR2, full rank
    return CORRLA_OK;
  } catch (...) { set_last_error("exception"); return CORRLA_ERR_ALLOC; }
}

void* corrla_host_alloc(size_t bytes) {
  if (bytes == 0) return nullptr;
  const size_t two_mb = (size_t)2 << 20;
  const size_t len = (bytes + two_mb - 1) / two_mb * two_mb;
  void* p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if (p == MAP_FAILED) return nullptr;
#ifdef MADV_HUGEPAGE
  madvise(p, len, MADV_HUGEPAGE);
#endif
  return p;
}
void corrla_host_free(void* p, size_t bytes) {
  if (p == nullptr) return;
  const size_t two_mb = (size_t)2 << 20;
  munmap(p, (bytes + two_mb - 1) / two_mb * two_mb);
}

int corrla_ctx_create(int device, corrla_ctx** out) {
  if (!out) return CORRLA_ERR_INVALID;
  try { return ctx_create(device, out); } catch (...) { return CORRLA_ERR_ALLOC; }
}
void corrla_ctx_destroy(corrla_ctx* ctx) { delete ctx; }

int corrla_comm_unique_id(unsigned char id[128]) { return id ? comm_unique_id(id) : CORRLA_ERR_INVALID; }
int corrla_comm_init(const unsigned char id[128], int rank, int nranks, int device, corrla_comm** out) {
  if (!id || !out) return CORRLA_ERR_INVALID;
  int st = ensure_device(device);
  if (st != CORRLA_OK) return st;
  return comm_init(id, rank, nranks, device, out);
}
void corrla_comm_destroy(corrla_comm* comm) { comm_destroy(comm); }
int corrla_comm_rank(const corrla_comm* comm) { return comm ? comm->rank : 0; }
int corrla_comm_size(const corrla_comm* comm) { return comm ? comm->nranks : 1; }

const char* corrla_status_str(int status) {
  switch (status) {
    case CORRLA_OK: return "ok";
    case CORRLA_ERR_INVALID: return "invalid argument";
    case CORRLA_ERR_RANK: return "n_rank exceeds min(n_rank + n_oversamples, ncols of the thin matrix)";
    case CORRLA_ERR_CUDA: return "CUDA error";
    case CORRLA_ERR_UNSUPPORTED: return "unsupported size (n_rank + n_oversamples > 128)";
    case CORRLA_ERR_ALLOC: return "allocation failed";
    case CORRLA_ERR_COMM: return "communicator (NCCL) error";
    case CORRLA_ERR_NO_DEVICE: return "no CUDA device (no CPU fallback exists)";
    default: return "unknown status";
  }
}
const char* corrla_last_error(void) { return last_error_cstr(); }
const char* corrla_version(void) { return "corrla_b200 0.1.0 (sm_100a, DMMA+TMA)"; }

}  // extern "C"
