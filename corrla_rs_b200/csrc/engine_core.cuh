// Shared internals of the engine (context, the per-call Core with its GEMM/QR building blocks, staging helpers).
// Included by engine.cu (RSVD / PCA / QR entry points) and rom.cu (DMDc and POD on top of the same kernels).
#pragma once
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
  std::recursive_mutex mu;
  struct Buf { void* p = nullptr; size_t bytes = 0; };
  std::map<std::string, Buf> pool;
  void* pinned = nullptr; size_t pinned_bytes = 0;
  BounceBuffers bounce;
  int* hflag = nullptr;              // pinned, 64 ints: [0,2) decisions read back inside a QR, [16,49) end-of-call flags
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
  // Give the cached device buffers (and the pinned staging memory) back to the driver.  cudaFree waits for the device,
  // so buffers still referenced by queued work are safe.  The next call re-allocates what it needs.
  size_t trim() {
    size_t freed = 0;
    for (auto& kv : pool) if (kv.second.p) { cudaFree(kv.second.p); freed += kv.second.bytes; kv.second.p = nullptr; kv.second.bytes = 0; }
    pool.clear();
    if (pinned) { cudaFreeHost(pinned); freed += pinned_bytes; pinned = nullptr; pinned_bytes = 0; }
    freed += 2 * bounce.bytes;
    bounce.release();
    cudaGetLastError();
    return freed;
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

namespace corrla_eng {

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

// l columns as P panels of equal padded width w (a multiple of 8, <= 128); only the last panel has padding columns
inline void panel_plan(int l, int* P, int* w) {
  const int lc = (l + 7) / 8 * 8;
  *P = (lc + 127) / 128;
  *w = ((l + *P - 1) / *P + 7) / 8 * 8;
}

inline int ensure_device(int device) {
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

inline int ctx_create(int device, corrla_ctx** out) {
  ST_TRY(ensure_device(device));
  corrla_ctx* c = new corrla_ctx();
  if (device < 0) cudaGetDevice(&device);
  c->device = device;
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, device) == cudaSuccess) c->num_sms = p.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaMallocHost(reinterpret_cast<void**>(&c->hflag), 256) != cudaSuccess) {
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
    deadmask = reinterpret_cast<int*>(getz("deadmask", 2 * (size_t)L16));     // ints: [0, L16) main phase, [L16, 2*L16) refill phase
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
    e = refill_phase ? hqr_inv_launch(SK, ld, s_rows, l, T1, L16, ld, flags + 12, deadmask + L16, flags + 14, cond, st)
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

  // basis_only: the caller only needs a well-conditioned basis of range(X), not an orthonormal one (the re-orthonormalisation
  // INSIDE the power loop, random_svd.rs:38: the next product A^T(X*Tfold) depends on the range alone).  When the probe
  // passes (cond(X) below ~1e4) one Cholesky pass is then enough: X stays untouched, Tfold = R^-1, X*Tfold is orthonormal to
  // cond(X)^2 * eps <~ 1e-8 -- a basis of condition 1 + 1e-8.  That removes the apply, the second Gram product and the
  // first-order correction from every in-loop QR (two passes over the m x l iterate).
  int qr_inplace(double* X, int64_t rows, bool distributed, double rows_for_shift, double* Tfold,
                 bool refill_from_a = false, bool complete = false, bool basis_only = false) {
    const MatView vx = view_rows(X, rows);
    const size_t gx = distributed ? (size_t)Lc * ld : 0;
    CU_TRY(cudaMemsetAsync(flags, 0, 8 * sizeof(int), st));      // f2 (0), liveness (1..3), refill-stage f2 (6) start clear
    int* fs = flags + 16;                                        // [0] robust stage needed, [1] fast path ok
    ST_TRY(gram(vx, X, nullptr, gx));
    ST_TRY(chol(kCholProbe, rows_for_shift, basis_only ? Tfold : T1, nullptr, true, fs));
    int robust = 0;
    ST_TRY(read_flags(fs, &robust, nullptr));                    // identical on every rank: G is the all-reduced Gram
    if (!robust && basis_only) { ++qr_calls; return CORRLA_OK; }
    if (!robust) {
      // fast path: CholeskyQR2
      // ... with the second Cholesky replaced by its first-order expansion: the probe bounds cond(G) by ~1e8, so after
      // the first pass G2 = I + E with |E| <~ 1e-8 and X*(3/2 I - 1/2 G2) is orthonormal to (3/8)E^2 ~ 1e-16.  Tfold is
      // then symmetric, not triangular: every consumer of Tfold multiplies with the general GEMM.
      ST_TRY(apply_tri(vx, T1, X, nullptr));
      ST_TRY(gram(vx, X, nullptr, gx));
      {
        cudaError_t e = lowdin_launch(G, ld, l, Tfold, L16, ld, flags + 18, st);
        ++launches;
        if (e != cudaSuccess) { set_last_error("lowdin launch failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return CORRLA_ERR_CUDA; }
      }
      // ||E||_F above 2e-8 (possible near the probe threshold on very tall matrices): the real second Cholesky runs,
      // behind the device flag -- an empty launch otherwise
      ST_TRY(chol(kCholPlain, rows_for_shift, Tfold, flags + 18, true, nullptr));
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
      if (refill_from_a && complete) {
        // range(A) itself has fewer than l dimensions: the columns that are still dead get arbitrary vectors, made
        // orthogonal to the live ones -- Householder's completion; Q keeps l orthonormal columns (random_svd.rs:57).
        int still_dead = 0;
        ST_TRY(read_flags(flags + 14, &still_dead, nullptr));
        if (still_dead) {
          ST_TRY(apply_tri(vx, Tfold, X, nullptr));
          cudaError_t e = refill_dead_launch(X, rows, l, ld, deadmask + L16, refill_seed + 0x9e37ull + (uint64_t)qr_calls, refill_stream, nullptr, st);
          ++launches;
          if (e != cudaSuccess) { set_last_error("refill launch failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return CORRLA_ERR_CUDA; }
          ST_TRY(qr_stage(X, rows, distributed, rows_for_shift, Tfold, nullptr, flags + 6, true));
        }
      }
    }
    ++qr_calls;
    return CORRLA_OK;
  }

  int n_robust = 0, n_refill = 0;
  bool basis_only_qr = true;      // in-loop QR as one Cholesky pass when the probe allows (CORRLA_B200_INLOOP_CHOLQR2=1: two)


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
        ST_TRY(qr_inplace(Y, m, true, grows, Tf, true, false, basis_only_qr));   // :38
        ST_TRY(mm_AtY(Y, Zb));                                // :42-46 (on the pre-fold iterate), all-reduced
        ST_TRY(mm(view_rows(Zb, n), true, Tf, Za, ld, 1, Lc)); // fold R^-1 into the small side
        ST_TRY(mm_AX(Za, Y, nullptr, nu2));                   // :47-51
      } else {
        if (!(i == 0 && resume == 2)) ST_TRY(mm_AtY(Y, Zb));
        ST_TRY(mm_AX(Zb, Y, nu2, nu2));                       // :47-51 with the deferred :53-55 scaling
      }
    }
    ST_TRY(qr_inplace(Y, m, true, grows, Tf, true, true));    // :57
    return CORRLA_OK;
  }
};


struct Timer {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  double ms() const { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};

// Describe a strided matrix as a TMA-friendly view if possible.
// rows x cols with element strides (rs, cs).  Returns 1 = row-major-like, 2 = column-major-like, 0 = needs repack.
inline int classify(const double* p, int64_t rows, int64_t cols, int64_t rs, int64_t cs, MatView* v) {
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
inline int stage_matrix(corrla_ctx* ctx, cudaStream_t st, const char* bufname, const double* a, int64_t rows, int64_t cols,
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
inline int pack_small(corrla_ctx* ctx, cudaStream_t st, const double* src, int64_t rows, int64_t cols, int64_t rs, int64_t cs,
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
  std::unique_lock<std::recursive_mutex> lock;
  std::unique_lock<std::recursive_mutex> comm_lock;      // opts.comm, if any: one call at a time per communicator
  ~Scope() { if (comm_lock.owns_lock()) comm_lock.unlock(); if (lock.owns_lock()) lock.unlock(); if (owned) delete ctx; }
};

inline int open_scope(const corrla_rsvd_opts* o, Scope* s) {
  const int device = o ? o->device : -1;
  ST_TRY(ensure_device(device));
  if (o && o->ctx) { s->ctx = o->ctx; s->owned = false; CU_TRY(cudaSetDevice(s->ctx->device)); }
  else { ST_TRY(ctx_create(device, &s->ctx)); s->owned = true; }
  s->lock = std::unique_lock<std::recursive_mutex>(s->ctx->mu);
  if (o && o->comm) s->comm_lock = std::unique_lock<std::recursive_mutex>(o->comm->call_mu);
  s->st = (o && o->stream) ? static_cast<cudaStream_t>(o->stream) : s->ctx->own_stream;
  return CORRLA_OK;
}

inline corrla_rsvd_opts default_opts() {
  corrla_rsvd_opts o;
  memset(&o, 0, sizeof(o));
  o.device = -1;
  return o;
}

// (U, S, Vt) / Q of a host- or device-resident matrix; the body of corrla_rsvd_f64, corrla_power_iter_f64 and
// corrla_rpca_f64 (engine.cu).  The context mutex is recursive: composite entry points call this while holding it.
int rsvd_impl(const double* a, int64_t nrows, int64_t ncols, int64_t rs, int64_t cs, size_t n_rank, size_t n_iter,
              size_t n_oversamples, const corrla_rsvd_opts* opts_in, double* u, double* s, double* vt,
              corrla_timings* tm, bool power_only, double* q_out, bool u_optional = false, double* means_out = nullptr);

}  // namespace corrla_eng
