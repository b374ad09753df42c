// Host-side engine and C ABI of libcorrla_b200.so: orchestrates the RSVD of the reference
// (random_svd.rs:15-110) as a sequence of skinny DMMA GEMMs, CholeskyQR and a Jacobi SVD, all on one
// CUDA stream, with NCCL all-reduces of the small replicated factors when the rows are sharded.
#include "engine_core.cuh"
#include "fused_small.cuh"
#include "wide.cuh"

namespace corrla_eng {

// Every call ends here, with or without a timings struct: the device-side decision and failure flags come back through
// pinned memory behind ONE stream synchronisation (the one host outputs need anyway before their copy), and a failure
// reported by a kernel -- a peer that never published its epoch in the fused all-reduce, a lost DSMEM transaction in
// the Jacobi cluster -- becomes a status code instead of silently wrong numbers.  The communicator's error flag is
// cleared once reported, so the next call on it starts clean.
static int finish_device(Scope& sc, Core& c, corrla_comm* comm, bool power_only, int* hflags /* 32 */) {
  int* pin = sc.ctx->hflag + 16;                    // [16, 48): the run's flags, [48]: peer-exchange error
  pin[32] = 0;
  CU_TRY(cudaMemcpyAsync(pin, c.flags, 32 * sizeof(int), cudaMemcpyDeviceToHost, sc.st));
  const bool p2p = comm != nullptr && comm->p2p && comm->err_flag != nullptr;
  if (p2p) CU_TRY(cudaMemcpyAsync(pin + 32, comm->err_flag, sizeof(int), cudaMemcpyDeviceToHost, sc.st));
  CU_TRY(cudaStreamSynchronize(sc.st));
  memcpy(hflags, pin, 32 * sizeof(int));
  if (p2p && pin[32] != 0) {
    cudaMemsetAsync(comm->err_flag, 0, sizeof(int), sc.st);
    set_last_error("peer-memory exchange timed out: a rank never published its epoch (results are invalid)");
    return CORRLA_ERR_COMM;
  }
  if (!power_only && hflags[5] < 0) {
    set_last_error("the Jacobi kernel's cluster exchange timed out (results are invalid)");
    return CORRLA_ERR_CUDA;
  }
  return CORRLA_OK;
}

// Tiny problems: the whole call as ONE kernel with the matrix in shared memory (fused_small.cu).  Returns true when the
// call was served (its status is in *status); false when the path does not apply or the kernel asked for the general one
// (an exactly zero singular value among the first k: no direction to normalise).
static bool try_fused_small(Scope& sc, const corrla_rsvd_opts& o, const double* a, int64_t m, int64_t n, int64_t trs, int64_t tcs,
                            bool fat, int l, size_t k, size_t n_iter, int64_t nrows, int64_t ncols, double* u, double* s,
                            double* vt, corrla_timings* tm, bool power_only, double* q_out, Timer& total, int* status) {
  const char* env_off = getenv("CORRLA_B200_NO_FUSED_SMALL");      // read per call: tests compare the two paths in one process
  const bool disabled = env_off != nullptr && env_off[0] == '1';
  if (disabled || o.comm != nullptr || o.center != 0 || l > kFusedMaxL || m > 4096) return false;
  const int kk = power_only ? l : (int)k;
  if (fused_small_smem_bytes((int)m, (int)n, l, kk) == 0) return false;
  corrla_ctx* ctx = sc.ctx;
  cudaStream_t st = sc.st;
  auto fail = [&](int code) { *status = code; return true; };
  auto cu_fail = [&](cudaError_t e, const char* what) {
    set_last_error("%s failed: %s", what, cudaGetErrorString(e)); cudaGetLastError(); *status = CORRLA_ERR_CUDA; return true;
  };
  const bool in_dev = o.a_on_device != 0, out_dev = o.out_on_device != 0;
  FusedSmallArgs fa{};
  fa.m = (int)m; fa.n = (int)n; fa.l = l; fa.k = kk; fa.n_iter = (int)n_iter; fa.schedule = o.schedule; fa.power_only = power_only ? 1 : 0;
  fa.seed = o.seed;
  fa.debug = getenv("CORRLA_B200_FUSED_PROFILE") != nullptr ? 1 : 0;
  { const char* e = getenv("CORRLA_B200_FUSED_NO_CHOL"); fa.no_chol = (e != nullptr && e[0] == '1') ? 1 : 0; }
  { const char* e = getenv("CORRLA_B200_INLOOP_CHOLQR2"); fa.basis_only = (e != nullptr && e[0] == '1') ? 0 : 1; }
  // staging: [A m*n | Omega n*l] in, [U m*k | V n*k | S k | Q m*l] out
  const size_t in_elems = (in_dev ? 0 : (size_t)m * n) + ((o.omega != nullptr && !o.omega_on_device) ? (size_t)n * l : 0);
  const size_t out_elems = out_dev ? 0 : (power_only ? (size_t)m * l : (size_t)(m + n + 1) * kk);
  double* dstage = static_cast<double*>(ctx->get("fs_stage", (in_elems + out_elems + 8) * 8));
  int* info = reinterpret_cast<int*>(ctx->get("fs_info", 64));
  double* hstage = (in_elems + out_elems) ? static_cast<double*>(ctx->get_pinned((in_elems + out_elems + 8) * 8)) : nullptr;
  if (!dstage || !info || ((in_elems + out_elems) && !hstage)) { set_last_error("allocation failed (fused small path)"); return fail(CORRLA_ERR_ALLOC); }
  Timer th2d;
  size_t off = 0;
  if (in_dev) { fa.a = a; fa.a_rs = trs; fa.a_cs = tcs; }
  else {
    for (int64_t i = 0; i < m; ++i)
      for (int64_t j = 0; j < n; ++j) hstage[off + i * n + j] = a[i * trs + j * tcs];
    fa.a = dstage + off; fa.a_rs = n; fa.a_cs = 1;
    off += (size_t)m * n;
  }
  if (o.omega != nullptr) {
    if (o.omega_on_device) { fa.omega = o.omega; fa.om_rs = o.omega_rs; fa.om_cs = o.omega_cs; }
    else {
      for (int64_t i = 0; i < n; ++i)
        for (int j = 0; j < l; ++j) hstage[off + i * l + j] = o.omega[i * o.omega_rs + j * o.omega_cs];
      fa.omega = dstage + off; fa.om_rs = l; fa.om_cs = 1;
      off += (size_t)n * l;
    }
  }
  if (off > 0) {
    cudaError_t e = cudaMemcpyAsync(dstage, hstage, off * 8, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return cu_fail(e, "cudaMemcpyAsync (fused small input)");
  }
  const double h2d_ms = th2d.ms();
  // outputs (:96-109): thin-U is m x k, thin-V is n x k; a fat input swaps their roles
  double* out_base = dstage + in_elems;
  double *sd = nullptr, *qd = nullptr;
  if (power_only) { qd = out_dev ? q_out : out_base; fa.qout = qd; }
  else {
    double* udst = out_dev ? u : (u != nullptr ? out_base : nullptr);               // nrows x k column-major
    double* vdst = out_dev ? vt : out_base + (size_t)nrows * kk;                    // k x ncols column-major
    sd = out_dev ? s : out_base + (size_t)(nrows + ncols) * kk;
    if (!fat) { fa.u = udst; fa.u_rs = 1; fa.u_cs = m; fa.v = vdst; fa.v_rs = kk; fa.v_cs = 1; }
    else      { fa.u = vdst; fa.u_rs = kk; fa.u_cs = 1; fa.v = udst; fa.v_rs = 1; fa.v_cs = n; }
    fa.s = sd;
  }
  fa.info = info;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (tm) { ev0 = ctx->event(0); ev1 = ctx->event(1); if (ev0 && ev1) cudaEventRecord(ev0, st); }
  cudaError_t e = fused_small_launch(fa, st);
  if (e != cudaSuccess) return cu_fail(e, "fused small kernel launch");
  if (tm && ev0 && ev1) cudaEventRecord(ev1, st);
  int* pin = ctx->hflag + 16;
  e = cudaMemcpyAsync(pin, info, 4 * sizeof(int), cudaMemcpyDeviceToHost, st);
  Timer td2h;
  if (e == cudaSuccess && out_elems > 0) e = cudaMemcpyAsync(hstage + in_elems, out_base, out_elems * 8, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return cu_fail(e, "fused small kernel");
  if (pin[2] != 0) return false;                               // general path (it owns the orthonormal completion)
  if (!out_dev) {
    const double* h = hstage + in_elems;
    if (power_only) memcpy(q_out, h, (size_t)m * l * 8);
    else {
      if (u != nullptr) memcpy(u, h, (size_t)nrows * kk * 8);
      memcpy(vt, h + (size_t)nrows * kk, (size_t)ncols * kk * 8);
      memcpy(s, h + (size_t)(nrows + ncols) * kk, (size_t)kk * 8);
    }
  }
  (void)qd;
  if (tm) {
    float ms = 0.f;
    if (ev0 && ev1 && cudaEventElapsedTime(&ms, ev0, ev1) != cudaSuccess) { cudaGetLastError(); ms = 0.f; }
    tm->h2d_ms = h2d_ms; tm->device_ms = ms; tm->d2h_ms = out_dev ? 0.0 : td2h.ms(); tm->gpu_launches = 1;
    tm->passes_over_a = 2 + 2 * (int)n_iter - (power_only ? 1 : 0);
    tm->jacobi_sweeps = pin[0]; tm->jacobi_converged = power_only ? 1 : pin[1]; tm->live_columns = l;
    tm->pass_launches = 0; tm->pass_ms = 0.0; tm->pass_flops = 2.0 * (double)m * (double)n * (double)l;
    tm->fused_small = 1;
    tm->total_ms = total.ms();
  }
  *status = CORRLA_OK;
  return true;
}

int rsvd_impl(const double* a, int64_t nrows, int64_t ncols, int64_t rs, int64_t cs, size_t n_rank, size_t n_iter,
              size_t n_oversamples, const corrla_rsvd_opts* opts_in, double* u, double* s, double* vt,
              corrla_timings* tm, bool power_only, double* q_out, bool u_optional, double* means_out) {
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
  {
    int fs_status = CORRLA_OK;
    if (try_fused_small(sc, o, a, m, n, trs, tcs, fat, l, k, n_iter, nrows, ncols, u, s, vt, tm, power_only, q_out, total, &fs_status))
      return fs_status;
  }
  Core c;
  c.ctx = sc.ctx; c.st = sc.st; c.comm = o.comm;
  c.refill_seed = o.seed ^ 0x9e3779b97f4a7c15ull;
  c.refill_stream = o.comm ? (uint64_t)o.comm->rank : 0;
  { const char* e = getenv("CORRLA_B200_INLOOP_CHOLQR2"); c.basis_only_qr = !(e != nullptr && e[0] == '1'); }   // A/B switch, read per call
  // more than 128 sketch columns: P column panels of padded width w <= 128 (wide.cuh); c then describes ONE panel
  const bool wide = l > 8 * kMaxNblk;
  if (wide && l > 2048) { set_last_error("n_rank + n_oversamples = %d exceeds the supported 2048 sketch columns", l); return CORRLA_ERR_UNSUPPORTED; }
  int wide_P = 1, wide_w = l;
  if (wide) Wide::plan(l, &wide_P, &wide_w);
  ST_TRY(c.setup_dims(m, n, wide ? wide_w : l));

  const bool want_center = (o.center != 0) && !power_only;
  if (want_center && fat && o.comm != nullptr) { set_last_error("centring of a fat matrix is not supported with a communicator"); return CORRLA_ERR_UNSUPPORTED; }
  c.center = (want_center && !fat) ? 1 : 0;       // tall: rank-1 corrections inside the passes

  // Large host-resident inputs are streamed: the first product (and the first transposed product when the schedule
  // starts without a QR and no other rank is involved) runs chunk by chunk behind the host->device copies.
  const bool host_rowmajor = (tcs == 1 && trs >= n), host_colmajor = !host_rowmajor && (trs == 1 && tcs >= m);
  int64_t stream_rows = 0;
  if (!o.a_on_device && !want_center && !wide && (host_rowmajor || host_colmajor)) {
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

  Wide wd(c);
  if (wide) ST_TRY(wd.alloc(l));
  const double* omega_packed = nullptr;
  if (o.omega != nullptr && !wide) {
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
  if (wide) ST_TRY(wd.power_iter(o, (int)n_iter));
  else ST_TRY(c.power_iter(omega_packed, o.seed, (int)n_iter, o.schedule, resume));

  const bool out_dev = o.out_on_device != 0;
  double d2h_ms = 0.0;
  int hflags[32] = {0};
  if (power_only) {
    // Q = Y * Tf, column-major m x l
    double* qd = out_dev ? q_out : static_cast<double*>(sc.ctx->get("Uout", (size_t)m * l * 8));
    if (!qd) { set_last_error("device allocation failed"); return CORRLA_ERR_ALLOC; }
    if (wide) ST_TRY(wd.scatter_q(qd));
    else ST_TRY(c.mm(c.view_rows(c.Y, m), true, c.Tf, qd, 1, m, l));
    if (tm) { CU_TRY(cudaEventRecord(ev1, sc.st)); }
    ST_TRY(finish_device(sc, c, o.comm, true, hflags));
    if (!out_dev) {
      pretouch.join();
      Timer t; CU_TRY(copy_d2h_2d(sc.ctx->bounce, sc.st, q_out, (size_t)m * l * 8, qd, (size_t)m * l * 8, (size_t)m * l * 8, 1)); d2h_ms = t.ms();
    }
  } else {
    if (!wide) {
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
    }
    // W = Ur S Vr^T  =>  B = Vr S (Qz Ur)^T : U_thin = Q Vr[:, :k] = Y (Tf Vr)   (:92),  V_thin = Qz Ur[:, :k]
    if (!wide) ST_TRY(c.mm(MatView{c.Tf, (int64_t)c.Lc, (int64_t)c.Lc, (int64_t)c.ld}, true, c.Vr, c.M1, c.ld, 1, c.Lc));
    const int kk = (int)k;
    // output placement (:96-109): thin-U is m x k, thin-V is n x k
    //   tall input : u <- U (col-major m x k),  vt <- V^T (col-major k x n  == V row-major n x k)
    //   fat input  : u <- V (col-major n x k),  vt <- U^T (col-major k x m  == U row-major m x k)
    const bool want_u = (u != nullptr);
    double* ud = !want_u ? nullptr : (out_dev ? u : static_cast<double*>(sc.ctx->get("Uout", (size_t)nrows * kk * 8)));
    double* vd = out_dev ? vt : static_cast<double*>(sc.ctx->get("Vout", (size_t)ncols * kk * 8));
    const double* sig_dev = wide ? wd.sig : c.sig;
    double* sd = out_dev ? s : const_cast<double*>(sig_dev);
    if ((want_u && !ud) || !vd) { set_last_error("device allocation of outputs failed"); return CORRLA_ERR_ALLOC; }
    double* Uthin_dst = fat ? vd : ud;
    double* Vthin_dst = fat ? ud : vd;
    const int64_t u_rs = fat ? kk : 1, u_cs = fat ? 1 : m;
    const int64_t v_rs = fat ? 1 : kk, v_cs = fat ? n : 1;
    if (wide) {
      ST_TRY(wd.finish(kk, Uthin_dst, u_rs, u_cs, Vthin_dst, v_rs, v_cs));
    } else {
      if (Uthin_dst != nullptr) ST_TRY(c.mm(c.view_rows(c.Y, m), true, c.M1, Uthin_dst, u_rs, u_cs, kk));
      if (Vthin_dst != nullptr) ST_TRY(c.mm(c.view_rows(c.Qz, n), true, c.Ur, Vthin_dst, v_rs, v_cs, kk));
    }
    if (out_dev) CU_TRY(cudaMemcpyAsync(s, sig_dev, (size_t)kk * 8, cudaMemcpyDeviceToDevice, sc.st));
    if (tm) { CU_TRY(cudaEventRecord(ev1, sc.st)); }
    ST_TRY(finish_device(sc, c, o.comm, false, hflags));
    if (!out_dev) {
      pretouch.join();
      Timer t;
      if (want_u) CU_TRY(copy_d2h_2d(sc.ctx->bounce, sc.st, u, (size_t)nrows * kk * 8, ud, (size_t)nrows * kk * 8, (size_t)nrows * kk * 8, 1));
      CU_TRY(copy_d2h_2d(sc.ctx->bounce, sc.st, vt, (size_t)ncols * kk * 8, vd, (size_t)ncols * kk * 8, (size_t)ncols * kk * 8, 1));
      CU_TRY(cudaMemcpy(s, sd, (size_t)kk * 8, cudaMemcpyDeviceToHost));
      d2h_ms = t.ms();
    }
  }
  if (tm) {
    float ms = 0.f;
    CU_TRY(cudaEventElapsedTime(&ms, ev0, ev1));
    double pass_ms = 0.0;
    for (int i = 0; i < c.n_pass_events; ++i) {
      float pm = 0.f;
      if (cudaEventElapsedTime(&pm, sc.ctx->event(2 + 2 * (size_t)i), sc.ctx->event(3 + 2 * (size_t)i)) == cudaSuccess) pass_ms += pm;
      else cudaGetLastError();
    }
    tm->pass_launches = c.n_pass_events; tm->pass_ms = pass_ms; tm->pass_flops = 2.0 * (double)m * (double)n * (double)l / (double)wide_P;
    tm->p2p_exchanges = c.p2p_exchanges;
    tm->streamed_chunks = n_chunks;
    tm->device_ms = ms; tm->d2h_ms = d2h_ms; tm->gpu_launches = c.launches;
    tm->passes_over_a = 2 + 2 * (int)n_iter - (power_only ? 1 : 0);     // each pass is wide_P launches when the sketch is cut into panels
    tm->qr_third_passes = c.n_robust; tm->qr_refills = c.n_refill; tm->jacobi_sweeps = hflags[4]; tm->live_columns = wide ? l : (hflags[1] ? hflags[1] : l);
    tm->jacobi_converged = power_only ? 1 : (hflags[5] > 0 ? 1 : 0);
    tm->total_ms = total.ms();
  }
  return CORRLA_OK;
}

}  // namespace corrla_eng

using namespace corrla_eng;


// ---------------------------------------------------------------------------------------------
// f32 instantiation of the generic reference functions (random_svd<T>, power_iter<T>: T = f32).  The contraction
// hardware of this path is the FP64 tensor pipe, so single-precision data are widened once on the device, the f64
// engine runs unchanged, and the factors are rounded to f32 on the way out: results at least as accurate as an
// all-f32 evaluation (faer's f32 GEMM / QR / SVD) at the cost of an f64 copy of A on the device.
// ---------------------------------------------------------------------------------------------
namespace {

__global__ void __launch_bounds__(256)
widen_kernel(const float* __restrict__ src, int64_t rows, int64_t cols, int64_t rs, int64_t cs, double* __restrict__ dst) {
  const int64_t total = rows * cols;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int64_t i = e / cols, j = e - i * cols;
    dst[e] = (double)src[i * rs + j * cs];
  }
}
__global__ void __launch_bounds__(256)
narrow_kernel(const double* __restrict__ src, int64_t count, float* __restrict__ dst) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += stride) dst[e] = (float)src[e];
}

int f32_impl(const float* a, int64_t nrows, int64_t ncols, int64_t rs, int64_t cs, size_t n_rank, size_t n_iter,
             size_t n_oversamples, const corrla_rsvd_opts* opts_in, float* u, float* s, float* vt, corrla_timings* tm,
             bool power_only, float* q_out) {
  corrla_rsvd_opts o = opts_in ? *opts_in : default_opts();
  if (a == nullptr || nrows <= 0 || ncols <= 0) { set_last_error("empty or null input matrix"); return CORRLA_ERR_INVALID; }
  if (power_only ? (q_out == nullptr) : (u == nullptr || s == nullptr || vt == nullptr)) { set_last_error("null output"); return CORRLA_ERR_INVALID; }
  if (o.comm != nullptr) { set_last_error("the f32 entry points do not take a communicator"); return CORRLA_ERR_UNSUPPORTED; }
  Scope sc;
  ST_TRY(open_scope(&o, &sc));
  corrla_ctx* ctx = sc.ctx;
  cudaStream_t st = sc.st;
  const bool in_dev = o.a_on_device != 0, out_dev = o.out_on_device != 0;
  const size_t k = n_rank;
  const size_t n_a = (size_t)nrows * ncols;
  const size_t n_u = power_only ? (size_t)nrows * k : (size_t)nrows * k, n_v = power_only ? 0 : (size_t)ncols * k;
  double* a64 = static_cast<double*>(ctx->get("f32_a64", n_a * 8));
  double* o64 = static_cast<double*>(ctx->get("f32_o64", (n_u + n_v + k + 8) * 8));
  if (!a64 || !o64) { set_last_error("device allocation failed (f32 widening buffers)"); return CORRLA_ERR_ALLOC; }
  Timer th2d;
  const float* asrc = a;
  int64_t srs = rs, scs = cs;
  if (!in_dev) {
    // host: move the floats as they lie when they are one dense block (either orientation), else pack on the host
    float* a32 = static_cast<float*>(ctx->get("f32_a32", n_a * 4));
    if (!a32) { set_last_error("device allocation failed (f32 staging)"); return CORRLA_ERR_ALLOC; }
    const bool dense_rm = (cs == 1 && rs == ncols), dense_cm = (rs == 1 && cs == nrows);
    if (dense_rm || dense_cm) {
      CU_TRY(cudaMemcpyAsync(a32, a, n_a * 4, cudaMemcpyHostToDevice, st));
    } else {
      std::vector<float> tmp;
      try { tmp.resize(n_a); } catch (...) { set_last_error("host allocation failed"); return CORRLA_ERR_ALLOC; }
      for (int64_t i = 0; i < nrows; ++i)
        for (int64_t j = 0; j < ncols; ++j) tmp[(size_t)i * ncols + j] = a[i * rs + j * cs];
      CU_TRY(cudaMemcpyAsync(a32, tmp.data(), n_a * 4, cudaMemcpyHostToDevice, st));
      CU_TRY(cudaStreamSynchronize(st));
      srs = ncols; scs = 1;
    }
    asrc = a32;
  }
  widen_kernel<<<(unsigned)std::min<int64_t>((int64_t)(n_a + 255) / 256, 148 * 32), 256, 0, st>>>(asrc, nrows, ncols, srs, scs, a64);
  CU_TRY(cudaGetLastError());
  const double h2d_ms = in_dev ? 0.0 : th2d.ms();
  corrla_rsvd_opts oi = o;
  oi.a_on_device = 1; oi.out_on_device = 1; oi.ctx = ctx; oi.stream = st; oi.device = ctx->device;
  double *u64 = o64, *v64 = o64 + n_u, *s64 = o64 + n_u + n_v;
  int status;
  if (power_only) status = rsvd_impl(a64, nrows, ncols, ncols, 1, n_rank, n_iter, 0, &oi, nullptr, nullptr, nullptr, tm, true, u64);
  else status = rsvd_impl(a64, nrows, ncols, ncols, 1, n_rank, n_iter, n_oversamples, &oi, u64, s64, v64, tm, false, nullptr);
  if (status != CORRLA_OK) return status;
  const size_t n_out = power_only ? n_u : n_u + n_v + k;
  float* o32 = out_dev ? nullptr : static_cast<float*>(ctx->get("f32_o32", (n_out + 8) * 4));
  if (!out_dev && !o32) { set_last_error("device allocation failed (f32 outputs)"); return CORRLA_ERR_ALLOC; }
  auto narrow = [&](const double* src, size_t count, float* dst) -> int {
    if (count == 0) return CORRLA_OK;
    narrow_kernel<<<(unsigned)std::min<size_t>((count + 255) / 256, 148 * 32), 256, 0, st>>>(src, (int64_t)count, dst);
    CU_TRY(cudaGetLastError());
    return CORRLA_OK;
  };
  Timer td2h;
  if (out_dev) {
    if (power_only) ST_TRY(narrow(u64, n_u, q_out));
    else { ST_TRY(narrow(u64, n_u, u)); ST_TRY(narrow(v64, n_v, vt)); ST_TRY(narrow(s64, k, s)); }
    CU_TRY(cudaStreamSynchronize(st));
  } else {
    ST_TRY(narrow(o64, n_out, o32));
    if (power_only) CU_TRY(cudaMemcpyAsync(q_out, o32, n_u * 4, cudaMemcpyDeviceToHost, st));
    else {
      CU_TRY(cudaMemcpyAsync(u, o32, n_u * 4, cudaMemcpyDeviceToHost, st));
      CU_TRY(cudaMemcpyAsync(vt, o32 + n_u, n_v * 4, cudaMemcpyDeviceToHost, st));
      CU_TRY(cudaMemcpyAsync(s, o32 + n_u + n_v, k * 4, cudaMemcpyDeviceToHost, st));
    }
    CU_TRY(cudaStreamSynchronize(st));
  }
  if (tm) { tm->h2d_ms = h2d_ms; tm->d2h_ms = out_dev ? 0.0 : td2h.ms(); }
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

int corrla_rsvd_f32(const float* a, int64_t nrows, int64_t ncols, int64_t row_stride, int64_t col_stride, size_t n_rank,
                    size_t n_iter, size_t n_oversamples, const corrla_rsvd_opts* opts, float* u, float* s, float* vt,
                    corrla_timings* timings) {
  try {
    return f32_impl(a, nrows, ncols, row_stride, col_stride, n_rank, n_iter, n_oversamples, opts, u, s, vt, timings, false, nullptr);
  } catch (const std::exception& e) { set_last_error("exception: %s", e.what()); return CORRLA_ERR_ALLOC; }
  catch (...) { set_last_error("unknown exception"); return CORRLA_ERR_INVALID; }
}

int corrla_power_iter_f32(const float* a, int64_t nrows, int64_t ncols, int64_t row_stride, int64_t col_stride,
                          size_t omega_rank, size_t n_iter, const corrla_rsvd_opts* opts, float* q, corrla_timings* timings) {
  try {
    return f32_impl(a, nrows, ncols, row_stride, col_stride, omega_rank, n_iter, 0, opts, nullptr, nullptr, nullptr, timings, true, q);
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
    if (!res || !lhs || !rhs || lhs_rows <= 0 || lhs_cols <= 0 || rhs_cols <= 0 || rhs_cols > (1 << 20)) { set_last_error("bad argument"); return CORRLA_ERR_INVALID; }
    Scope sc;
    ST_TRY(open_scope(opts, &sc));
    Core c; c.ctx = sc.ctx; c.st = sc.st;
    // more than 128 right-hand columns run as column panels of equal width, one launch each
    int P = 1, w = (int)rhs_cols;
    if (rhs_cols > 8 * kMaxNblk) Wide::plan((int)rhs_cols, &P, &w);
    ST_TRY(c.setup_dims(lhs_rows, lhs_cols, w));
    ST_TRY(stage_matrix(sc.ctx, sc.st, "A", lhs, lhs_rows, lhs_cols, lhs_rs, lhs_cs, on_device != 0, &c.av, &c.a_rowmajor, nullptr, &c.launches));
    ST_TRY(c.alloc_workspace(false));
    double* X = static_cast<double*>(sc.ctx->get("Za", (size_t)c.n16 * c.ld * 8));
    const int nc = (int)rhs_cols;
    double* out = on_device ? nullptr : static_cast<double*>(sc.ctx->get("Uout", (size_t)lhs_rows * nc * 8));
    if (!X || (!on_device && !out)) { set_last_error("device allocation failed"); return CORRLA_ERR_ALLOC; }
    for (int p = 0; p < P; ++p) {
      const int c0 = p * w, wc = std::min(w, nc - c0);
      CU_TRY(cudaMemsetAsync(X, 0, (size_t)c.n16 * c.ld * 8, sc.st));
      ST_TRY(pack_small(sc.ctx, sc.st, rhs + (int64_t)c0 * rhs_cs, lhs_cols, wc, rhs_rs, rhs_cs, on_device != 0, X, c.ld, beta, &c.launches));
      if (on_device) ST_TRY(c.mm(c.av, c.a_rowmajor, X, res + (int64_t)c0 * res_cs, res_rs, res_cs, wc));
      else ST_TRY(c.mm(c.av, c.a_rowmajor, X, out + c0, nc, 1, wc));
    }
    if (on_device) {
      CU_TRY(cudaStreamSynchronize(sc.st));
    } else {
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
    if (!a || !q || nrows <= 0 || ncols <= 0 || ncols > 2048) { set_last_error("bad argument"); return CORRLA_ERR_INVALID; }
    Scope sc;
    ST_TRY(open_scope(opts, &sc));
    Core c; c.ctx = sc.ctx; c.st = sc.st; c.comm = opts ? opts->comm : nullptr;
    const bool wide = ncols > 8 * kMaxNblk;
    int P = 1, w = (int)ncols;
    if (wide) Wide::plan((int)ncols, &P, &w);
    ST_TRY(c.setup_dims(nrows, ncols, w));
    c.grows = (opts && opts->global_rows > 0) ? (double)opts->global_rows : (double)nrows;
    ST_TRY(c.alloc_workspace(wide));
    ST_TRY(c.alloc_buffers(wide));
    double* qd = on_device ? q : static_cast<double*>(sc.ctx->get("Uout", (size_t)nrows * ncols * 8));
    if (!qd) { set_last_error("device allocation failed"); return CORRLA_ERR_ALLOC; }
    int hinfo[4] = {0, 0, 0, 0};
    if (wide) {
      Wide wd(c);
      ST_TRY(wd.alloc((int)ncols));
      for (int p = 0; p < P; ++p)
        ST_TRY(pack_small(sc.ctx, sc.st, a + (int64_t)p * w * col_stride, nrows, wd.lp(p), row_stride, col_stride, on_device != 0,
                          wd.Y[p], c.ld, 1.0, &c.launches));
      ST_TRY(wd.block_qr(wd.Y, nrows, c.comm != nullptr, c.grows, false, false));
      ST_TRY(wd.scatter_q(qd));
    } else {
      ST_TRY(pack_small(sc.ctx, sc.st, a, nrows, ncols, row_stride, col_stride, on_device != 0, c.Y, c.ld, 1.0, &c.launches));
      ST_TRY(c.qr_inplace(c.Y, nrows, c.comm != nullptr, c.grows, c.Tf));
      ST_TRY(c.mm(c.view_rows(c.Y, nrows), true, c.Tf, qd, 1, nrows, (int)ncols));
      CU_TRY(cudaMemcpyAsync(hinfo, c.flags + 1, 8, cudaMemcpyDeviceToHost, sc.st));
    }
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
size_t corrla_ctx_trim(corrla_ctx* ctx) {
  if (ctx == nullptr) return 0;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(ctx->device);
  const size_t freed = ctx->trim();
  if (prev >= 0) cudaSetDevice(prev);
  return freed;
}

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
    case CORRLA_ERR_UNSUPPORTED: return "unsupported size or option combination";
    case CORRLA_ERR_ALLOC: return "allocation failed";
    case CORRLA_ERR_COMM: return "communicator (NCCL) error";
    case CORRLA_ERR_NO_DEVICE: return "no CUDA device (no CPU fallback exists)";
    default: return "unknown status";
  }
}
const char* corrla_last_error(void) { return last_error_cstr(); }
const char* corrla_version(void) { return "corrla_b200 0.1.0 (sm_100a, DMMA+TMA)"; }

}  // extern "C"
