// Reduced-order models that sit directly on the RSVD: DMD with control
// (dmd_rom.rs:46-146) and POD modes/weights (pod_rom.rs:53-75).  Everything that touches the tall snapshot matrix
// runs through the same skinny DMMA GEMM as the RSVD passes; the r x r eigendecomposition (DMDc) and the
// n_snap x n_snap RBF interpolation (POD) stay with the caller, as they are negligible and not data-parallel.
#include <string>
#include <vector>

#include "engine_core.cuh"
#include "gradients.cuh"

using namespace corrla_eng;

namespace {

// X[i][j] *= pinv_diag(s)[j] for j < r  (mat_pinv_diag, mat_utils.rs:386-402: 0 if |s| < 1e-20 else 1/(s + 1e-20))
__global__ void __launch_bounds__(256)
scale_cols_pinv_kernel(double* __restrict__ X, int64_t rows, int r, int64_t ld, const double* __restrict__ s) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * r) return;
  const int64_t i = idx / r;
  const int j = (int)(idx - i * r);
  const double sv = s[j];
  const double inv = (sv < 1.0e-20 && sv > -1.0e-20) ? 0.0 : 1.0 / (sv + 1.0e-20);
  X[i * ld + j] *= inv;
}

// X[i][j] -= mu[j]
__global__ void __launch_bounds__(256)
sub_col_means_kernel(double* __restrict__ X, int64_t rows, int cols, int64_t ld, const double* __restrict__ mu) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols) return;
  const int64_t i = idx / cols;
  const int j = (int)(idx - i * cols);
  X[i * ld + j] -= mu[j];
}

// G holds the upper triangle of a symmetric product (pitch ld): S (pitch ld, both triangles) and out (column-major d x d)
// = scale * G, or the correlation matrix G_ij / sqrt(G_ii G_jj) when normalise is set
__global__ void __launch_bounds__(256)
finish_cov_kernel(const double* __restrict__ G, int d, int64_t ld, double scale, int normalise, double* __restrict__ S,
                  double* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= d * d) return;
  const int i = idx / d, j = idx - i * d;
  const double g = (i <= j) ? G[(int64_t)i * ld + j] : G[(int64_t)j * ld + i];
  double v = g * scale;
  if (normalise) {
    const double den = sqrt(G[(int64_t)i * ld + i]) * sqrt(G[(int64_t)j * ld + j]);
    v = den > 0.0 ? g / den : 0.0;
  }
  S[(int64_t)i * ld + j] = v;
  if (out != nullptr) out[(int64_t)j * d + i] = v;
}

struct RomBufs {
  corrla_ctx* ctx;
  cudaStream_t st;
  // zero-filled pool buffer
  double* zeros(const char* name, size_t elems) {
    double* p = static_cast<double*>(ctx->get(name, elems * 8));
    if (p != nullptr && cudaMemsetAsync(p, 0, elems * 8, st) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
  }
  double* raw(const char* name, size_t elems) { return static_cast<double*>(ctx->get(name, elems * 8)); }
};

// Copy one strided block (rows x T) into rows [off, off + rows) of the column-major stack D (pitch ld).
int stage_block(corrla_ctx* ctx, cudaStream_t st, const double* src, int64_t rows, int64_t T, int64_t rs, int64_t cs,
                bool on_device, double* D, int64_t ld, int64_t off, double* h2d_ms, int* launches) {
  if (rows == 0) return CORRLA_OK;
  cudaError_t e;
  if (on_device && rs == 1 && cs >= rows) {
    e = cudaMemcpy2DAsync(D + off, ld * 8, src, cs * 8, rows * 8, T, cudaMemcpyDeviceToDevice, st);   // already column-major
  } else if (on_device) {
    e = repack_launch(src, T, rows, cs, rs, D + off, ld, st);      // D[t*ld + off + i] = src[i*rs + t*cs]
    ++*launches;
  } else if (rs == 1 && cs >= rows) {
    Timer t;
    e = copy_h2d_2d(ctx->bounce, st, D + off, ld * 8, src, cs * 8, rows * 8, T);
    *h2d_ms += t.ms();
  } else {
    MatView v; bool rm = true;
    ST_TRY(stage_matrix(ctx, st, "rom_raw", src, rows, T, rs, cs, false, &v, &rm, h2d_ms, launches));
    e = rm ? repack_launch(v.p, T, rows, 1, v.ld, D + off, ld, st) : repack_launch(v.p, T, rows, v.ld, 1, D + off, ld, st);
    ++*launches;
  }
  if (e != cudaSuccess) { set_last_error("staging of the snapshot matrix failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return CORRLA_ERR_CUDA; }
  return CORRLA_OK;
}

int copy_out(corrla_ctx* ctx, cudaStream_t st, double* dst_host, const double* src_dev, size_t elems) {
  if (dst_host == nullptr || elems == 0) return CORRLA_OK;
  CU_TRY(copy_d2h_2d(ctx->bounce, st, dst_host, elems * 8, src_dev, elems * 8, elems * 8, 1));
  return CORRLA_OK;
}

void add_timings(corrla_timings* acc, const corrla_timings& t) {
  acc->device_ms += t.device_ms; acc->gpu_launches += t.gpu_launches; acc->passes_over_a += t.passes_over_a;
  acc->qr_third_passes += t.qr_third_passes; acc->qr_refills += t.qr_refills; acc->jacobi_sweeps += t.jacobi_sweeps;
  acc->live_columns = t.live_columns; acc->pass_launches += t.pass_launches; acc->pass_ms += t.pass_ms;
  acc->pass_flops = t.pass_flops;
}

int dmdc_impl(const double* x, int64_t n_x, int64_t T, int64_t x_rs, int64_t x_cs, const double* u, int64_t n_u,
              int64_t u_rs, int64_t u_cs, size_t n_modes, size_t n_iters, const corrla_rsvd_opts* opts_in,
              const double* omega_y, double* a_til, double* b, double* modes_scale, double* s_til, double* u_hat,
              corrla_timings* tm) {
  Timer total;
  corrla_rsvd_opts o = opts_in ? *opts_in : default_opts();
  corrla_timings acc;
  memset(&acc, 0, sizeof(acc));
  if (tm) memset(tm, 0, sizeof(*tm));
  if (x == nullptr || n_x <= 0 || T < 2 || n_u < 0 || (n_u > 0 && u == nullptr)) { set_last_error("bad snapshot / control matrix"); return CORRLA_ERR_INVALID; }
  if (o.center != 0) { set_last_error("centring is not part of DMDc"); return CORRLA_ERR_INVALID; }
  // With a communicator every rank passes its block of state rows of x and the whole u; the control rows are stacked
  // under the LAST rank's block, so the global input-space matrix is [x; u] exactly as in the reference.
  corrla_comm* comm = (o.comm != nullptr && o.comm->nranks > 1) ? o.comm : nullptr;
  const int64_t n_u_all = n_u;
  if (comm != nullptr && comm->rank != comm->nranks - 1) n_u = 0;
  const int64_t M = n_x + n_u, nn = T - 1;
  const int r = (int)std::min<size_t>(n_modes, 1 << 20);
  if (r <= 0) { set_last_error("n_modes must be positive"); return CORRLA_ERR_INVALID; }
  // the reference indexes r columns out of l = min(r + 12, thin columns) (random_svd.rs:77,98): it panics beyond
  if (comm == nullptr && (int64_t)r > std::min<int64_t>(n_x, nn)) {
    set_last_error("n_modes=%d exceeds min(n_x, n_snapshots - 1)=%lld (the reference panics here)", r, (long long)std::min<int64_t>(n_x, nn));
    return CORRLA_ERR_RANK;
  }

  Scope sc;
  ST_TRY(open_scope(&o, &sc));
  corrla_ctx* ctx = sc.ctx;
  cudaStream_t st = sc.st;
  RomBufs rb{ctx, st};
  int launches = 0;
  double h2d_ms = 0.0;
  const bool in_dev = o.a_on_device != 0, out_dev = o.out_on_device != 0;

  // ---- the stacked snapshots [x; u], column-major, once (mat_vstack, dmd_rom.rs:66)
  const int64_t ldD = round_up(M, 2);
  double* D = rb.raw("rom_stack", (size_t)T * ldD);
  if (!D) { set_last_error("device allocation for the snapshot stack failed (%lld x %lld)", (long long)M, (long long)T); return CORRLA_ERR_ALLOC; }
  ST_TRY(stage_block(ctx, st, x, n_x, T, x_rs, x_cs, in_dev, D, ldD, 0, &h2d_ms, &launches));
  ST_TRY(stage_block(ctx, st, u, n_u, T, u_rs, u_cs, in_dev, D, ldD, n_x, &h2d_ms, &launches));
  const double* Xv = D;            // _X: all rows, snapshots 0 .. T-2      (:148-153)
  const double* Yv = D + ldD;      // _Y: state rows, snapshots 1 .. T-1    (:156-162)

  // ---- the two RSVDs (:72, :82), results stay on the device
  double* Util = rb.raw("rom_util", (size_t)M * r);
  double* Stil = rb.raw("rom_stil", (size_t)r + 8);
  double* Vttil = rb.raw("rom_vttil", (size_t)r * nn);
  double* Uhat = (out_dev && u_hat) ? u_hat : rb.raw("rom_uhat", (size_t)n_x * r);
  double* Shat = rb.raw("rom_shat", (size_t)r + 8);
  double* Vthat = rb.raw("rom_vthat", (size_t)r * nn);
  if (!Util || !Stil || !Vttil || !Uhat || !Shat || !Vthat) { set_last_error("device allocation failed (DMDc factors)"); return CORRLA_ERR_ALLOC; }
  corrla_rsvd_opts oi = o;
  oi.a_on_device = 1; oi.out_on_device = 1; oi.ctx = ctx; oi.stream = st; oi.device = ctx->device;
  corrla_timings t1, t2;
  if (comm != nullptr && o.global_rows > 0) oi.global_rows = o.global_rows + n_u_all;     // o.global_rows counts state rows
  ST_TRY(rsvd_impl(Xv, M, nn, 1, ldD, (size_t)r, n_iters, 12, &oi, Util, Stil, Vttil, &t1, false, nullptr));
  oi.omega = omega_y;
  oi.seed = o.seed + 1;
  oi.global_rows = o.global_rows;
  ST_TRY(rsvd_impl(Yv, n_x, nn, 1, ldD, (size_t)r, n_iters, 12, &oi, Uhat, Shat, Vthat, &t2, false, nullptr));
  add_timings(&acc, t1);
  add_timings(&acc, t2);

  // ---- products on the tall side.  r columns run as P column panels of padded width w <= 128 (P = 1 up to 128
  // modes): an n x r matrix is P buffers in the engine layout, an r x r matrix one row-major buffer of pitch ldW whose
  // w x w blocks are cut out (or assembled) as the right-hand operands of the panel products.
  int P = 1, w = r;
  panel_plan(r, &P, &w);
  auto lp = [&](int pi) { return pi == P - 1 ? r - (P - 1) * w : w; };
  Core c;
  c.ctx = ctx; c.st = st; c.comm = comm;
  ST_TRY(c.setup_dims(std::max(n_x, nn), std::min(n_x, nn), w));
  ST_TRY(c.alloc_workspace(true));
  const int Lc = c.Lc, ld = c.ld, L16 = c.L16;
  const size_t gx = comm ? (size_t)Lc * ld : 0;            // r x r cross products are summed over the ranks
  const int64_t nx16 = round_up(n_x, 16), nn16 = round_up(nn, 16);
  const int ldW = P * w + 4;
  const int64_t rW16 = round_up((int64_t)P * w, 16);
  auto panels = [&](const char* name, size_t elems, std::vector<double*>* out) -> bool {
    out->assign(P, nullptr);
    for (int pi = 0; pi < P; ++pi) {
      const std::string nm = pi == 0 ? std::string(name) : std::string(name) + "_p" + std::to_string(pi);
      (*out)[pi] = rb.zeros(nm.c_str(), elems);
      if ((*out)[pi] == nullptr) return false;
    }
    return true;
  };
  std::vector<double*> Vs, G2, P1, Uh, U1;
  bool ok = panels("rom_vs", (size_t)nn16 * ld, &Vs) && panels("rom_g2", (size_t)nn16 * ld, &G2) &&
            panels("rom_p1", (size_t)nx16 * ld, &P1) && panels("rom_uh", (size_t)nx16 * ld, &Uh) &&
            panels("rom_u1", (size_t)nx16 * ld, &U1);
  double* T0 = rb.zeros("rom_t0", (size_t)rW16 * ldW);      // tmp_op_scale, r x r
  double* C1 = rb.zeros("rom_c1", (size_t)rW16 * ldW);      // u_til_1^T u_hat, r x r
  double* Blk = rb.zeros("rom_blk", (size_t)L16 * ld);      // one w x w block as a product result / right-hand operand
  double* Bcol = rb.zeros("rom_bcol", (size_t)rW16 * ld);   // one r x w block column as a right-hand operand
  double* U2t = rb.zeros("rom_u2t", (size_t)L16 * ld);
  if (!ok || !T0 || !C1 || !Blk || !Bcol || !U2t) { set_last_error("device allocation failed (DMDc products)"); return CORRLA_ERR_ALLOC; }
  std::vector<double*>& H = P1;   // reused once tmp_op_scale is formed
  cudaError_t e = cudaSuccess;
  auto chk = [&](const char* what) -> int {
    if (e != cudaSuccess) { set_last_error("%s failed: %s", what, cudaGetErrorString(e)); cudaGetLastError(); return CORRLA_ERR_CUDA; }
    return CORRLA_OK;
  };
  // block (pi, pj) of an r x r matrix as a right-hand operand (lp(pi) x lp(pj), engine layout)
  auto load_block = [&](const double* Wd, int pi, int pj) -> int {
    CU_TRY(cudaMemsetAsync(Blk, 0, (size_t)L16 * ld * 8, st));
    e = repack_launch(Wd + (size_t)pi * w * ldW + (size_t)pj * w, lp(pi), lp(pj), ldW, 1, Blk, ld, st);
    ++launches;
    return chk("repack");
  };
  const MatView yv{Yv, n_x, nn, ldD};                       // column-major: inner = state rows
  for (int pi = 0; pi < P; ++pi) {
    // Vs = v_til * pinv(diag(s_til))      v_til(i, j) = Vttil[j + i*r]                     (:74, :86-87)
    e = repack_launch(Vttil + (size_t)pi * w, nn, lp(pi), r, 1, Vs[pi], ld, st); ST_TRY(chk("repack"));
    scale_cols_pinv_kernel<<<(unsigned)((nn * lp(pi) + 255) / 256), 256, 0, st>>>(Vs[pi], nn, lp(pi), ld, Stil + (size_t)pi * w);
    e = cudaGetLastError(); ST_TRY(chk("column scaling"));
    // row-major padded copies of u_hat and u_til_1 (n_x x r)                                (:75-77)
    e = repack_launch(Uhat + (size_t)pi * w * n_x, n_x, lp(pi), 1, n_x, Uh[pi], ld, st); ST_TRY(chk("repack"));
    e = repack_launch(Util + (size_t)pi * w * M, n_x, lp(pi), 1, M, U1[pi], ld, st); ST_TRY(chk("repack"));
    launches += 4;
    // P1 = Y * v_til * s_inv  (one pass over Y per panel)                                   (:90-94)
    ST_TRY(c.mm(yv, false, Vs[pi], P1[pi], ld, 1, Lc, nullptr, nullptr, nullptr, 0, true));
  }
  // tmp_op_scale = u_hat^T * P1;  C1 = u_til_1^T * u_hat   (w x w blocks, summed over the ranks)   (:90-97)
  for (int pi = 0; pi < P; ++pi)
    for (int pj = 0; pj < P; ++pj) {
      ST_TRY(c.mm(c.view_rows(Uh[pi], n_x), false, P1[pj], Blk, ld, 1, Lc, nullptr, nullptr, nullptr, 0, false, nullptr, gx));
      e = repack_launch(Blk, lp(pi), lp(pj), ld, 1, T0 + (size_t)pi * w * ldW + (size_t)pj * w, ldW, st); ST_TRY(chk("repack"));
      ST_TRY(c.mm(c.view_rows(U1[pi], n_x), false, Uh[pj], Blk, ld, 1, Lc, nullptr, nullptr, nullptr, 0, false, nullptr, gx));
      e = repack_launch(Blk, lp(pi), lp(pj), ld, 1, C1 + (size_t)pi * w * ldW + (size_t)pj * w, ldW, st); ST_TRY(chk("repack"));
      launches += 2;
    }
  // a_til = tmp_op_scale * C1: the r x r left operand whole, C1 one block column at a time        (:95-97)
  double* a_dev = (out_dev && a_til) ? a_til : rb.raw("rom_atil", (size_t)r * r + 8);
  if (!a_dev) { set_last_error("device allocation failed"); return CORRLA_ERR_ALLOC; }
  for (int pj = 0; pj < P; ++pj) {
    CU_TRY(cudaMemsetAsync(Bcol, 0, (size_t)rW16 * ld * 8, st));
    e = repack_launch(C1 + (size_t)pj * w, r, lp(pj), ldW, 1, Bcol, ld, st); ST_TRY(chk("repack"));
    ++launches;
    ST_TRY(c.mm(MatView{T0, (int64_t)P * w, (int64_t)r, (int64_t)ldW}, true, Bcol, a_dev + (size_t)pj * w * r, 1, r, lp(pj)));
  }
  // modes_scale = Y * (v_til * s_inv * C1)  (second pass over Y per panel)                  (:133-139)
  double* ms_dev = nullptr;
  if (modes_scale != nullptr) {
    ms_dev = out_dev ? modes_scale : rb.raw("rom_modes", (size_t)n_x * r);
    if (!ms_dev) { set_last_error("device allocation failed"); return CORRLA_ERR_ALLOC; }
    for (int pj = 0; pj < P; ++pj) {
      for (int pk = 0; pk < P; ++pk) {
        ST_TRY(load_block(C1, pk, pj));
        ST_TRY(c.mm(c.view_rows(Vs[pk], nn), true, Blk, G2[pj], ld, 1, Lc, nullptr, nullptr, nullptr, 0, false, nullptr, 0, 0, pk > 0));
      }
      ST_TRY(c.mm(yv, false, G2[pj], ms_dev + (size_t)pj * w * n_x, 1, n_x, lp(pj), nullptr, nullptr, nullptr, 0, true));
    }
  }
  // _B = u_hat * (tmp_op_scale * u_til_2^T) = (u_hat * tmp_op_scale) * u_til_2^T            (:100-106)
  double* b_dev = nullptr;
  if (b != nullptr && n_u_all > 0) {
    b_dev = out_dev ? b : rb.raw("rom_b", (size_t)n_x * n_u_all);
    if (!b_dev) { set_last_error("device allocation failed"); return CORRLA_ERR_ALLOC; }
    for (int pj = 0; pj < P; ++pj)                                                            // H = u_hat * tmp_op_scale
      for (int pk = 0; pk < P; ++pk) {
        ST_TRY(load_block(T0, pk, pj));
        ST_TRY(c.mm(c.view_rows(Uh[pk], n_x), true, Blk, H[pj], ld, 1, Lc, nullptr, nullptr, nullptr, 0, false, nullptr, 0, 0, pk > 0));
      }
    for (int64_t j0 = 0; j0 < n_u_all; j0 += Lc) {
      const int wu = (int)std::min<int64_t>(Lc, n_u_all - j0);
      for (int pk = 0; pk < P; ++pk) {
        CU_TRY(cudaMemsetAsync(U2t, 0, (size_t)L16 * ld * 8, st));
        if (n_u > 0) {                                                                        // the rank that holds u_til_2
          e = repack_launch(Util + n_x + j0 + (size_t)pk * w * M, lp(pk), wu, M, 1, U2t, ld, st); ST_TRY(chk("repack"));   // u_til_2^T block
          ++launches;
        }
        if (comm != nullptr) ST_TRY(c.allreduce(U2t, (size_t)L16 * ld));                      // zeros elsewhere: a broadcast
        ST_TRY(c.mm(c.view_rows(H[pk], n_x), true, U2t, b_dev + (size_t)j0 * n_x, 1, n_x, wu, nullptr, nullptr, nullptr, 0, false,
                    nullptr, 0, 0, pk > 0));
      }
    }
  }
  launches += c.launches;

  // ---- results
  if (out_dev) {
    if (s_til) CU_TRY(cudaMemcpyAsync(s_til, Stil, (size_t)r * 8, cudaMemcpyDeviceToDevice, st));
    if (tm) CU_TRY(cudaStreamSynchronize(st));
  } else {
    CU_TRY(cudaStreamSynchronize(st));
    Timer t;
    if (a_til) CU_TRY(cudaMemcpy(a_til, a_dev, (size_t)r * r * 8, cudaMemcpyDeviceToHost));
    if (s_til) CU_TRY(cudaMemcpy(s_til, Stil, (size_t)r * 8, cudaMemcpyDeviceToHost));
    ST_TRY(copy_out(ctx, st, b_dev ? b : nullptr, b_dev, (size_t)n_x * n_u_all));
    ST_TRY(copy_out(ctx, st, modes_scale, ms_dev, (size_t)n_x * r));
    ST_TRY(copy_out(ctx, st, u_hat, Uhat, (size_t)n_x * r));
    acc.d2h_ms = t.ms();
  }
  if (tm) {
    *tm = acc;
    tm->h2d_ms = h2d_ms;
    tm->gpu_launches += launches;
    tm->passes_over_a += 2;                 // the two extra passes over Y
    tm->total_ms = total.ms();
  }
  return CORRLA_OK;
}

int pod_impl(const double* x, int64_t n_snap, int64_t n_points, int64_t rs, int64_t cs, size_t n_modes,
             const corrla_rsvd_opts* opts_in, double* modes, double* weights, double* s, corrla_timings* tm) {
  Timer total;
  corrla_rsvd_opts o = opts_in ? *opts_in : default_opts();
  if (tm) memset(tm, 0, sizeof(*tm));
  if (x == nullptr || n_snap <= 0 || n_points <= 0) { set_last_error("empty or null snapshot matrix"); return CORRLA_ERR_INVALID; }
  if (o.center != 0) { set_last_error("centring is not part of PodI"); return CORRLA_ERR_INVALID; }
  // With a communicator every rank passes its block of POINTS (columns of x): n_points is the local count, the modes
  // come back for those points, the weights (n_snap x r) are summed over the ranks and replicated.
  corrla_comm* comm = (o.comm != nullptr && o.comm->nranks > 1) ? o.comm : nullptr;
  const int r = (int)std::min<size_t>(n_modes, 1 << 20);
  const int64_t thin_cols = std::min(n_snap, n_points);
  if (r <= 0) { set_last_error("n_modes must be positive"); return CORRLA_ERR_INVALID; }
  if ((int64_t)r > (comm ? n_snap : thin_cols)) {
    set_last_error("n_modes=%d exceeds min(n_snapshots, n_points)=%lld (the reference panics here)", r, (long long)thin_cols);
    return CORRLA_ERR_RANK;
  }
  Scope sc;
  ST_TRY(open_scope(&o, &sc));
  corrla_ctx* ctx = sc.ctx;
  cudaStream_t st = sc.st;
  RomBufs rb{ctx, st};
  int launches = 0;
  double h2d_ms = 0.0;
  const bool out_dev = o.out_on_device != 0;

  MatView xv; bool rm = true;
  ST_TRY(stage_matrix(ctx, st, "rom_stack", x, n_snap, n_points, rs, cs, o.a_on_device != 0, &xv, &rm, &h2d_ms, &launches));
  const int64_t drs = rm ? xv.ld : 1, dcs = rm ? 1 : xv.ld;

  double* Sd = rb.raw("rom_stil", (size_t)r + 8);
  double* Vt = rb.raw("rom_vttil", (size_t)r * n_points);
  if (!Sd || !Vt) { set_last_error("device allocation failed (POD factors)"); return CORRLA_ERR_ALLOC; }
  corrla_rsvd_opts oi = o;
  oi.a_on_device = 1; oi.out_on_device = 1; oi.ctx = ctx; oi.stream = st; oi.device = ctx->device;
  corrla_timings t1;
  double* m_col = nullptr;     // communicator path: the thin matrix is x_local^T, its left vectors ARE the local modes
  if (comm == nullptr) {
    ST_TRY(rsvd_impl(xv.p, n_snap, n_points, drs, dcs, (size_t)r, 10, 10, &oi, nullptr, Sd, Vt, &t1, false, nullptr, true));   // pod_rom.rs:56
  } else {
    m_col = (out_dev && modes) ? modes : rb.raw("rom_modes", (size_t)n_points * r);
    double* vt_small = rb.raw("rom_vthat", (size_t)r * n_snap);
    if (!m_col || !vt_small) { set_last_error("device allocation failed (POD factors)"); return CORRLA_ERR_ALLOC; }
    ST_TRY(rsvd_impl(xv.p, n_points, n_snap, dcs, drs, (size_t)r, 10, 10, &oi, m_col, Sd, vt_small, &t1, false, nullptr));
  }

  // modes and weights in P column panels of padded width w <= 128 (P = 1 up to 128 modes)
  int P = 1, w = r;
  panel_plan(r, &P, &w);
  auto lp = [&](int pi) { return pi == P - 1 ? r - (P - 1) * w : w; };
  Core c;
  c.ctx = ctx; c.st = st; c.comm = comm;
  ST_TRY(c.setup_dims(std::max(n_snap, n_points), thin_cols, w));
  ST_TRY(c.alloc_workspace(true));
  const int64_t np16 = round_up(n_points, 16);
  double* Mp = rb.raw("rom_uh", (size_t)np16 * c.ld);
  double* m_dev = m_col ? m_col : ((out_dev && modes) ? modes : rb.raw("rom_modes", (size_t)n_points * r));
  double* w_dev = (out_dev && weights) ? weights : rb.raw("rom_b", (size_t)n_snap * r);
  if (!Mp || !m_dev || !w_dev) { set_last_error("device allocation failed (POD products)"); return CORRLA_ERR_ALLOC; }
  const int64_t ns16 = round_up(n_snap, 16);
  double* Wt = nullptr;
  if (weights != nullptr && comm != nullptr) {
    Wt = rb.raw("rom_p1", (size_t)ns16 * c.ld);
    if (!Wt) { set_last_error("device allocation failed (POD weights)"); return CORRLA_ERR_ALLOC; }
  }
  for (int pi = 0; pi < P; ++pi) {
    cudaError_t e;
    CU_TRY(cudaMemsetAsync(Mp, 0, (size_t)np16 * c.ld * 8, st));
    if (comm == nullptr) {
      // modes(i, j) = Vt[j + i*r]  (v.transpose().to_owned(), :57)
      e = repack_launch(Vt + (size_t)pi * w, n_points, lp(pi), r, 1, Mp, c.ld, st);
      if (e == cudaSuccess && modes != nullptr) e = scatter_launch(Mp, n_points, lp(pi), c.ld, m_dev + (size_t)pi * w * n_points, 1, n_points, st);
    } else {
      e = repack_launch(m_col + (size_t)pi * w * n_points, n_points, lp(pi), 1, n_points, Mp, c.ld, st);
    }
    launches += 2;
    if (e != cudaSuccess) { set_last_error("repack failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return CORRLA_ERR_CUDA; }
    // weights = x * modes: one more pass over the snapshots per panel (:61-75); summed over the point blocks of the ranks
    if (weights != nullptr && comm == nullptr)
      ST_TRY(c.mm(xv, rm, Mp, w_dev + (size_t)pi * w * n_snap, 1, n_snap, lp(pi), nullptr, nullptr, nullptr, 0, true));
    if (weights != nullptr && comm != nullptr) {
      CU_TRY(cudaMemsetAsync(Wt, 0, (size_t)ns16 * c.ld * 8, st));
      ST_TRY(c.mm(xv, rm, Mp, Wt, c.ld, 1, c.Lc, nullptr, nullptr, nullptr, 0, true, nullptr, (size_t)ns16 * c.ld));
      e = scatter_launch(Wt, n_snap, lp(pi), c.ld, w_dev + (size_t)pi * w * n_snap, 1, n_snap, st);
      ++launches;
      if (e != cudaSuccess) { set_last_error("scatter failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return CORRLA_ERR_CUDA; }
    }
  }
  launches += c.launches;

  double d2h_ms = 0.0;
  if (out_dev) {
    if (s) CU_TRY(cudaMemcpyAsync(s, Sd, (size_t)r * 8, cudaMemcpyDeviceToDevice, st));
    if (tm) CU_TRY(cudaStreamSynchronize(st));
  } else {
    CU_TRY(cudaStreamSynchronize(st));
    Timer t;
    if (s) CU_TRY(cudaMemcpy(s, Sd, (size_t)r * 8, cudaMemcpyDeviceToHost));
    ST_TRY(copy_out(ctx, st, modes, m_dev, (size_t)n_points * r));
    ST_TRY(copy_out(ctx, st, weights, w_dev, (size_t)n_snap * r));
    d2h_ms = t.ms();
  }
  if (tm) {
    *tm = t1;
    tm->h2d_ms = h2d_ms; tm->d2h_ms = d2h_ms;
    tm->gpu_launches += launches;
    tm->passes_over_a += 1;
    tm->total_ms = total.ms();
  }
  return CORRLA_OK;
}

int cov_impl(const double* x, int64_t nrows, int64_t ncols, int64_t rs, int64_t cs, int kind, double scale,
             const corrla_rsvd_opts* opts_in, double* out, double* means, double* evals, double* evecs) {
  corrla_rsvd_opts o = opts_in ? *opts_in : default_opts();
  if (x == nullptr || out == nullptr || nrows <= 0 || ncols <= 0 || kind < 0 || kind > 2 || ((evals == nullptr) != (evecs == nullptr))) {
    set_last_error("bad argument");
    return CORRLA_ERR_INVALID;
  }
  const int d = (int)std::min<int64_t>(ncols, 1 << 20);
  Scope sc;
  ST_TRY(open_scope(&o, &sc));
  corrla_ctx* ctx = sc.ctx;
  cudaStream_t st = sc.st;
  RomBufs rb{ctx, st};
  Core c;
  c.ctx = ctx; c.st = st; c.comm = (o.comm != nullptr && o.comm->nranks > 1) ? o.comm : nullptr;
  ST_TRY(c.setup_dims(nrows, ncols, d));                       // d > 128: CORRLA_ERR_UNSUPPORTED
  ST_TRY(c.alloc_workspace(false));
  const int Lc = c.Lc, ld = c.ld, L16 = c.L16;
  const bool out_dev = o.out_on_device != 0;
  int launches = 0;

  // samples in the engine's padded row-major layout
  MatView xv; bool rm = true;
  ST_TRY(stage_matrix(ctx, st, "A", x, nrows, ncols, rs, cs, o.a_on_device != 0, &xv, &rm, nullptr, &launches));
  double* Xp = rb.zeros("cov_xp", (size_t)c.m16 * ld);
  double* G = rb.zeros("cov_g", (size_t)L16 * ld);
  double* S = rb.zeros("cov_s", (size_t)L16 * ld);
  double* mu = rb.zeros("cov_mu", (size_t)Lc + 8);
  double* part = rb.raw("cov_part", (size_t)sum_blocks(nrows) * Lc + 128);
  double* out_dev_buf = out_dev ? out : rb.raw("cov_out", (size_t)d * d);
  if (!Xp || !G || !S || !mu || !part || !out_dev_buf) { set_last_error("device allocation failed (covariance)"); return CORRLA_ERR_ALLOC; }
  cudaError_t e = rm ? repack_launch(xv.p, nrows, d, xv.ld, 1, Xp, ld, st) : repack_launch(xv.p, nrows, d, 1, xv.ld, Xp, ld, st);
  if (e != cudaSuccess) { set_last_error("repack failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return CORRLA_ERR_CUDA; }

  double n_total = (double)nrows;
  if (c.comm != nullptr) {
    if (o.global_rows > 0) n_total = (double)o.global_rows;
    else {
      CU_TRY(cudaMemcpyAsync(mu + Lc, &n_total, 8, cudaMemcpyHostToDevice, st));
      ST_TRY(c.allreduce(mu + Lc, 1));
      CU_TRY(cudaMemcpyAsync(&n_total, mu + Lc, 8, cudaMemcpyDeviceToHost, st));
      CU_TRY(cudaStreamSynchronize(st));
    }
  }
  if (kind != CORRLA_COV_GRAM) {
    // column means (mat_mean axis 1), then the explicit centred copy the reference also takes (center_mat_col)
    e = sum_over_outer_launch(Xp, Lc, nrows, ld, part, mu, st);
    if (e != cudaSuccess) { set_last_error("mean launch failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return CORRLA_ERR_CUDA; }
    ST_TRY(c.allreduce(mu, (size_t)Lc));
    e = scale_vec_launch(mu, Lc, 1.0 / n_total, st);
    if (e == cudaSuccess) {
      sub_col_means_kernel<<<(unsigned)((nrows * d + 255) / 256), 256, 0, st>>>(Xp, nrows, d, ld, mu);
      e = cudaGetLastError();
    }
    if (e != cudaSuccess) { set_last_error("centring launch failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return CORRLA_ERR_CUDA; }
    if (n_total < 2.0) { set_last_error("covariance needs at least two samples"); return CORRLA_ERR_INVALID; }
    scale = 1.0 / (n_total - 1.0);
  }
  // upper triangle of X^T X (symmetric-output mode of the GEMM), summed over the ranks
  ST_TRY(c.mm(c.view_rows(Xp, nrows), false, Xp, G, ld, 1, Lc, nullptr, nullptr, nullptr, 0, false, nullptr,
              c.comm ? (size_t)Lc * ld : 0, 0, false, 1));
  finish_cov_kernel<<<(unsigned)((d * d + 255) / 256), 256, 0, st>>>(G, d, ld, scale, kind == CORRLA_COV_PEARSON ? 1 : 0, S, out_dev_buf);
  CU_TRY(cudaGetLastError());

  double *sig = nullptr, *Ur = nullptr;
  int* jinfo = nullptr;
  if (evals != nullptr) {
    // symmetric positive semi-definite: the SVD is the eigendecomposition, already sorted in descending order
    sig = rb.zeros("cov_sig", (size_t)L16);
    Ur = rb.zeros("cov_ur", (size_t)L16 * ld);
    double* Vr = rb.zeros("cov_vr", (size_t)L16 * ld);
    double* js = rb.zeros("cov_js", 2 * (size_t)d * (d + 2) + 8);
    int* info = reinterpret_cast<int*>(rb.zeros("cov_info", 8));
    if (!sig || !Ur || !Vr || !js || !info) { set_last_error("device allocation failed (eigendecomposition)"); return CORRLA_ERR_ALLOC; }
    e = jacobi_svd_launch(S, ld, d, sig, Vr, Ur, L16, ld, js, info, st);
    if (e != cudaSuccess) { set_last_error("jacobi launch failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return CORRLA_ERR_CUDA; }
    jinfo = info;
  }
  if (out_dev) {
    if (means && kind != CORRLA_COV_GRAM) CU_TRY(cudaMemcpyAsync(means, mu, (size_t)d * 8, cudaMemcpyDeviceToDevice, st));
    if (evals) {
      CU_TRY(cudaMemcpyAsync(evals, sig, (size_t)d * 8, cudaMemcpyDeviceToDevice, st));
      e = scatter_launch(Ur, d, d, ld, evecs, 1, d, st);
      if (e != cudaSuccess) { set_last_error("scatter failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return CORRLA_ERR_CUDA; }
    }
    CU_TRY(cudaStreamSynchronize(st));      // small results: return them complete (the context stream is not the caller's)
  } else {
    if (evals) {
      double* ev_cm = rb.raw("cov_evcm", (size_t)d * d);
      if (!ev_cm) { set_last_error("device allocation failed"); return CORRLA_ERR_ALLOC; }
      e = scatter_launch(Ur, d, d, ld, ev_cm, 1, d, st);
      if (e != cudaSuccess) { set_last_error("scatter failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return CORRLA_ERR_CUDA; }
      CU_TRY(cudaMemcpyAsync(evecs, ev_cm, (size_t)d * d * 8, cudaMemcpyDeviceToHost, st));
      CU_TRY(cudaMemcpyAsync(evals, sig, (size_t)d * 8, cudaMemcpyDeviceToHost, st));
    }
    CU_TRY(cudaMemcpyAsync(out, out_dev_buf, (size_t)d * d * 8, cudaMemcpyDeviceToHost, st));
    if (means && kind != CORRLA_COV_GRAM) CU_TRY(cudaMemcpyAsync(means, mu, (size_t)d * 8, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
  }
  if (jinfo != nullptr) {
    int hj[2] = {0, 0};
    CU_TRY(cudaMemcpy(hj, jinfo, sizeof(hj), cudaMemcpyDeviceToHost));
    if (hj[1] < 0) { set_last_error("the Jacobi kernel's cluster exchange timed out (results are invalid)"); return CORRLA_ERR_CUDA; }
  }
  return CORRLA_OK;
}

int active_ss_impl(const double* x, int64_t n, int64_t nfeat, int64_t x_rs, int64_t x_cs, const double* y, int64_t y_stride,
                   int order, int n_nbr, const corrla_rsvd_opts* opts_in, double* evals, double* evecs, double* grad_mat,
                   int* n_deficient) {
  corrla_rsvd_opts o = opts_in ? *opts_in : default_opts();
  if (x == nullptr || y == nullptr || evals == nullptr || evecs == nullptr || n <= 0 || nfeat <= 0) { set_last_error("bad argument"); return CORRLA_ERR_INVALID; }
  if (o.comm != nullptr) { set_last_error("corrla_active_ss_f64 does not take a communicator"); return CORRLA_ERR_UNSUPPORTED; }
  if (order != 1 && order != 2) { set_last_error("Not implemented est order: %d (the reference panics here)", order); return CORRLA_ERR_INVALID; }
  const int d = (int)std::min<int64_t>(nfeat, 1 << 20);
  // the reference's asserts (active_subspaces.rs:115-116, :127-128)
  const int64_t need = (order == 1) ? (int64_t)d + 1 : (int64_t)d * (d + 3) / 2;
  if (!(n > need) || !(n_nbr > need)) {
    set_last_error("order %d fit in %d dimensions needs more than %lld samples and neighbours (got %lld, %d)", order, d,
                   (long long)need, (long long)n, n_nbr);
    return CORRLA_ERR_INVALID;
  }
  const int k = (int)std::min<int64_t>(n_nbr, n);                 // KdTree::nearest returns at most all points
  if (k > kKnnMaxK || poly_grad_num_coef(d, order) > kGradMaxCoef || poly_grad_smem_bytes(d, k, order) > 200 * 1024) {
    set_last_error("active subspace fit of order %d with %d features and %d neighbours exceeds the kernel limits "
                   "(<= %d neighbours, <= %d coefficients)", order, d, k, kKnnMaxK, kGradMaxCoef);
    return CORRLA_ERR_UNSUPPORTED;
  }
  Scope sc;
  ST_TRY(open_scope(&o, &sc));
  corrla_ctx* ctx = sc.ctx;
  cudaStream_t st = sc.st;
  RomBufs rb{ctx, st};
  Core c;
  c.ctx = ctx; c.st = st;
  ST_TRY(c.setup_dims(n, nfeat, d));
  ST_TRY(c.alloc_workspace(false));
  const int Lc = c.Lc, ld = c.ld, L16 = c.L16;
  const bool in_dev = o.a_on_device != 0, out_dev = o.out_on_device != 0;
  int launches = 0;

  MatView xv; bool rm = true;
  ST_TRY(stage_matrix(ctx, st, "A", x, n, nfeat, x_rs, x_cs, in_dev, &xv, &rm, nullptr, &launches));
  double* Xp = rb.zeros("cov_xp", (size_t)c.m16 * ld);
  double* Gp = rb.zeros("as_grad", (size_t)c.m16 * ld);
  double* yd = rb.raw("as_y", (size_t)n + 8);
  int* idx = reinterpret_cast<int*>(ctx->get("as_idx", (size_t)n * k * sizeof(int)));
  double* Gm = rb.zeros("cov_g", (size_t)L16 * ld);
  double* S = rb.zeros("cov_s", (size_t)L16 * ld);
  double* sig = rb.zeros("cov_sig", (size_t)L16);
  double* Ur = rb.zeros("cov_ur", (size_t)L16 * ld);
  double* Vr = rb.zeros("cov_vr", (size_t)L16 * ld);
  double* js = rb.zeros("cov_js", 2 * (size_t)d * (d + 2) + 8);
  int* info = reinterpret_cast<int*>(rb.zeros("cov_info", 8));
  if (!Xp || !Gp || !yd || !idx || !Gm || !S || !sig || !Ur || !Vr || !js || !info) { set_last_error("device allocation failed (active subspace)"); return CORRLA_ERR_ALLOC; }
  cudaError_t e = rm ? repack_launch(xv.p, n, d, xv.ld, 1, Xp, ld, st) : repack_launch(xv.p, n, d, 1, xv.ld, Xp, ld, st);
  if (e != cudaSuccess) { set_last_error("repack failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return CORRLA_ERR_CUDA; }
  ST_TRY(pack_small(ctx, st, y, n, 1, y_stride, 1, in_dev, yd, 1, 1.0, &launches));

  // neighbours, local fits, gradient matrix (row i = gradient at sample i)
  const size_t knn_bytes = knn_scratch_bytes(n, k, d);
  void* knn_scratch = ctx->get("as_knn", knn_bytes);       // optional: without it the exact kernel does the whole search
  int knn_fallback = 0;
  e = knn_launch(Xp, n, d, ld, k, idx, knn_scratch, knn_scratch ? knn_bytes : 0, &knn_fallback, st);
  if (e == cudaSuccess) e = poly_grad_launch(Xp, yd, n, d, ld, idx, k, order, Gp, ld, info + 2, st);
  if (e != cudaSuccess) { set_last_error("gradient kernels failed to launch: %s", cudaGetErrorString(e)); cudaGetLastError(); return CORRLA_ERR_CUDA; }
  // grad_mat grad_mat^T / N (:253) on the symmetric-output GEMM, sorted eigendecomposition (:259-271) by Jacobi
  ST_TRY(c.mm(c.view_rows(Gp, n), false, Gp, Gm, ld, 1, Lc, nullptr, nullptr, nullptr, 0, false, nullptr, 0, 0, false, 1));
  finish_cov_kernel<<<(unsigned)((d * d + 255) / 256), 256, 0, st>>>(Gm, d, ld, 1.0 / (double)n, 0, S, nullptr);
  CU_TRY(cudaGetLastError());
  e = jacobi_svd_launch(S, ld, d, sig, Vr, Ur, L16, ld, js, info, st);
  if (e != cudaSuccess) { set_last_error("jacobi launch failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return CORRLA_ERR_CUDA; }

  double* ev_cm = out_dev ? evecs : rb.raw("cov_evcm", (size_t)d * d);
  double* g_cm = grad_mat == nullptr ? nullptr : (out_dev ? grad_mat : rb.raw("as_gout", (size_t)n * d));
  if (!ev_cm || (grad_mat != nullptr && !g_cm)) { set_last_error("device allocation failed"); return CORRLA_ERR_ALLOC; }
  e = scatter_launch(Ur, d, d, ld, ev_cm, 1, d, st);
  if (e == cudaSuccess && g_cm != nullptr) e = scatter_launch(Gp, n, d, ld, g_cm, d, 1, st);   // k x N column-major
  if (e != cudaSuccess) { set_last_error("scatter failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return CORRLA_ERR_CUDA; }
  int hinfo[4] = {0, 0, 0, 0};
  if (out_dev) {
    CU_TRY(cudaMemcpyAsync(evals, sig, (size_t)d * 8, cudaMemcpyDeviceToDevice, st));
  } else {
    CU_TRY(cudaMemcpyAsync(evals, sig, (size_t)d * 8, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(evecs, ev_cm, (size_t)d * d * 8, cudaMemcpyDeviceToHost, st));
  }
  CU_TRY(cudaMemcpyAsync(hinfo, info, sizeof(hinfo), cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  if (hinfo[1] < 0) { set_last_error("the Jacobi kernel's cluster exchange timed out (results are invalid)"); return CORRLA_ERR_CUDA; }
  if (!out_dev && grad_mat != nullptr) ST_TRY(copy_out(ctx, st, grad_mat, g_cm, (size_t)n * d));
  if (n_deficient) *n_deficient = hinfo[2];
  return CORRLA_OK;
}


// PolyGradientEstimator::grad_at for a batch of evaluation points (active_subspaces.rs:66-141): neighbours of every
// query among the samples, one local fit per query, gradient taken at the query.
int poly_grad_at_impl(const double* x, int64_t n, int64_t nfeat, int64_t x_rs, int64_t x_cs, const double* y, int64_t y_stride,
                      int order, int n_nbr, const double* xq, int64_t nq, int64_t q_rs, int64_t q_cs,
                      const corrla_rsvd_opts* opts_in, double* grad_out, int* n_deficient) {
  corrla_rsvd_opts o = opts_in ? *opts_in : default_opts();
  if (x == nullptr || y == nullptr || xq == nullptr || grad_out == nullptr || n <= 0 || nfeat <= 0 || nq <= 0) { set_last_error("bad argument"); return CORRLA_ERR_INVALID; }
  if (o.comm != nullptr) { set_last_error("corrla_poly_grad_at_f64 does not take a communicator"); return CORRLA_ERR_UNSUPPORTED; }
  if (order != 1 && order != 2) { set_last_error("Not implemented est order: %d (the reference panics here)", order); return CORRLA_ERR_INVALID; }
  const int d = (int)std::min<int64_t>(nfeat, 1 << 20);
  const int64_t need = (order == 1) ? (int64_t)d + 1 : (int64_t)d * (d + 3) / 2;       // :115-116, :127-128
  if (!(n > need) || !(n_nbr > need)) {
    set_last_error("order %d fit in %d dimensions needs more than %lld samples and neighbours (got %lld, %d)", order, d,
                   (long long)need, (long long)n, n_nbr);
    return CORRLA_ERR_INVALID;
  }
  const int k = (int)std::min<int64_t>(n_nbr, n);
  if (k > kKnnMaxK || poly_grad_num_coef(d, order) > kGradMaxCoef || poly_grad_smem_bytes(d, k, order) > 200 * 1024) {
    set_last_error("gradient fit of order %d with %d features and %d neighbours exceeds the kernel limits "
                   "(<= %d neighbours, <= %d coefficients)", order, d, k, kKnnMaxK, kGradMaxCoef);
    return CORRLA_ERR_UNSUPPORTED;
  }
  Scope sc;
  ST_TRY(open_scope(&o, &sc));
  corrla_ctx* ctx = sc.ctx;
  cudaStream_t st = sc.st;
  RomBufs rb{ctx, st};
  Core c;
  c.ctx = ctx; c.st = st;
  ST_TRY(c.setup_dims(n, nfeat, d));
  const int ld = c.ld;
  const bool in_dev = o.a_on_device != 0, out_dev = o.out_on_device != 0;
  int launches = 0;
  const int64_t nq16 = (nq + 15) / 16 * 16;

  MatView xv, qv; bool rm = true, qrm = true;
  ST_TRY(stage_matrix(ctx, st, "A", x, n, nfeat, x_rs, x_cs, in_dev, &xv, &rm, nullptr, &launches));
  ST_TRY(stage_matrix(ctx, st, "as_qraw", xq, nq, nfeat, q_rs, q_cs, in_dev, &qv, &qrm, nullptr, &launches));
  double* Xp = rb.zeros("cov_xp", (size_t)c.m16 * ld);
  double* Qp = rb.zeros("as_qp", (size_t)nq16 * ld);
  double* Gp = rb.zeros("as_qgrad", (size_t)nq16 * ld);
  double* yd = rb.raw("as_y", (size_t)n + 8);
  int* idx = reinterpret_cast<int*>(ctx->get("as_qidx", (size_t)nq * k * sizeof(int)));
  int* info = reinterpret_cast<int*>(rb.zeros("cov_info", 8));
  double* g_rm = out_dev ? grad_out : rb.raw("as_qgout", (size_t)nq * d);
  if (!Xp || !Qp || !Gp || !yd || !idx || !info || !g_rm) { set_last_error("device allocation failed (gradient estimator)"); return CORRLA_ERR_ALLOC; }
  cudaError_t e = rm ? repack_launch(xv.p, n, d, xv.ld, 1, Xp, ld, st) : repack_launch(xv.p, n, d, 1, xv.ld, Xp, ld, st);
  if (e == cudaSuccess) e = qrm ? repack_launch(qv.p, nq, d, qv.ld, 1, Qp, ld, st) : repack_launch(qv.p, nq, d, 1, qv.ld, Qp, ld, st);
  if (e != cudaSuccess) { set_last_error("repack failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return CORRLA_ERR_CUDA; }
  ST_TRY(pack_small(ctx, st, y, n, 1, y_stride, 1, in_dev, yd, 1, 1.0, &launches));
  e = knn_query_launch(Xp, n, d, ld, Qp, nq, ld, k, idx, st);
  if (e == cudaSuccess) e = poly_grad_launch(Xp, yd, nq, d, ld, idx, k, order, Gp, ld, info + 2, st, Qp, ld);
  if (e == cudaSuccess) e = scatter_launch(Gp, nq, d, ld, g_rm, d, 1, st);              // n_query x n_features row-major
  if (e != cudaSuccess) { set_last_error("gradient kernels failed to launch: %s", cudaGetErrorString(e)); cudaGetLastError(); return CORRLA_ERR_CUDA; }
  int hinfo[4] = {0, 0, 0, 0};
  CU_TRY(cudaMemcpyAsync(hinfo, info, sizeof(hinfo), cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  if (!out_dev) ST_TRY(copy_out(ctx, st, grad_out, g_rm, (size_t)nq * d));
  if (n_deficient) *n_deficient = hinfo[2];
  return CORRLA_OK;
}
}  // namespace

extern "C" {

int corrla_active_ss_f64(const double* x, int64_t n_samples, int64_t n_features, int64_t x_rs, int64_t x_cs, const double* y,
                         int64_t y_stride, int order, int n_nbr, const corrla_rsvd_opts* opts, double* evals, double* evecs,
                         double* grad_mat, int* n_deficient) {
  try {
    return active_ss_impl(x, n_samples, n_features, x_rs, x_cs, y, y_stride, order, n_nbr, opts, evals, evecs, grad_mat, n_deficient);
  } catch (const std::exception& e) { set_last_error("exception: %s", e.what()); return CORRLA_ERR_ALLOC; }
  catch (...) { set_last_error("unknown exception"); return CORRLA_ERR_INVALID; }
}

int corrla_poly_grad_at_f64(const double* x, int64_t n_samples, int64_t n_features, int64_t x_rs, int64_t x_cs, const double* y,
                            int64_t y_stride, int order, int n_nbr, const double* xq, int64_t n_query, int64_t q_rs, int64_t q_cs,
                            const corrla_rsvd_opts* opts, double* grad_out, int* n_deficient) {
  try {
    return poly_grad_at_impl(x, n_samples, n_features, x_rs, x_cs, y, y_stride, order, n_nbr, xq, n_query, q_rs, q_cs, opts,
                             grad_out, n_deficient);
  } catch (const std::exception& e) { set_last_error("exception: %s", e.what()); return CORRLA_ERR_ALLOC; }
  catch (...) { set_last_error("unknown exception"); return CORRLA_ERR_INVALID; }
}

int corrla_cov_f64(const double* x, int64_t nrows, int64_t ncols, int64_t row_stride, int64_t col_stride, int kind,
                   double scale, const corrla_rsvd_opts* opts, double* out, double* means, double* evals, double* evecs) {
  try {
    return cov_impl(x, nrows, ncols, row_stride, col_stride, kind, scale, opts, out, means, evals, evecs);
  } catch (const std::exception& e) { set_last_error("exception: %s", e.what()); return CORRLA_ERR_ALLOC; }
  catch (...) { set_last_error("unknown exception"); return CORRLA_ERR_INVALID; }
}

int corrla_dmdc_f64(const double* x, int64_t n_x, int64_t n_snap, int64_t x_rs, int64_t x_cs, const double* u, int64_t n_u,
                    int64_t u_rs, int64_t u_cs, size_t n_modes, size_t n_iters, const corrla_rsvd_opts* opts,
                    const double* omega_y, double* a_til, double* b, double* modes_scale, double* s_til, double* u_hat,
                    corrla_timings* timings) {
  try {
    return dmdc_impl(x, n_x, n_snap, x_rs, x_cs, u, n_u, u_rs, u_cs, n_modes, n_iters, opts, omega_y, a_til, b,
                     modes_scale, s_til, u_hat, timings);
  } catch (const std::exception& e) { set_last_error("exception: %s", e.what()); return CORRLA_ERR_ALLOC; }
  catch (...) { set_last_error("unknown exception"); return CORRLA_ERR_INVALID; }
}

int corrla_pod_f64(const double* x, int64_t n_snap, int64_t n_points, int64_t row_stride, int64_t col_stride,
                   size_t n_modes, const corrla_rsvd_opts* opts, double* modes, double* weights, double* s,
                   corrla_timings* timings) {
  try {
    return pod_impl(x, n_snap, n_points, row_stride, col_stride, n_modes, opts, modes, weights, s, timings);
  } catch (const std::exception& e) { set_last_error("exception: %s", e.what()); return CORRLA_ERR_ALLOC; }
  catch (...) { set_last_error("unknown exception"); return CORRLA_ERR_INVALID; }
}

}  // extern "C"
