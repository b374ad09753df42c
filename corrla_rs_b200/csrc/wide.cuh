// l = n_rank + n_oversamples above the 128 columns one register-tiled GEMM holds: the sketch is cut into P column
// panels of equal padded width w <= 128 (only the last panel has padding columns), every n x l / m x l matrix becomes P
// panel buffers in the engine's usual layout, and the RSVD of random_svd.rs:15-110 runs panel by panel on the same
// kernels:
//   passes      Y_p = A X_p and Z_p = A^T Y_p for every panel (P launches per pass; the flops per pass are unchanged)
//   thin Q      block classical Gram-Schmidt, two projection sweeps against the finished panels, then the adaptive
//               CholeskyQR of the single-panel path on the panel itself; Q is formed explicitly (no R^-1 folding)
//   SVD of B    W = Q_z^T Z_B assembled from w x w blocks, one-sided Jacobi on the l x l core (global-memory variant)
// Included by engine.cu only.
#pragma once
#include <string>
#include <vector>

#include "engine_core.cuh"

namespace corrla_eng {

struct Wide {
  Core& c;
  int P = 0, w = 0, l_total = 0, l_last = 0, Ltot16 = 0, ldW = 0;
  std::vector<double*> Y, Za, Zb, Qz;
  double *Cb = nullptr, *Cn = nullptr, *Bp = nullptr, *Tq = nullptr;
  double *W = nullptr, *Vr = nullptr, *Ur = nullptr, *sig = nullptr, *jscratch = nullptr;
  double *nu_slots = nullptr, *nutot = nullptr, *omega_tmp = nullptr;

  explicit Wide(Core& core) : c(core) {}

  // panel geometry for l columns: P panels of padded width w (multiple of 8, <= 128)
  static void plan(int l, int* P, int* w) { panel_plan(l, P, w); }
  int lp(int p) const { return p == P - 1 ? l_last : w; }
  bool multi() const { return c.comm != nullptr && c.comm->nranks > 1; }

  double* zeros(const std::string& name, size_t elems) {
    double* p = static_cast<double*>(c.ctx->get(name.c_str(), elems * 8));
    if (p != nullptr && cudaMemsetAsync(p, 0, elems * 8, c.st) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
  }

  // c.setup_dims(m, n, w), c.alloc_workspace and c.alloc_buffers(true) have run
  int alloc(int l) {
    l_total = l;
    plan(l, &P, &w);
    l_last = l - (P - 1) * w;
    if (w != c.Lc || l_last <= 0) { set_last_error("internal: panel plan mismatch (l=%d P=%d w=%d Lc=%d)", l, P, w, c.Lc); return CORRLA_ERR_INVALID; }
    Ltot16 = (int)round_up((int64_t)P * w, 16);
    ldW = P * w + 4;
    Y.assign(P, nullptr); Za.assign(P, nullptr); Zb.assign(P, nullptr); Qz.assign(P, nullptr);
    bool ok = true;
    for (int p = 0; p < P; ++p) {
      const std::string s = std::to_string(p);
      Y[p] = p == 0 ? c.Y : zeros("wide_Y" + s, (size_t)c.m16 * c.ld);
      Zb[p] = p == 0 ? c.Zb : zeros("wide_Zb" + s, (size_t)c.n16 * c.ld + 256);
      Qz[p] = p == 0 ? c.Qz : zeros("wide_Qz" + s, (size_t)c.n16 * c.ld);
      Za[p] = zeros("wide_Za" + s, (size_t)c.n16 * c.ld + 256);        // c.Za stays free: the QR refill uses it as scratch
      ok = ok && Y[p] && Zb[p] && Qz[p] && Za[p];
    }
    Cb = zeros("wide_Cb", c.small_elems()); Cn = zeros("wide_Cn", c.small_elems()); Bp = zeros("wide_Bp", c.small_elems());
    Tq = zeros("wide_Tq", c.small_elems());
    const size_t big = (size_t)Ltot16 * ldW;
    W = zeros("wide_W", big); Vr = zeros("wide_Vr", big); Ur = zeros("wide_Ur", big);
    sig = zeros("wide_sig", (size_t)Ltot16);
    jscratch = zeros("wide_jscratch", 2 * (size_t)l * (l + 2) + 8);
    nu_slots = zeros("wide_nu", (size_t)P + 8);
    nutot = nu_slots + P;
    ok = ok && Cb && Cn && Bp && Tq && W && Vr && Ur && sig && jscratch && nu_slots;
    if (!ok) { set_last_error("device allocation failed (wide sketch: l=%d in %d panels)", l, P); return CORRLA_ERR_ALLOC; }
    return CORRLA_OK;
  }

  int cu(cudaError_t e, const char* what) {
    if (e != cudaSuccess) { set_last_error("%s failed: %s", what, cudaGetErrorString(e)); cudaGetLastError(); return CORRLA_ERR_CUDA; }
    return CORRLA_OK;
  }

  // Omega (n x l): injected (any strides, host or device) or Philox with the same element numbering as the
  // single-panel path (draw number i*l + j), cut into the panel buffers
  int set_omega(const corrla_rsvd_opts& o) {
    if (o.omega != nullptr) {
      for (int p = 0; p < P; ++p)
        ST_TRY(pack_small(c.ctx, c.st, o.omega + (int64_t)p * w * o.omega_cs, c.n, lp(p), o.omega_rs, o.omega_cs,
                          o.omega_on_device != 0, Za[p], c.ld, 1.0, &c.launches));
      return CORRLA_OK;
    }
    omega_tmp = static_cast<double*>(c.ctx->get("wide_omega", (size_t)c.n * l_total * 8));
    if (!omega_tmp) { set_last_error("device allocation failed (Omega)"); return CORRLA_ERR_ALLOC; }
    ST_TRY(cu(philox_normal_launch(omega_tmp, c.n, l_total, l_total, o.seed, c.st), "philox"));
    ++c.launches;
    for (int p = 0; p < P; ++p) {
      ST_TRY(cu(repack_launch(omega_tmp + (int64_t)p * w, c.n, lp(p), l_total, 1, Za[p], c.ld, c.st), "repack"));
      ++c.launches;
    }
    return CORRLA_OK;
  }

  // Y_p = alpha * A * X_p for every panel; the squared Frobenius norm of the whole Y ends up in *nutot (all ranks)
  int passes_AX(std::vector<double*>& X, bool scaled) {
    for (int p = 0; p < P; ++p) ST_TRY(c.mm_AX(X[p], Y[p], scaled ? nutot : nullptr, nu_slots + p));
    if (multi()) ST_TRY(c.allreduce(nu_slots, (size_t)P));
    ST_TRY(cu(sum_array_launch(nu_slots, P, nutot, c.st), "norm reduction"));
    ++c.launches;
    return CORRLA_OK;
  }
  int passes_AtY() {
    for (int p = 0; p < P; ++p) ST_TRY(c.mm_AtY(Y[p], Zb[p]));
    return CORRLA_OK;
  }

  // X (rows x l in P panels) <- its thin-Q factor, explicitly.  basis_only (the QR inside the power loop): a well-conditioned
  // basis of the same range is enough (engine_core.cuh::qr_inplace) -- one projection sweep against the finished panels
  // instead of two and one Cholesky pass per panel; the result is orthonormal to ~1e-8 instead of 1e-15.
  int block_qr(std::vector<double*>& X, int64_t rows, bool distributed, double rows_for_shift, bool from_a, bool complete,
               bool basis_only = false) {
    const size_t gx = (distributed && multi()) ? (size_t)c.Lc * c.ld : 0;
    const int l_keep = c.l;
    int status = CORRLA_OK;
    for (int j = 0; j < P && status == CORRLA_OK; ++j) {
      c.l = lp(j);
      for (int attempt = 0; attempt < 3 && status == CORRLA_OK; ++attempt) {
        for (int rep = 0; rep < (basis_only ? 1 : 2) && status == CORRLA_OK; ++rep)
          for (int i = 0; i < j && status == CORRLA_OK; ++i) {
            const MatView qi = c.view_rows(X[i], rows);
            status = c.mm(qi, false, X[j], Cb, c.ld, 1, c.Lc, nullptr, nullptr, nullptr, 0, false, nullptr, gx);   // Q_i^T X_j
            if (status != CORRLA_OK) break;
            status = cu(repack_launch(Cb, c.Lc, c.Lc, c.ld, 1, Cn, c.ld, c.st, -1.0), "repack");
            ++c.launches;
            if (status != CORRLA_OK) break;
            status = c.mm(qi, true, Cn, X[j], c.ld, 1, c.Lc, nullptr, nullptr, nullptr, 0, false, nullptr, 0, 0, true); // X_j -= Q_i (..)
          }
        if (status != CORRLA_OK) break;
        const int refills = c.n_refill, robusts = c.n_robust;
        status = c.qr_inplace(X[j], rows, distributed, rows_for_shift, Tq, from_a, complete, basis_only);
        if (status == CORRLA_OK)      // Tq may be the symmetric first-order factor of the fast path: general product, in place
          status = c.mm(c.view_rows(X[j], rows), true, Tq, X[j], c.ld, 1, c.Lc, nullptr, nullptr, nullptr, 1);
        // Columns refilled inside the panel QR are not orthogonal to the earlier panels yet.  And a panel that needed the
        // robust stage may hold columns that COLLAPSED in the projection (exactly rank-deficient input: a column that lay in
        // the span of the earlier panels is rounding noise afterwards, full rank relative to itself, and the normalisation
        // blows that noise up together with its components along the earlier panels -- cross-orthogonality 1e-10 instead
        // of 1e-16, round-2 finding on the CPU model).  In both cases: project and factor again.
        if ((c.n_refill == refills && c.n_robust == robusts) || j == 0) break;
      }
    }
    c.l = l_keep;
    return status;
  }

  // power_iter (random_svd.rs:15-59); on return the Y panels hold Q explicitly
  int power_iter(const corrla_rsvd_opts& o, int n_iter) {
    ST_TRY(set_omega(o));
    ST_TRY(passes_AX(Za, false));                                   // :31
    for (int i = 0; i < n_iter; ++i) {                              // :35
      const bool do_qr = (o.schedule == 1) || (i > 2);              // :37
      if (do_qr) ST_TRY(block_qr(Y, c.m, true, c.grows, true, false, c.basis_only_qr));
      ST_TRY(passes_AtY());                                         // :42-46
      ST_TRY(passes_AX(Zb, !do_qr));                                // :47-51 (+ the deferred :53-55 scaling)
    }
    return block_qr(Y, c.m, true, c.grows, true, true);             // :57
  }

  // Q (m x l, column-major) out of the panels
  int scatter_q(double* qd) {
    for (int p = 0; p < P; ++p) {
      ST_TRY(cu(scatter_launch(Y[p], c.m, lp(p), c.ld, qd + (size_t)p * w * c.m, 1, c.m, c.st), "scatter"));
      ++c.launches;
    }
    return CORRLA_OK;
  }

  // out[:, :k] = [X_0 .. X_{P-1}] * M[:, :k]   (M: l x l, pitch ldW) with arbitrary output strides
  int combine(std::vector<double*>& X, int64_t rows, const double* M, double* out, int64_t ors, int64_t ocs, int k) {
    for (int c0 = 0; c0 < k; c0 += w) {
      const int wc = std::min(w, k - c0);
      for (int p = 0; p < P; ++p) {
        CU_TRY(cudaMemsetAsync(Bp, 0, c.small_elems() * 8, c.st));
        ST_TRY(cu(repack_launch(M + (size_t)p * w * ldW + c0, lp(p), wc, ldW, 1, Bp, c.ld, c.st), "repack"));
        ++c.launches;
        ST_TRY(c.mm(c.view_rows(X[p], rows), true, Bp, out + (int64_t)c0 * ocs, ors, ocs, wc, nullptr, nullptr, nullptr, 0,
                    false, nullptr, 0, 0, p > 0));
      }
    }
    return CORRLA_OK;
  }

  // B = Q^T A (:80), its SVD (:89), U = Q U~ (:92); thin-U / thin-V / sigma placement as in the single-panel path
  int finish(int k, double* Uthin_dst, int64_t u_rs, int64_t u_cs, double* Vthin_dst, int64_t v_rs, int64_t v_cs) {
    ST_TRY(passes_AtY());                                           // Z_B = A^T Q = B^T, n x l
    for (int p = 0; p < P; ++p)
      CU_TRY(cudaMemcpyAsync(Qz[p], Zb[p], (size_t)c.n16 * c.ld * 8, cudaMemcpyDeviceToDevice, c.st));
    ST_TRY(block_qr(Qz, c.n, false, (double)c.n, false, false));
    for (int pi = 0; pi < P; ++pi)
      for (int pj = 0; pj < P; ++pj)
        ST_TRY(c.mm(c.view_rows(Qz[pi], c.n), false, Zb[pj], W + (size_t)pi * w * ldW + (size_t)pj * w, ldW, 1, c.Lc));
    ST_TRY(cu(jacobi_svd_launch(W, ldW, l_total, sig, Vr, Ur, Ltot16, ldW, jscratch, c.flags + 4, c.st), "jacobi"));
    ++c.launches;
    if (Uthin_dst != nullptr) ST_TRY(combine(Y, c.m, Vr, Uthin_dst, u_rs, u_cs, k));
    if (Vthin_dst != nullptr) ST_TRY(combine(Qz, c.n, Ur, Vthin_dst, v_rs, v_cs, k));
    return CORRLA_OK;
  }
};

}  // namespace corrla_eng
