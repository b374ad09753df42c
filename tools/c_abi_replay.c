/* Replays, in plain C, the exact call sequences of the (uncompilable here) Rust wrapper crate rust/corrla-b200/src/lib.rs:
 * default options + a seed, every `timings` pointer NULL, column-major faer-style outputs, optional outputs NULL.
 *   random_svd        -> corrla_rsvd_f64        (lib.rs: random_svd)
 *   power_iter        -> corrla_power_iter_f64  (lib.rs: power_iter)
 *   par_matmul_helper -> corrla_par_matmul_f64  (lib.rs: par_matmul_helper; on_device 0, opts NULL)
 *   dmdc_operators    -> corrla_dmdc_f64        (lib.rs: dmdc_operators; omega_y NULL)
 *   pod_modes_weights -> corrla_pod_f64         (lib.rs: pod_modes_weights; s NULL)
 * and checks what a caller can check without a reference: orthonormal factors, reconstruction, exact products.
 * Build: gcc -O2 -Iinclude -o c_abi_replay tools/c_abi_replay.c -Lcorrla_rs_b200/lib -lcorrla_b200 -lm */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "corrla_b200.h"

static double urand(unsigned long long* s) {
  *s = *s * 6364136223846793005ULL + 1442695040888963407ULL;
  return (double)(*s >> 11) / 9007199254740992.0 - 0.5;
}
#define CHECK(st, what) do { if ((st) != CORRLA_OK) { fprintf(stderr, "%s: %s: %s\n", what, corrla_status_str(st), corrla_last_error()); return 1; } } while (0)

/* max |Q^T Q - I| for a column-major rows x cols matrix */
static double ortho_err(const double* q, int rows, int cols) {
  double worst = 0;
  for (int i = 0; i < cols; ++i) for (int j = 0; j <= i; ++j) {
    double d = 0; for (int r = 0; r < rows; ++r) d += q[(size_t)i * rows + r] * q[(size_t)j * rows + r];
    d = fabs(d - (i == j ? 1.0 : 0.0)); if (d > worst) worst = d;
  }
  return worst;
}

int main(void) {
  unsigned long long seed = 99;
  /* ---- random_svd on a faer Mat (column-major: row_stride 1, col_stride nrows), exact rank 6 */
  const int m = 3000, n = 80, r = 6, k = 6;
  double* lf = malloc(sizeof(double) * m * r); double* rf = malloc(sizeof(double) * r * n);
  for (int i = 0; i < m * r; ++i) lf[i] = urand(&seed);
  for (int i = 0; i < r * n; ++i) rf[i] = urand(&seed);
  double* a = calloc((size_t)m * n, sizeof(double));
  for (int c = 0; c < n; ++c) for (int j = 0; j < r; ++j) for (int i = 0; i < m; ++i) a[(size_t)c * m + i] += lf[(size_t)j * m + i] * rf[j * n + c];
  corrla_rsvd_opts opts; corrla_rsvd_opts_default(&opts); opts.seed = 0x1234abcdULL;
  double* u = calloc((size_t)m * k, sizeof(double)); double* s = calloc(k, sizeof(double)); double* vt = calloc((size_t)k * n, sizeof(double));
  int st = corrla_rsvd_f64(a, m, n, 1, m, k, 4, 10, &opts, u, s, vt, NULL);
  CHECK(st, "random_svd");
  double rec = 0, amax = 0;
  for (int c = 0; c < n; ++c) for (int i = 0; i < m; ++i) {
    double v = 0; for (int j = 0; j < k; ++j) v += u[(size_t)j * m + i] * s[j] * vt[(size_t)c * k + j];
    double d = fabs(v - a[(size_t)c * m + i]); if (d > rec) rec = d;
    if (fabs(a[(size_t)c * m + i]) > amax) amax = fabs(a[(size_t)c * m + i]);
  }
  printf("random_svd: ortho(U) %.1e, reconstruction %.1e (max |a| %.2f)\n", ortho_err(u, m, k), rec, amax);
  if (ortho_err(u, m, k) > 1e-12 || rec > 1e-10 * amax) return 2;

  /* ---- power_iter */
  const int l = 12;
  double* q = calloc((size_t)m * l, sizeof(double));
  st = corrla_power_iter_f64(a, m, n, 1, m, l, 3, &opts, q, NULL);
  CHECK(st, "power_iter");
  printf("power_iter: ortho(Q) %.1e\n", ortho_err(q, m, l));
  if (ortho_err(q, m, l) > 1e-12) return 3;

  /* ---- par_matmul_helper: res = beta * lhs * rhs, host pointers, opts NULL */
  const int mm = 500, kk = 40, nn = 7;
  double* lhs = malloc(sizeof(double) * mm * kk); double* rhs = malloc(sizeof(double) * kk * nn); double* res = calloc((size_t)mm * nn, sizeof(double));
  for (int i = 0; i < mm * kk; ++i) lhs[i] = urand(&seed);
  for (int i = 0; i < kk * nn; ++i) rhs[i] = urand(&seed);
  st = corrla_par_matmul_f64(res, 1, mm, lhs, mm, kk, 1, mm, rhs, nn, 1, kk, 2.0, 0, NULL);
  CHECK(st, "par_matmul_helper");
  double worst = 0;
  for (int j = 0; j < nn; ++j) for (int i = 0; i < mm; ++i) {
    double v = 0; for (int t = 0; t < kk; ++t) v += lhs[(size_t)t * mm + i] * rhs[(size_t)j * kk + t];
    double d = fabs(res[(size_t)j * mm + i] - 2.0 * v); if (d > worst) worst = d;
  }
  printf("par_matmul_helper: max err %.1e\n", worst);
  if (worst > 1e-12) return 4;

  /* ---- dmdc_operators: x (n_x x n_snap) and u (n_u x n_snap) column-major, x_{t+1} = A x_t + B u_t with rank-3 A */
  const int nx = 400, nu = 1, ns = 30, rr = 4;
  double* basis = malloc(sizeof(double) * nx * 3);
  for (int i = 0; i < nx * 3; ++i) basis[i] = urand(&seed);
  for (int j = 0; j < 3; ++j) {            /* Gram-Schmidt */
    for (int i = 0; i < j; ++i) { double d = 0; for (int t = 0; t < nx; ++t) d += basis[i * nx + t] * basis[j * nx + t]; for (int t = 0; t < nx; ++t) basis[j * nx + t] -= d * basis[i * nx + t]; }
    double nr = 0; for (int t = 0; t < nx; ++t) nr += basis[j * nx + t] * basis[j * nx + t]; nr = sqrt(nr); for (int t = 0; t < nx; ++t) basis[j * nx + t] /= nr;
  }
  const double lam[3] = {0.9, -0.7, 0.5};
  double* x = calloc((size_t)nx * ns, sizeof(double)); double* uu = malloc(sizeof(double) * nu * ns);
  double coef[3] = {1.0, -0.5, 0.25};
  const double bcoef[3] = {0.3, 0.2, -0.4};
  for (int t = 0; t < ns; ++t) {
    uu[t] = urand(&seed);
    for (int i = 0; i < nx; ++i) x[(size_t)t * nx + i] = coef[0] * basis[i] + coef[1] * basis[nx + i] + coef[2] * basis[2 * nx + i];
    for (int j = 0; j < 3; ++j) coef[j] = lam[j] * coef[j] + bcoef[j] * uu[t];
  }
  double* a_til = calloc(rr * rr, sizeof(double)); double* b = calloc((size_t)nx * nu, sizeof(double));
  double* ms = calloc((size_t)nx * rr, sizeof(double)); double* s_til = calloc(rr, sizeof(double)); double* u_hat = calloc((size_t)nx * rr, sizeof(double));
  st = corrla_dmdc_f64(x, nx, ns, 1, nx, uu, nu, 1, nu, rr, 4, &opts, NULL, a_til, b, ms, s_til, u_hat, NULL);
  CHECK(st, "dmdc_operators");
  /* B must be sum_j bcoef[j] * basis_j */
  worst = 0;
  for (int i = 0; i < nx; ++i) { double v = bcoef[0] * basis[i] + bcoef[1] * basis[nx + i] + bcoef[2] * basis[2 * nx + i]; double d = fabs(b[i] - v); if (d > worst) worst = d; }
  double tr = 0; for (int i = 0; i < rr; ++i) tr += a_til[i * rr + i];
  printf("dmdc_operators: |B - B_true| %.1e, trace(A~) %.6f (expected %.6f), ortho(U^) %.1e\n", worst, tr, lam[0] + lam[1] + lam[2], ortho_err(u_hat, nx, rr));
  if (worst > 1e-8 || fabs(tr - (lam[0] + lam[1] + lam[2])) > 1e-8) return 5;

  /* ---- pod_modes_weights: x is n_snap x n_points (fat), column-major */
  const int psn = 16, ppt = 5000, pm = 5;
  double* px = malloc(sizeof(double) * psn * ppt);
  for (int i = 0; i < psn * ppt; ++i) px[i] = urand(&seed);
  double* modes = calloc((size_t)ppt * pm, sizeof(double)); double* weights = calloc((size_t)psn * pm, sizeof(double));
  st = corrla_pod_f64(px, psn, ppt, 1, psn, pm, &opts, modes, weights, NULL, NULL);
  CHECK(st, "pod_modes_weights");
  worst = 0;                                /* weights = x * modes */
  for (int j = 0; j < pm; ++j) for (int i = 0; i < psn; ++i) {
    double v = 0; for (int t = 0; t < ppt; ++t) v += px[(size_t)t * psn + i] * modes[(size_t)j * ppt + t];
    double d = fabs(weights[j * psn + i] - v); if (d > worst) worst = d;
  }
  printf("pod_modes_weights: ortho(modes) %.1e, |weights - x*modes| %.1e\n", ortho_err(modes, ppt, pm), worst);
  if (ortho_err(modes, ppt, pm) > 1e-12 || worst > 1e-10) return 6;
  printf("C ABI replay OK\n");
  return 0;
}
