/* Plain-C caller of libcorrla_b200.so: what a non-Python, non-Rust host (or the Rust -sys crate) does.
 * Build:  gcc -O2 -Iinclude -o tools/c_abi_example tools/c_abi_example.c -Lcorrla_rs_b200/lib -lcorrla_b200 \
 *             -Wl,-rpath,'$ORIGIN/../corrla_rs_b200/lib' -lm
 * Runs random_svd on a 2000 x 64 matrix of exact rank 12 with planted singular values 64, 63, ..., 53 (rank <= l = 18,
 * so the range finder captures it exactly), and on the reference's
 * known-answer 5 x 5 matrix (random_svd.rs:155-161), checks the results and prints the timings struct. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "corrla_b200.h"

static double urand(unsigned long long* s) {
  *s = *s * 6364136223846793005ULL + 1442695040888963407ULL;
  return (double)(*s >> 11) / 9007199254740992.0 - 0.5;
}

/* modified Gram-Schmidt, in place, column-major rows x cols */
static void orthonormalise(double* a, int rows, int cols) {
  for (int j = 0; j < cols; ++j) {
    for (int i = 0; i < j; ++i) {
      double d = 0; for (int r = 0; r < rows; ++r) d += a[i * rows + r] * a[j * rows + r];
      for (int r = 0; r < rows; ++r) a[j * rows + r] -= d * a[i * rows + r];
    }
    double n = 0; for (int r = 0; r < rows; ++r) n += a[j * rows + r] * a[j * rows + r];
    n = sqrt(n); for (int r = 0; r < rows; ++r) a[j * rows + r] /= n;
  }
}

int main(void) {
  printf("%s\n", corrla_version());
  const int m = 2000, n = 64, k = 8;
  unsigned long long seed = 12345;
  double* u0 = malloc(sizeof(double) * m * n); double* v0 = malloc(sizeof(double) * n * n);
  for (int i = 0; i < m * n; ++i) u0[i] = urand(&seed);
  for (int i = 0; i < n * n; ++i) v0[i] = urand(&seed);
  orthonormalise(u0, m, n); orthonormalise(v0, n, n);
  /* A = U0[:, :12] diag(64 - j) V0[:, :12]^T, stored ROW-major (row_stride = n, col_stride = 1), like a numpy array */
  double* a = calloc((size_t)m * n, sizeof(double));
  for (int r = 0; r < m; ++r) for (int c = 0; c < n; ++c) {
    double s = 0; for (int j = 0; j < 12; ++j) s += u0[j * m + r] * (double)(64 - j) * v0[j * n + c];
    a[(size_t)r * n + c] = s;
  }
  corrla_rsvd_opts opts; corrla_rsvd_opts_default(&opts); opts.seed = 7;
  corrla_timings t;
  double* u = malloc(sizeof(double) * m * k); double s[8]; double* vt = malloc(sizeof(double) * k * n);
  int st = corrla_rsvd_f64(a, m, n, n, 1, k, 4, 10, &opts, u, s, vt, &t);
  if (st != CORRLA_OK) { fprintf(stderr, "corrla_rsvd_f64: %s: %s\n", corrla_status_str(st), corrla_last_error()); return 1; }
  double worst = 0;
  for (int j = 0; j < k; ++j) { double e = fabs(s[j] - (64.0 - j)) / (64.0 - j); if (e > worst) worst = e; }
  printf("sigma[0..3] = %.12f %.12f %.12f %.12f   worst rel err = %.2e\n", s[0], s[1], s[2], s[3], worst);
  printf("passes over A = %d, kernel launches = %d, device ms = %.3f, h2d ms = %.3f\n", t.passes_over_a, t.gpu_launches,
         t.device_ms, t.h2d_ms);
  if (worst > 1e-10) { fprintf(stderr, "FAIL: singular values off\n"); return 2; }

  /* the reference's known-answer matrix, column-major (row_stride 1, col_stride 5) */
  const double a5[25] = {1, 0, 0, 0, 0,  0, 0, 0, 0, 2,  0, 3, 0, 0, 0,  0, 0, 0, 0, 0,  2, 0, 0, 0, 0};
  double u5[25], s5[5], v5[25];
  st = corrla_rsvd_f64(a5, 5, 5, 1, 5, 5, 12, 10, &opts, u5, s5, v5, NULL);
  if (st != CORRLA_OK) { fprintf(stderr, "5x5: %s: %s\n", corrla_status_str(st), corrla_last_error()); return 1; }
  printf("5x5 sigma = %.7f %.7f %.7f %.1e %.1e (expected 3, 2.2360679, 2, 0, 0)\n", s5[0], s5[1], s5[2], s5[3], s5[4]);
  if (fabs(s5[0] - 3) > 1e-3 || fabs(s5[1] - 2.2360679) > 1e-3 || fabs(s5[2] - 2) > 1e-3 || fabs(s5[3]) > 1e-3 || fabs(s5[4]) > 1e-3) return 3;

  /* error path: n_rank > l must come back as CORRLA_ERR_RANK, not crash */
  st = corrla_rsvd_f64(a, m, n, n, 1, 70, 2, 10, &opts, u, s, vt, NULL);
  printf("n_rank=70 on 64 columns -> %d (%s)\n", st, corrla_status_str(st));
  if (st != CORRLA_ERR_RANK) return 4;
  printf("C ABI example OK\n");
  return 0;
}
