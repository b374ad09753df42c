// Standalone correctness + speed check of the skinny DMMA GEMM (no Python, no engine).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -Icorrla_rs_b200/csrc \
//        -o tools/test_gemm tools/test_gemm.cu corrla_rs_b200/csrc/skinny_gemm.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "skinny_gemm.cuh"
using namespace corrla;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

static double urand() { return (double)rand() / RAND_MAX - 0.5; }

__global__ void fill_kernel(double* p, size_t n, unsigned seed) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    unsigned x = (unsigned)(i * 2654435761u) ^ seed; x ^= x >> 13; x *= 0x5bd1e995u; x ^= x >> 15;
    p[i] = (double)(x & 0xffff) / 65536.0 - 0.5;
  }
}

// outer x inner matrix (ld), contraction over inner or outer with B (K x l), compared with a CPU loop.
static int check(int64_t outer, int64_t inner, int64_t ld, int l, bool reduce_inner, int force_splits, GemmWorkspace& w) {
  const int nblk = (l + 7) / 8, Lc = nblk * 8, ldb = Lc + 4;
  const int64_t K = reduce_inner ? inner : outer, M = reduce_inner ? outer : inner;
  const int64_t Kp = (K + 15) / 16 * 16;
  std::vector<double> hA(outer * ld), hB(Kp * ldb, 0.0), hO(M * ldb, -7.0), ref(M * l, 0.0);
  for (auto& v : hA) v = urand();
  for (int64_t k = 0; k < K; ++k) for (int j = 0; j < l; ++j) hB[k * ldb + j] = urand();
  for (int64_t o = 0; o < outer; ++o) for (int64_t i = 0; i < inner; ++i) {
    const double a = hA[o * ld + i];
    const int64_t m = reduce_inner ? o : i, k = reduce_inner ? i : o;
    for (int j = 0; j < l; ++j) ref[m * l + j] += a * hB[k * ldb + j];
  }
  double *dA, *dB, *dO, *dS;
  CK(cudaMalloc(&dA, hA.size() * 8)); CK(cudaMalloc(&dB, hB.size() * 8)); CK(cudaMalloc(&dO, hO.size() * 8));
  CK(cudaMalloc(&dS, 8));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dO, hO.data(), hO.size() * 8, cudaMemcpyHostToDevice));
  GemmCall c{};
  c.a = MatView{dA, inner, outer, ld}; c.reduce_inner = reduce_inner;
  c.B = dB; c.ldb = ldb; c.nblk = nblk; c.out = dO; c.out_rs = ldb; c.out_cs = 1; c.ncols_out = Lc;
  c.sumsq_slot = dS; c.force_splits = force_splits;
  cudaError_t e = gemm_launch(c, w, 0);
  if (e != cudaSuccess) { printf("launch error: %s\n", cudaGetErrorString(e)); return 1; }
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("exec error: %s\n", cudaGetErrorString(e)); exit(2); }
  double ss = 0;
  CK(cudaMemcpy(hO.data(), dO, hO.size() * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&ss, dS, 8, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0, refss = 0; int bad_pad = 0;
  for (int64_t m = 0; m < M; ++m) {
    for (int j = 0; j < l; ++j) {
      maxerr = fmax(maxerr, fabs(hO[m * ldb + j] - ref[m * l + j])); maxref = fmax(maxref, fabs(ref[m * l + j]));
      refss += ref[m * l + j] * ref[m * l + j];
    }
    for (int j = l; j < Lc; ++j) if (hO[m * ldb + j] != 0.0) ++bad_pad;
    for (int j = Lc; j < ldb; ++j) if (hO[m * ldb + j] != -7.0) ++bad_pad;
  }
  const double rel = maxerr / (maxref + 1e-300), ssrel = fabs(ss - refss) / refss;
  const bool ok = rel < 1e-13 && bad_pad == 0 && ssrel < 1e-12;
  printf("%s outer=%lld inner=%lld ld=%lld l=%d %s splits=%d : rel_err=%.2e sumsq_rel=%.2e bad_pad=%d\n", ok ? "PASS" : "FAIL",
         (long long)outer, (long long)inner, (long long)ld, l, reduce_inner ? "reduce_inner" : "reduce_outer", force_splits,
         rel, ssrel, bad_pad);
  cudaFree(dA); cudaFree(dB); cudaFree(dO); cudaFree(dS);
  return ok ? 0 : 1;
}

static void speed(int64_t outer, int64_t inner, int l, bool reduce_inner, GemmWorkspace& w, const char* name) {
  const int nblk = (l + 7) / 8, Lc = nblk * 8, ldb = Lc + 4;
  const int64_t K = reduce_inner ? inner : outer, M = reduce_inner ? outer : inner;
  const int64_t Kp = (K + 15) / 16 * 16;
  double *dA, *dB, *dO;
  CK(cudaMalloc(&dA, outer * inner * 8)); CK(cudaMalloc(&dB, Kp * ldb * 8)); CK(cudaMalloc(&dO, M * ldb * 8));
  fill_kernel<<<1184, 256>>>(dA, outer * inner, 1u); fill_kernel<<<1184, 256>>>(dB, Kp * ldb, 2u);
  CK(cudaMemset(dO, 0, M * ldb * 8));
  GemmCall c{};
  c.a = MatView{dA, inner, outer, inner}; c.reduce_inner = reduce_inner;
  c.B = dB; c.ldb = ldb; c.nblk = nblk; c.out = dO; c.out_rs = ldb; c.out_cs = 1; c.ncols_out = Lc;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 2; ++i) { cudaError_t e = gemm_launch(c, w, 0); if (e != cudaSuccess) { printf("launch error %s\n", cudaGetErrorString(e)); return; } }
  CK(cudaDeviceSynchronize());
  const int reps = 5;
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) gemm_launch(c, w, 0);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
  int tilesM, splits; int64_t cps; size_t wsb, np;
  gemm_plan(M, K, nblk, w.num_sms, 0, &tilesM, &splits, &cps, &wsb, &np);
  printf("SPEED %-28s outer=%lld inner=%lld l=%d splits=%d : %.3f ms  %.2f TFLOP/s (useful, l=%d)  %.1f GB/s of A\n", name,
         (long long)outer, (long long)inner, l, splits, ms, 2.0 * outer * inner * l / ms * 1e-9, l, outer * inner * 8.0 / ms * 1e-6);
  cudaFree(dA); cudaFree(dB); cudaFree(dO);
}

int main(int argc, char** argv) {
  GemmWorkspace w;
  w.ws_bytes = (size_t)600 << 20; CK(cudaMalloc(&w.ws, w.ws_bytes));
  w.n_partials = 1 << 22; CK(cudaMalloc(&w.sumsq_partials, w.n_partials * 8));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0)); w.num_sms = p.multiProcessorCount;
  int fails = 0;
  srand(1);
  // small shapes first: every kernel variant (nblk 1..16), both contractions, with and without split-K
  for (int l : {110, 8, 18, 64, 128, 3, 40, 50, 76, 90, 100, 120, 24, 30, 56}) {
    fails += check(300, 200, 200, l, true, 0, w);
    fails += check(300, 200, 204, l, false, 0, w);
  }
  fails += check(1000, 333 * 2, 666, 110, true, 3, w);
  fails += check(1000, 333 * 2, 666, 110, false, 5, w);
  fails += check(5000, 130, 132, 110, true, 0, w);
  fails += check(5000, 130, 132, 110, false, 0, w);
  fails += check(5000, 130, 132, 110, false, 7, w);
  fails += check(129, 1030, 1030, 18, true, 2, w);
  fails += check(17, 5, 6, 5, true, 0, w);
  fails += check(17, 5, 6, 5, false, 0, w);
  // short output side of a reduce_outer product with a narrow sketch: the k-group path (1, 2 and 4 boxes), few and many
  // splits (serial and warp-parallel reduction), ragged K
  fails += check(100000, 64, 64, 18, false, 0, w);
  fails += check(100003, 24, 24, 18, false, 0, w);
  fails += check(50000, 12, 12, 10, false, 0, w);
  fails += check(3000, 40, 40, 30, false, 3, w);
  fails += check(77, 64, 64, 18, false, 0, w);
  fails += check(5000, 16, 16, 50, false, 0, w);
  fails += check(70000, 100, 100, 18, false, 0, w);
  printf("correctness failures: %d\n", fails);
  if (argc > 1) {
    speed(1 << 20, 1024, 110, true, w, "Y=A*X row-major (K2)");
    speed(1 << 20, 1024, 110, false, w, "Z=A^T*Y row-major (K3)");
    speed(1024, 1 << 20, 110, false, w, "Y=A*X col-major (K2)");
    speed(1024, 1 << 20, 110, true, w, "Z=A^T*Y col-major (K3)");
    speed(20000, 20000, 110, true, w, "C2 Y=A*X");
    speed(20000, 20000, 110, false, w, "C2 Z=A^T*Y");
    speed(1 << 20, 112, 110, false, w, "Gram Y^T*Y");
    speed(1 << 20, 112, 110, true, w, "apply Y*T");
    speed(1 << 20, 64, 18, true, w, "C5 Y=A*X");
    speed(1 << 20, 64, 18, false, w, "C5 Z=A^T*Y");
    speed(1 << 20, 24, 18, false, w, "C5 Gram Y^T*Y");
    speed(1 << 20, 24, 18, true, w, "C5 apply Y*T");
  }
  return fails ? 1 : 0;
}
