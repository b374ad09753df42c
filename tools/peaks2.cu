// Second-round FP64 pipe microbenchmarks: cycles per DMMA/DFMA (clock64), larger mma shapes,
// warp-specialised and in-warp mixes of DMMA and DFMA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/peaks2 tools/peaks2.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double* c, double a0, double a1, double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a0), "d"(a1), "d"(b));
}
__device__ __forceinline__ void dmma16816(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

struct Stat { unsigned long long cyc; unsigned long long ns; };

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t;
}

// mode 0: m8n8k4 x NACC; mode 1: m16n8k4; mode 2: m16n8k16
template <int NACC, int MODE>
__global__ void __launch_bounds__(512) k_mma(double* out, Stat* st, int iters, double a0, double b0) {
  double c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0; }
  double a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = a0 + threadIdx.x * 1e-9 + i;
#pragma unroll
  for (int i = 0; i < 4; ++i) b[i] = b0 + i * 1e-4;
  __syncthreads();
  unsigned long long t0 = clock64(), g0 = gtimer();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      if (MODE == 0) dmma884(c[i][0], c[i][1], a[0], b[0]);
      if (MODE == 1) dmma1684(c[i], a[0], a[1], b[0]);
      if (MODE == 2) dmma16816(c[i], a, b);
    }
  }
  unsigned long long t1 = clock64(), g1 = gtimer();
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 123.456) out[0] = s;
  if (blockIdx.x == 0 && threadIdx.x == 0) { st->cyc = t1 - t0; st->ns = g1 - g0; }
}

template <int NACC>
__global__ void __launch_bounds__(512) k_dfma(double* out, Stat* st, int iters, double a0, double b0) {
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i] = i;
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  __syncthreads();
  unsigned long long t0 = clock64(), g0 = gtimer();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], a, b);
  }
  unsigned long long t1 = clock64(), g1 = gtimer();
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i];
  if (s == 123.456) out[0] = s;
  if (blockIdx.x == 0 && threadIdx.x == 0) { st->cyc = t1 - t0; st->ns = g1 - g0; }
}

// warp-specialised: warps [0, wd) run DMMA (8 acc), others DFMA (16 acc); each runs `iters` trips
// of 16 instructions.  Reports per-role cycles via st[0] (dmma warp 0) and st[1] (first dfma warp).
__global__ void __launch_bounds__(512) k_spec(double* out, Stat* st, int iters, int wd, int dfma_mult, double a0, double b0) {
  int warp = threadIdx.x >> 5;
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  __syncthreads();
  unsigned long long t0 = clock64();
  double s = 0;
  if (warp < wd) {
    double c[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) { c[i][0] = 0; c[i][1] = 0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
  } else {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = i;
    for (int it = 0; it < iters * dfma_mult; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
  }
  unsigned long long t1 = clock64();
  if (s == 123.456) out[0] = s;
  if (blockIdx.x == 0 && (threadIdx.x == 0)) { st[0].cyc = t1 - t0; }
  if (blockIdx.x == 0 && (threadIdx.x == wd * 32)) { st[1].cyc = t1 - t0; }
}

// in-warp mix: per trip NM dmma + NF dfma, all independent
template <int NM, int NF>
__global__ void __launch_bounds__(512) k_mix(double* out, Stat* st, int iters, double a0, double b0) {
  double c[NM][2]; double f[NF];
#pragma unroll
  for (int i = 0; i < NM; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
#pragma unroll
  for (int i = 0; i < NF; ++i) f[i] = i;
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  __syncthreads();
  unsigned long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    constexpr int R = NF / NM;  // dfma per dmma
#pragma unroll
    for (int i = 0; i < NM; ++i) {
      dmma884(c[i][0], c[i][1], a, b);
#pragma unroll
      for (int r = 0; r < R; ++r) f[i * R + r] = fma(f[i * R + r], a, b);
    }
  }
  unsigned long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < NM; ++i) s += c[i][0] + c[i][1];
#pragma unroll
  for (int i = 0; i < NF; ++i) s += f[i];
  if (s == 123.456) out[0] = s;
  if (blockIdx.x == 0 && threadIdx.x == 0) { st->cyc = t1 - t0; }
}

template <typename F>
static float time_ms(F f, int reps) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, 1024));
  Stat* st; CK(cudaMallocManaged(&st, 4 * sizeof(Stat)));
  const int iters = 8192;
  printf("# name warps_per_sm tflops cyc_per_instr_per_smsp eff_mhz\n");
  int wps[] = {4, 8, 16};
  for (int w : wps) {
    float ms = time_ms([&] { k_mma<16, 0><<<sms, w * 32>>>(out, st, iters, 1.0, 1e-3); }, 3);
    double n_instr = 16.0 * iters;  // per warp
    printf("m8n8k4   w%-2d %7.2f  cyc/instr/smsp=%6.2f  mhz=%7.1f\n", w, 2.0 * 256 * n_instr * w * sms / ms * 1e-9,
           (double)st->cyc / (n_instr * (w / 4.0)), (double)st->cyc / st->ns * 1e3);
    ms = time_ms([&] { k_mma<8, 1><<<sms, w * 32>>>(out, st, iters, 1.0, 1e-3); }, 3);
    n_instr = 8.0 * iters;
    printf("m16n8k4  w%-2d %7.2f  cyc/instr/smsp=%6.2f  mhz=%7.1f\n", w, 2.0 * 512 * n_instr * w * sms / ms * 1e-9,
           (double)st->cyc / (n_instr * (w / 4.0)), (double)st->cyc / st->ns * 1e3);
    ms = time_ms([&] { k_mma<8, 2><<<sms, w * 32>>>(out, st, iters / 4, 1.0, 1e-3); }, 3);
    n_instr = 8.0 * (iters / 4);
    printf("m16n8k16 w%-2d %7.2f  cyc/instr/smsp=%6.2f  mhz=%7.1f\n", w, 2.0 * 2048 * n_instr * w * sms / ms * 1e-9,
           (double)st->cyc / (n_instr * (w / 4.0)), (double)st->cyc / st->ns * 1e3);
    ms = time_ms([&] { k_dfma<16><<<sms, w * 32>>>(out, st, iters, 1.0, 1e-3); }, 3);
    n_instr = 16.0 * iters;
    printf("dfma     w%-2d %7.2f  cyc/instr/smsp=%6.2f  mhz=%7.1f\n", w, 2.0 * 32 * n_instr * w * sms / ms * 1e-9,
           (double)st->cyc / (n_instr * (w / 4.0)), (double)st->cyc / st->ns * 1e3);
  }
  // one DMMA warp per SMSP with few accumulators: latency view
  {
    float ms = time_ms([&] { k_mma<1, 0><<<sms, 32>>>(out, st, iters, 1.0, 1e-3); }, 3);
    printf("m8n8k4 dependent chain: cyc/instr=%6.2f\n", (double)st->cyc / iters);
    ms = time_ms([&] { k_mma<2, 0><<<sms, 32>>>(out, st, iters, 1.0, 1e-3); }, 3);
    printf("m8n8k4 2 chains: cyc/instr=%6.2f\n", (double)st->cyc / (2.0 * iters));
    ms = time_ms([&] { k_mma<4, 0><<<sms, 32>>>(out, st, iters, 1.0, 1e-3); }, 3);
    printf("m8n8k4 4 chains: cyc/instr=%6.2f\n", (double)st->cyc / (4.0 * iters));
    ms = time_ms([&] { k_dfma<1><<<sms, 32>>>(out, st, iters, 1.0, 1e-3); }, 3);
    printf("dfma dependent chain: cyc/instr=%6.2f\n", (double)st->cyc / iters);
    (void)ms;
  }
  // warp-specialised mixes, 16 warps/SM: wd DMMA warps, rest DFMA. dfma_mult sets DFMA trips per DMMA trip.
  for (int wd : {4, 8, 12}) for (int mult : {1, 4, 8}) {
    int w = 16;
    float ms = time_ms([&] { k_spec<<<sms, w * 32>>>(out, st, iters, wd, mult, 1.0, 1e-3); }, 3);
    double fl = (2.0 * 256 * 16 * iters * wd + 2.0 * 32 * 16 * iters * mult * (w - wd)) * sms;
    printf("spec wd=%-2d mult=%d total=%7.2f TF (dmma part %.2f, dfma part %.2f)  cyc dmma-warp=%llu dfma-warp=%llu\n", wd, mult, fl / ms * 1e-9,
           2.0 * 256 * 16 * iters * wd * sms / ms * 1e-9, 2.0 * 32 * 16 * iters * mult * (w - wd) * sms / ms * 1e-9,
           st[0].cyc, st[1].cyc);
  }
  // in-warp mixes
  {
    int w = 16;
    float ms = time_ms([&] { k_mix<8, 8><<<sms, w * 32>>>(out, st, iters, 1.0, 1e-3); }, 3);
    printf("mix 1:1  total=%7.2f TF\n", (2.0 * 256 * 8 + 2.0 * 32 * 8) * iters * w * sms / ms * 1e-9);
    ms = time_ms([&] { k_mix<8, 16><<<sms, w * 32>>>(out, st, iters, 1.0, 1e-3); }, 3);
    printf("mix 1:2  total=%7.2f TF\n", (2.0 * 256 * 8 + 2.0 * 32 * 16) * iters * w * sms / ms * 1e-9);
    ms = time_ms([&] { k_mix<8, 32><<<sms, w * 32>>>(out, st, iters, 1.0, 1e-3); }, 3);
    printf("mix 1:4  total=%7.2f TF\n", (2.0 * 256 * 8 + 2.0 * 32 * 32) * iters * w * sms / ms * 1e-9);
    ms = time_ms([&] { k_mix<4, 32><<<sms, w * 32>>>(out, st, iters, 1.0, 1e-3); }, 3);
    printf("mix 1:8  total=%7.2f TF\n", (2.0 * 256 * 4 + 2.0 * 32 * 32) * iters * w * sms / ms * 1e-9);
  }
  return 0;
}
