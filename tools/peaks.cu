// Microbenchmarks that pin the FP64 roofline denominators on this B200:
//   DMMA (mma.sync m8n8k4 f64) issue rate, DFMA rate, both together, HBM read rate.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/peaks tools/peaks.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void __launch_bounds__(1024) k_dmma(double* out, int iters, double a0, double b0) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}

template <int NACC>
__global__ void __launch_bounds__(1024) k_dfma(double* out, int iters, double a0, double b0) {
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i] = i;
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i];
  if (s == 123.456) out[0] = s;
}

// DMMA and DFMA interleaved in one warp: NM dmma + NF dfma per inner trip
template <int NM, int NF>
__global__ void __launch_bounds__(1024) k_mixed(double* out, int iters, double a0, double b0) {
  double c[NM][2]; double f[NF];
#pragma unroll
  for (int i = 0; i < NM; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
#pragma unroll
  for (int i = 0; i < NF; ++i) f[i] = i;
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < (NM > NF ? NM : NF); ++i) {
      if (i < NM) dmma884(c[i][0], c[i][1], a, b);
      if (i < NF) f[i] = fma(f[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NM; ++i) s += c[i][0] + c[i][1];
#pragma unroll
  for (int i = 0; i < NF; ++i) s += f[i];
  if (s == 123.456) out[0] = s;
}

__global__ void k_read(const double2* __restrict__ p, size_t n2, double* out) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  double s = 0;
  for (; i + 3 * stride < n2; i += 4 * stride) {
    double2 a = p[i], b = p[i + stride], c = p[i + 2 * stride], d = p[i + 3 * stride];
    s += a.x + a.y + b.x + b.y + c.x + c.y + d.x + d.y;
  }
  for (; i < n2; i += stride) { double2 a = p[i]; s += a.x + a.y; }
  if (s == 123.456) out[0] = s;
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
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d,\n", p.name, sms, p.clockRate);
  double* out; CK(cudaMalloc(&out, 1024));
  const int iters = 4096;
  // --- DMMA: sweep warps per SM
  int wps[] = {4, 8, 16, 32};
  printf(" \"dmma_tflops\": {");
  for (int wi = 0; wi < 4; ++wi) {
    int w = wps[wi];
    float ms = time_ms([&] { k_dmma<16><<<sms, w * 32>>>(out, iters, 1.0, 1e-3); }, 5);
    double fl = 2.0 * 256 * 16.0 * iters * w * sms;
    printf("%s\"w%d\": %.2f", wi ? ", " : "", w, fl / ms * 1e-9);
  }
  printf("},\n \"dfma_tflops\": {");
  for (int wi = 0; wi < 4; ++wi) {
    int w = wps[wi];
    float ms = time_ms([&] { k_dfma<16><<<sms, w * 32>>>(out, iters, 1.0, 1e-3); }, 5);
    double fl = 2.0 * 32 * 16.0 * iters * w * sms;
    printf("%s\"w%d\": %.2f", wi ? ", " : "", w, fl / ms * 1e-9);
  }
  printf("},\n \"mixed_tflops\": {");
  {
    float ms = time_ms([&] { k_mixed<8, 8><<<sms, 16 * 32>>>(out, iters, 1.0, 1e-3); }, 5);
    double fl = (2.0 * 256 * 8 + 2.0 * 32 * 8) * iters * 16.0 * sms;
    printf("\"m8f8_w16\": %.2f", fl / ms * 1e-9);
    ms = time_ms([&] { k_mixed<8, 64><<<sms, 16 * 32>>>(out, iters / 4, 1.0, 1e-3); }, 5);
    fl = (2.0 * 256 * 8 + 2.0 * 32 * 64) * (iters / 4) * 16.0 * sms;
    printf(", \"m8f64_w16\": %.2f", fl / ms * 1e-9);
    ms = time_ms([&] { k_mixed<8, 32><<<sms, 16 * 32>>>(out, iters / 4, 1.0, 1e-3); }, 5);
    fl = (2.0 * 256 * 8 + 2.0 * 32 * 32) * (iters / 4) * 16.0 * sms;
    printf(", \"m8f32_w16\": %.2f", fl / ms * 1e-9);
  }
  printf("},\n");
  // sustained DMMA (about 2 s) to see the power-capped rate
  {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    int launches = 0;
    for (; launches < 400; ++launches) k_dmma<16><<<sms, 16 * 32>>>(out, iters * 4, 1.0, 1e-3);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    double fl = 2.0 * 256 * 16.0 * iters * 4 * 16 * sms * launches;
    printf(" \"dmma_sustained_tflops\": %.2f, \"dmma_sustained_seconds\": %.2f,\n", fl / ms * 1e-9, ms * 1e-3);
  }
  // --- HBM read
  {
    size_t bytes = (size_t)8 << 30; double2* buf; CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 0, bytes));
    float ms = time_ms([&] { k_read<<<sms * 16, 512>>>(buf, bytes / 16, out); }, 5);
    printf(" \"hbm_read_gbs\": %.1f}\n", bytes / ms * 1e-6);
    CK(cudaFree(buf));
  }
  return 0;
}
