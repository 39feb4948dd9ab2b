// Measures the FP64 ceilings of this B200 that MEASURED_PEAKS.json does not hold:
//   DFMA (CUDA-core FP64 pipe) and DMMA (mma.sync f64, the only FP64 tensor path on sm_100a; tcgen05 has no f64 kind).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/fp64_peak tools/fp64_peak.cu ; prints one JSON line.
#include <cuda_runtime.h>
#include <cstdio>

__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

template <int K>
__device__ __forceinline__ void dmma(double (&c)[4], const double (&a)[K / 2], const double (&b)[K / 4]);

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

__global__ void dmma884_kernel(double* out, int iters) {
  double c[8][2] = {};
  double a = threadIdx.x * 1e-3, b = 1e-3;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) dmma884(c[j][0], c[j][1], a, b);
  }
  double s = 0;
  for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void dmma1688_kernel(double* out, int iters) {
  double c[8][4] = {};
  double a[4] = {1e-3, 2e-3, 3e-3, 4e-3}, b[2] = {1e-3, 2e-3};
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) dmma1688(c[j], a, b);
  }
  double s = 0;
  for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void dmma16816_kernel(double* out, int iters) {
  double c[8][4] = {};
  double a[8] = {1e-3, 2e-3, 3e-3, 4e-3, 1e-3, 2e-3, 3e-3, 4e-3}, b[4] = {1e-3, 2e-3, 1e-3, 2e-3};
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) dmma16816(c[j], a, b);
  }
  double s = 0;
  for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
double time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); f();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount, blocks = sms * 8, threads = 512, iters = 20000;
  double* out; cudaMalloc(&out, (size_t)blocks * threads * 8);
  double t;
  t = time_ms([&] { dfma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
  double dfma = 2.0 * 8 * iters * (double)blocks * threads / (t * 1e-3) / 1e12;
  const double warps = (double)blocks * threads / 32;
  t = time_ms([&] { dmma884_kernel<<<blocks, threads>>>(out, iters); });
  double d884 = 2.0 * 8 * 8 * 4 * 8 * iters * warps / (t * 1e-3) / 1e12;
  t = time_ms([&] { dmma1688_kernel<<<blocks, threads>>>(out, iters); });
  double d1688 = 2.0 * 16 * 8 * 8 * 8 * iters * warps / (t * 1e-3) / 1e12;
  t = time_ms([&] { dmma16816_kernel<<<blocks, threads>>>(out, iters); });
  double d16816 = 2.0 * 16 * 8 * 16 * 8 * iters * warps / (t * 1e-3) / 1e12;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"dfma_tflops\": %.2f, \"dmma_m8n8k4_tflops\": %.2f, \"dmma_m16n8k8_tflops\": %.2f, \"dmma_m16n8k16_tflops\": %.2f}\n",
         p.name, sms, dfma, d884, d1688, d16816);
  if (cudaGetLastError() != cudaSuccess) return 1;
  return 0;
}
