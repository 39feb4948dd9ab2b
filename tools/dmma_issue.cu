// How many warps per scheduler does the FP64 tensor pipe need, and do the larger shapes need fewer?
// One CTA per SM; warps per SM = 4, 8, 16, 32; 8 independent accumulators per warp; operands refreshed from
// shared memory every iteration like a real kernel does.  Prints TFLOP/s per configuration.
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

template <int NACC, bool SMEM>
__global__ void k884(double* out, int iters) {
  __shared__ double sb[32 * 16];
  for (int i = threadIdx.x; i < 512; i += blockDim.x) sb[i] = 1e-3 * i;
  __syncthreads();
  double c[NACC][2] = {};
  double a = threadIdx.x * 1e-3;
  const int lane = threadIdx.x & 31;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < NACC; ++j) {
      const double b = SMEM ? sb[((i + j) & 15) * 32 + lane] : 1e-3;
      dmma884(c[j][0], c[j][1], a, b);
    }
  }
  double s = 0;
  for (int j = 0; j < NACC; ++j) s += c[j][0] + c[j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k1688(double* out, int iters) {
  __shared__ double sb[32 * 16 * 2];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sb[i] = 1e-3 * i;
  __syncthreads();
  double c[NACC][4] = {};
  double a[4] = {1e-3, 2e-3, 3e-3, 4e-3};
  const int lane = threadIdx.x & 31;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < NACC; ++j) {
      double b[2] = {sb[((i + j) & 15) * 64 + lane], sb[((i + j) & 15) * 64 + 32 + lane]};
      dmma1688(c[j], a, b);
    }
  }
  double s = 0;
  for (int j = 0; j < NACC; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
double best_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount, iters = 20000;
  double* out; cudaMalloc(&out, (size_t)sms * 1024 * 8);
  printf("{\"gpu\": \"%s\"", p.name);
  for (int warps : {4, 8, 16, 32}) {
    const double nw = (double)sms * warps;
    double t = best_ms([&] { k884<8, true><<<sms, warps * 32>>>(out, iters); });
    printf(", \"m8n8k4_acc8_smemB_%dw\": %.2f", warps, 2.0 * 256 * 8 * iters * nw / (t * 1e-3) / 1e12);
    t = best_ms([&] { k884<16, true><<<sms, warps * 32>>>(out, iters); });
    printf(", \"m8n8k4_acc16_smemB_%dw\": %.2f", warps, 2.0 * 256 * 16 * iters * nw / (t * 1e-3) / 1e12);
    t = best_ms([&] { k884<8, false><<<sms, warps * 32>>>(out, iters); });
    printf(", \"m8n8k4_acc8_regB_%dw\": %.2f", warps, 2.0 * 256 * 8 * iters * nw / (t * 1e-3) / 1e12);
    t = best_ms([&] { k1688<8><<<sms, warps * 32>>>(out, iters); });
    printf(", \"m16n8k8_acc8_smemB_%dw\": %.2f", warps, 2.0 * 1024 * 8 * iters * nw / (t * 1e-3) / 1e12);
  }
  printf("}\n");
  return cudaGetLastError() != cudaSuccess;
}
