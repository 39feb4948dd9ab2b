// FP64 tensor-core primitives for sm_100a.
//
// tcgen05.mma has no f64 kind; the FP64 tensor path of Blackwell is the warp-level mma.sync DMMA.
// Measured on this pool's B200 (tools/fp64_peak.cu): DMMA m8n8k4 37.1 TFLOP/s, m16n8k8 36.8, DFMA 33.8,
// cuBLAS DGEMM 8192^3 35.4 -- all shapes share one FP64 datapath, so the kernels use the finest shape
// (m8n8k4), which tiles S = 20, 64 and 200 without padding along K.
//
// Fragment layout of mma.sync.m8n8k4.f64 (g = lane / 4, q = lane % 4):
//   A (8x4, row)  : a  = A[g][q]
//   B (4x8, col)  : b  = B[q][g]
//   C (8x8)       : c0 = C[g][2q], c1 = C[g][2q+1]
#pragma once
#include "common.cuh"

namespace bppgpu {

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

}  // namespace bppgpu
