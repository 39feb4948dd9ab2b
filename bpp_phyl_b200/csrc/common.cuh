// Shared device/host helpers for the bppgpu kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

namespace bppgpu {

// ---- thread-local error string behind bppgpu_last_error() -------------------
std::string& last_error();

#define BPP_FAIL(code, ...)                            \
  do {                                                 \
    char _b[512];                                      \
    snprintf(_b, sizeof(_b), __VA_ARGS__);             \
    ::bppgpu::last_error() = _b;                       \
    return (code);                                     \
  } while (0)

#define BPP_CUDA(expr)                                                                  \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess)                                                              \
      BPP_FAIL(_e == cudaErrorMemoryAllocation ? BPPGPU_E_NOMEM : BPPGPU_E_CUDA,        \
               "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__,  \
               cudaGetErrorString(_e));                                                 \
  } while (0)

// ---- 256-bit global accesses (LDG.E.256 / STG.E.256 on sm_100a) -------------
__device__ __forceinline__ void ld256(const double* p, double& a, double& b, double& c, double& d) {
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
// read-only data (P tables, tip tables): non-coherent path
__device__ __forceinline__ void ld256nc(const double* p, double& a, double& b, double& c, double& d) {
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
__device__ __forceinline__ void st256(double* p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%4], {%0,%1,%2,%3};" ::"d"(a), "d"(b), "d"(c), "d"(d), "l"(p) : "memory");
}

// ---- programmatic dependent launch (one kernel per tree node, hundreds per evaluation) --------------------------------------
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while its predecessor in the stream is
// still draining: its CTAs take the SMs the predecessor frees, stage what does not depend on it (P operands), and only then
// wait for the predecessor's results.  pdl_wait() returns once the predecessor grid has completed and its writes are visible
// (it returns at once in a kernel launched without the attribute); every kernel of such a chain must call it, so that
// "my predecessor is complete" implies "all earlier kernels are complete".
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- mbarrier / TMA bulk copy (PTX ISA 8.x, sm_90+) ------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "W4C_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra W4C_DONE_%=;\n"
      "bra W4C_WAIT_%=;\n"
      "W4C_DONE_%=:\n"
      "}\n" ::"r"(smem_u32(b)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(b))
               : "memory");
}

// ---- power-of-two rescaling (exact in FP64) ----------------------------------
// A pattern is rescaled when its largest CLV entry drops below 2^-256; the rule
// (and therefore every exponent) is identical in oracle/ref_likelihood.py::_rescale.
constexpr int kScaleThresholdExp = -256;
constexpr int kScaleThresholdHi = (1023 + kScaleThresholdExp) << 20;

// For non-negative doubles the IEEE ordering equals the integer ordering of the
// high word, so the per-pattern maximum is tracked on the integer pipe.
__device__ __forceinline__ int hi_word(double v) { return __double2hiint(v); }

// k such that max * 2^k lies in [0.5, 1)  (frexp convention), from the max's high word
__device__ __forceinline__ int rescale_shift(int max_hi) { return 1022 - (max_hi >> 20); }
__device__ __forceinline__ double pow2(int k) { return __hiloint2double((k + 1023) << 20, 0); }

constexpr double kLn2 = 0.693147180559945309417232121458;

// 2^-d for d >= 0 (aligning a class row to the pattern's smallest exponent); 0 once it would be subnormal
__device__ __forceinline__ double align_factor(int d) { return d > 1000 ? 0.0 : pow2(-d); }

// ---- deterministic block sum (fixed shuffle tree, then warp 0) ---------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// all threads of the block must call; result valid in thread 0
__device__ __forceinline__ double block_sum(double v, double* sm /* >= 32 doubles */) {
  v = warp_sum(v);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sm[w] = v;
  __syncthreads();
  int nw = (blockDim.x + 31) >> 5;
  double r = 0.0;
  if (w == 0) {
    r = l < nw ? sm[l] : 0.0;
    r = warp_sum(r);
  }
  return r;
}

}  // namespace bppgpu
