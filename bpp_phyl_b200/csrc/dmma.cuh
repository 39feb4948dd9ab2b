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

// ---- skinny panel GEMM on the FP64 tensor cores -----------------------------------------------------------------------
// C[M x 32] = A[M x K] . Bs[K x 32]: A row-major in global memory (L2-resident), the 32-column panel of B in shared memory
// with a leading dimension of 36 doubles (= 4 mod 16: conflict-free B fragments), 8 warps, warp w owns the row blocks
// w, w + 8, w + 16, w + 24 (M <= 256) times the panel's four column blocks.  Used by chr_level_kernel (columns = the
// branches of a tree level) and by the series kernel's matrix products (columns = a panel of the right factor).
constexpr int kChrCols = 32;      // columns (branches) per tile
constexpr int kChrLD = 36;        // leading dimension of the shared column tiles (= 4 mod 16 doubles: conflict-free B fragments)
constexpr int kChrWarps = 8;
constexpr int kChrMaxRB = 4;      // row blocks of 8 per warp -> S <= 8 * 8 * 4 = 256
// C[rb][cb] += A[rows of rb][k] . Bs[k][cols of cb]   for this warp's row blocks; A row-major [S][S] in global memory
// COHERENT: A was written earlier by this same kernel (ld.global.cg: L2, never a stale L1 line); otherwise it is read-only for
// the launch (ld.global.nc).
// NCB = column blocks of 8 in use (a tree level with 5 branches needs one, not four), KU = k-steps whose A fragments are fetched
// together.  The A fragments of the NEXT KU k-steps are in flight while the current ones feed the tensor cores (two register
// sets), so the L2 latency of A -- the only operand that does not sit in shared memory -- is covered by 4 KU loads per lane;
// a narrow tile has few DMMAs per load and takes a deep KU, a full one the opposite.  Accumulation order over k is the same for
// every (NCB, KU): results do not depend on them.
template <bool COHERENT, int NCB, int KU>
__device__ __forceinline__ void chr_gemm_t(const double* __restrict__ A, int S, int K4, const double* Bs, int nrb, int warp, int g,
                                           int q, double (&acc)[kChrMaxRB][kChrCols / 8][2]) {
#pragma unroll
  for (int i = 0; i < kChrMaxRB; ++i)
#pragma unroll
    for (int cb = 0; cb < NCB; ++cb) acc[i][cb][0] = acc[i][cb][1] = 0.0;
  const double* arow[kChrMaxRB];
  bool rok[kChrMaxRB];
#pragma unroll
  for (int i = 0; i < kChrMaxRB; ++i) {
    const int row = (warp + i * kChrWarps) * 8 + g;
    rok[i] = warp + i * kChrWarps < nrb && row < S;
    arow[i] = A + (size_t)(rok[i] ? row : 0) * S + q;
  }
  auto fetch = [&](int kbase, double (&a)[KU][kChrMaxRB]) {
#pragma unroll
    for (int u = 0; u < KU; ++u) {
      const int k = kbase + 4 * u;
#pragma unroll
      for (int i = 0; i < kChrMaxRB; ++i) {
        const bool in = rok[i] && k + q < S;
        a[u][i] = !in ? 0.0 : (COHERENT ? __ldcg(arow[i] + k) : __ldg(arow[i] + k));   // L2 / read-only path
      }
    }
  };
  auto feed = [&](int kbase, const double (&a)[KU][kChrMaxRB]) {
#pragma unroll
    for (int u = 0; u < KU; ++u) {
      const int k = kbase + 4 * u;
      if (k < K4) {   // uniform
        double b[NCB];
#pragma unroll
        for (int cb = 0; cb < NCB; ++cb) b[cb] = Bs[(k + q) * kChrLD + cb * 8 + g];
#pragma unroll
        for (int i = 0; i < kChrMaxRB; ++i)
          if (warp + i * kChrWarps < nrb) {
#pragma unroll
            for (int cb = 0; cb < NCB; ++cb) dmma884(acc[i][cb][0], acc[i][cb][1], a[u][i], b[cb]);
          }
      }
    }
  };
  double a0[KU][kChrMaxRB], a1[KU][kChrMaxRB];
  fetch(0, a0);
  for (int k0 = 0; k0 < K4; k0 += 8 * KU) {
    fetch(k0 + 4 * KU, a1);
    feed(k0, a0);
    fetch(k0 + 8 * KU, a0);
    feed(k0 + 4 * KU, a1);
  }
}
template <bool COHERENT = false>
__device__ __forceinline__ void chr_gemm(const double* __restrict__ A, int S, int K4, const double* Bs, int nrb, int warp, int g,
                                         int q, double (&acc)[kChrMaxRB][kChrCols / 8][2]) {
  chr_gemm_t<COHERENT, kChrCols / 8, 1>(A, S, K4, Bs, nrb, warp, g, q, acc);
}
// the same with the number of column blocks chosen at run time (uniform over the CTA)
__device__ __forceinline__ void chr_gemm_ncb(int ncb, const double* __restrict__ A, int S, int K4, const double* Bs, int nrb, int warp,
                                             int g, int q, double (&acc)[kChrMaxRB][kChrCols / 8][2]) {
  switch (ncb) {
    case 1: chr_gemm_t<false, 1, 4>(A, S, K4, Bs, nrb, warp, g, q, acc); break;
    case 2: chr_gemm_t<false, 2, 2>(A, S, K4, Bs, nrb, warp, g, q, acc); break;
    case 3: chr_gemm_t<false, 3, 1>(A, S, K4, Bs, nrb, warp, g, q, acc); break;
    default: chr_gemm_t<false, 4, 1>(A, S, K4, Bs, nrb, warp, g, q, acc); break;
  }
}
__device__ __forceinline__ void chr_store_acc(double* Cs, int nrb, int warp, int g, int q,
                                              const double (&acc)[kChrMaxRB][kChrCols / 8][2], int ncb = kChrCols / 8) {
#pragma unroll
  for (int i = 0; i < kChrMaxRB; ++i)
    if (warp + i * kChrWarps < nrb) {
      const int row = (warp + i * kChrWarps) * 8 + g;
#pragma unroll
      for (int cb = 0; cb < kChrCols / 8; ++cb)
        if (cb < ncb) *reinterpret_cast<double2*>(Cs + row * kChrLD + cb * 8 + 2 * q) = make_double2(acc[i][cb][0], acc[i][cb][1]);
    }
}


}  // namespace bppgpu
