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
// ---- the same product with A streamed through a warp-private cp.async ring --------------------------------------------------------
// A warp only ever reads the A rows of its own row blocks, so each warp brings them in by itself: per k-step the 32 rows x 4
// doubles (1 KB) of its fragments, into a ring of kChrStages slots in shared memory, kChrStages - 1 k-steps ahead of the tensor
// cores.  The bytes in flight live in shared memory instead of registers (5 KB per warp against 0.25 KB with register double
// buffering), which is what covers the L2 / HBM latency of A; no block-level barrier sits in the k loop (cp.async.wait_group +
// __syncwarp).  Fragment reads of the ring are conflict-free (row stride 32 B: bank pair 8 g + 2 q within a half warp).
// Rows / k beyond S are zero-filled by the copy itself (src-size operand).  Accumulation order over k as in chr_gemm_t.
constexpr int kChrStages = 6;
constexpr int kChrRingDoubles = kChrStages * 32 * 4;   // per warp
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gmem_src, int valid_bytes) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem_src), "r"(valid_bytes) : "memory");
}
__device__ __forceinline__ void cp_async8_zfill(void* smem_dst, const void* gmem_src, int valid_bytes) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(gmem_src), "r"(valid_bytes) : "memory");
}
template <int NCB>
__device__ __forceinline__ void chr_gemm_ring(const double* __restrict__ A, int S, int K4, const double* Bs, double* ring, int nrb,
                                              int warp, int lane, double (&acc)[kChrMaxRB][kChrCols / 8][2]) {
  const int g = lane >> 2, q = lane & 3;
#pragma unroll
  for (int i = 0; i < kChrMaxRB; ++i)
#pragma unroll
    for (int cb = 0; cb < NCB; ++cb) acc[i][cb][0] = acc[i][cb][1] = 0.0;
  const int nk = K4 >> 2;
  const bool al16 = (S & 1) == 0 && (reinterpret_cast<size_t>(A) & 15) == 0;   // uniform: 16-byte pieces need aligned rows
  auto issue = [&](int t, int slot) {
    double* dst = ring + slot * 128;
    const int k0 = t * 4;
    if (al16) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int p = lane + 32 * j, r = p >> 1, h = p & 1;
        const int rb = warp + (r >> 3) * kChrWarps, row = rb * 8 + (r & 7), kc = k0 + 2 * h;
        int valid = (rb < nrb && row < S) ? (S - kc) * 8 : 0;
        valid = valid < 0 ? 0 : (valid > 16 ? 16 : valid);
        cp_async16_zfill(dst + r * 4 + 2 * h, valid ? A + (size_t)row * S + kc : A, valid);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int p = lane + 32 * j, r = p >> 2, c = p & 3;
        const int rb = warp + (r >> 3) * kChrWarps, row = rb * 8 + (r & 7), kc = k0 + c;
        const int valid = (rb < nrb && row < S && kc < S) ? 8 : 0;
        cp_async8_zfill(dst + r * 4 + c, valid ? A + (size_t)row * S + kc : A, valid);
      }
    }
  };
#pragma unroll
  for (int t = 0; t < kChrStages - 1; ++t) {
    if (t < nk) issue(t, t);
    cp_async_commit();
  }
  int slot = 0, fill = kChrStages - 1;   // slot read by this k-step; slot the next copy goes to (the one read one k-step ago)
  for (int t = 0; t < nk; ++t) {
    cp_async_wait<kChrStages - 2>();   // this lane's pieces of k-step t have landed ...
    __syncwarp();                      // ... and so have the other lanes'; everyone is done reading the slot of k-step t - 1
    if (t + kChrStages - 1 < nk) issue(t + kChrStages - 1, fill);
    cp_async_commit();
    const double* as = ring + slot * 128;
    double a[kChrMaxRB], b[NCB];
#pragma unroll
    for (int i = 0; i < kChrMaxRB; ++i) a[i] = as[(i * 8 + g) * 4 + q];
#pragma unroll
    for (int cb = 0; cb < NCB; ++cb) b[cb] = Bs[(4 * t + q) * kChrLD + cb * 8 + g];
#pragma unroll
    for (int i = 0; i < kChrMaxRB; ++i)
      if (warp + i * kChrWarps < nrb) {
#pragma unroll
        for (int cb = 0; cb < NCB; ++cb) dmma884(acc[i][cb][0], acc[i][cb][1], a[i], b[cb]);
      }
    slot = slot + 1 == kChrStages ? 0 : slot + 1;
    fill = fill + 1 == kChrStages ? 0 : fill + 1;
  }
  cp_async_wait<0>();
  __syncwarp();
}
// ---- the same product with A streamed by the TMA engine in k-slabs ------------------------------------------------------------
// The warp-private ring above keeps 5 KB per warp in flight in 16-byte pieces from 32 different rows per k-step: the A stream
// (2 S^2 doubles per tile, nothing else comes from L2) ran at 16 GB/s per SM, and a tile with a handful of columns -- 43 of the 47
// levels of the 500-taxon benchmark tree -- spent 8x longer waiting for A than on its DMMAs.  Here every model keeps a second copy
// of V^-1 and V in SLAB order (chr_slab_kernel): slab ks = the K8 rows x 8 k-columns [8 ks, 8 ks + 8) as one contiguous block of
// K8 x 8 doubles (12.8 KB at S = 200), so a slab is ONE cp.async.bulk.  A producer warp streams the slabs of the CTA's whole tile
// sequence ([V^-1 | V] per dense tile, V per observed-tip tile) through an mbarrier full/empty ring; it never waits for the
// element-wise phases between the GEMMs, so the ring is full again when the next product starts.  7 consumer warps (25 row blocks:
// the busiest warp owns 4 with 7 warps as with 8).  Within a slab the 8 doubles of row r are stored XOR-swizzled,
// position = kk ^ (4 * ((r >> 1) & 1)): the fragment load a = A[rb * 8 + g][4 h + q] of a half warp (g = 0..3, q = 0..3) then
// touches 16 distinct bank pairs.  Accumulation order over k as in chr_gemm_t: results are bit-identical to the other variants.
constexpr int kChrCons = 7;   // consumer warps (warp kChrCons is the producer)
struct ChrRing {
  unsigned long long* full;    // [nst] count 1 (+ transaction bytes)
  unsigned long long* empty;   // [nst] count kChrCons
  double* slab;                // [nst][K8 * 8]
  int nst;
  int stage;
  unsigned phase;
};
__host__ __device__ inline int chr_slab_doubles(int K8) { return K8 * 8; }
template <int NCB>
__device__ __forceinline__ void chr_gemm_slab(ChrRing& r, int K8, const double* Bs, int nrb, int warp, int lane,
                                              double (&acc)[kChrMaxRB][kChrCols / 8][2]) {
  const int g = lane >> 2, q = lane & 3;
#pragma unroll
  for (int i = 0; i < kChrMaxRB; ++i)
#pragma unroll
    for (int cb = 0; cb < NCB; ++cb) acc[i][cb][0] = acc[i][cb][1] = 0.0;
  const int sw = 4 * ((g >> 1) & 1);
  const int slabD = chr_slab_doubles(K8);
  const int nslab = K8 >> 3;
  for (int ks = 0; ks < nslab; ++ks) {
    mbar_wait(r.full + r.stage, r.phase);
    const double* as = r.slab + (size_t)r.stage * slabD + g * 8;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      double a[kChrMaxRB], b[NCB];
#pragma unroll
      for (int i = 0; i < kChrMaxRB; ++i) a[i] = (warp + i * kChrCons < nrb) ? as[(warp + i * kChrCons) * 64 + ((4 * h + q) ^ sw)] : 0.0;
#pragma unroll
      for (int cb = 0; cb < NCB; ++cb) b[cb] = Bs[(8 * ks + 4 * h + q) * kChrLD + cb * 8 + g];
#pragma unroll
      for (int i = 0; i < kChrMaxRB; ++i)
        if (warp + i * kChrCons < nrb) {
#pragma unroll
          for (int cb = 0; cb < NCB; ++cb) dmma884(acc[i][cb][0], acc[i][cb][1], a[i], b[cb]);
        }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(r.empty + r.stage);
    if (++r.stage == r.nst) { r.stage = 0; r.phase ^= 1u; }
  }
}
__device__ __forceinline__ void chr_gemm_slab_ncb(int ncb, ChrRing& r, int K8, const double* Bs, int nrb, int warp, int lane,
                                                  double (&acc)[kChrMaxRB][kChrCols / 8][2]) {
  switch (ncb) {
    case 1: chr_gemm_slab<1>(r, K8, Bs, nrb, warp, lane, acc); break;
    case 2: chr_gemm_slab<2>(r, K8, Bs, nrb, warp, lane, acc); break;
    case 3: chr_gemm_slab<3>(r, K8, Bs, nrb, warp, lane, acc); break;
    default: chr_gemm_slab<4>(r, K8, Bs, nrb, warp, lane, acc); break;
  }
}
// producer side: n slabs starting at src (contiguous) into the ring; one thread
__device__ __forceinline__ void chr_ring_produce(ChrRing& r, const double* src, int n, int K8) {
  const int slabD = chr_slab_doubles(K8);
  for (int s = 0; s < n; ++s) {
    mbar_wait(r.empty + r.stage, r.phase ^ 1u);   // (a fresh barrier passes a wait on parity 1: the first lap does not block)
    mbar_expect_tx(r.full + r.stage, (unsigned)slabD * 8u);
    bulk_g2s(r.slab + (size_t)r.stage * slabD, src + (size_t)s * slabD, (unsigned)slabD * 8u, r.full + r.stage);
    if (++r.stage == r.nst) { r.stage = 0; r.phase ^= 1u; }
  }
}
template <int NW>
__device__ __forceinline__ void chr_store_acc_w(double* Cs, int nrb, int warp, int g, int q,
                                                const double (&acc)[kChrMaxRB][kChrCols / 8][2], int ncb) {
#pragma unroll
  for (int i = 0; i < kChrMaxRB; ++i)
    if (warp + i * NW < nrb) {
      const int row = (warp + i * NW) * 8 + g;
#pragma unroll
      for (int cb = 0; cb < kChrCols / 8; ++cb)
        if (cb < ncb) *reinterpret_cast<double2*>(Cs + row * kChrLD + cb * 8 + 2 * q) = make_double2(acc[i][cb][0], acc[i][cb][1]);
    }
}

// the number of column blocks chosen at run time (uniform over the CTA)
__device__ __forceinline__ void chr_gemm_ncb(int ncb, const double* __restrict__ A, int S, int K4, const double* Bs, double* ring, int nrb,
                                             int warp, int lane, double (&acc)[kChrMaxRB][kChrCols / 8][2]) {
  switch (ncb) {
    case 1: chr_gemm_ring<1>(A, S, K4, Bs, ring, nrb, warp, lane, acc); break;
    case 2: chr_gemm_ring<2>(A, S, K4, Bs, ring, nrb, warp, lane, acc); break;
    case 3: chr_gemm_ring<3>(A, S, K4, Bs, ring, nrb, warp, lane, acc); break;
    default: chr_gemm_ring<4>(A, S, K4, Bs, ring, nrb, warp, lane, acc); break;
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
