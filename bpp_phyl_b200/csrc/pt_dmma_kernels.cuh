// K1 on the FP64 tensor cores: branch-batched P(t) = (V . diag(f(lambda tau))) . V^-1 for S >= 32
// (codon S = 64, chromosome S ~ 200), one batched GEMM launch for every (point, branch, class).
//   f = exp(x)            -> pxy_    (Model/AbstractSubstitutionModel.cpp:436)
//   f = r  (rate lambda)   exp(x)   -> dpxy_   (:505, times r_c: AbstractHomogeneousTreeLikelihood.cpp:389)
//   f = r^2 (rate lambda)^2 exp(x)  -> d2pxy_  (:576, :409)
// Conjugate eigen-pairs use the real block form (:438-468, :507-537, :578-612): P = V.T.V^-1 with 2x2 blocks
// [[dia, up], [-up, dia]], i.e. column k of the left factor is dia_k V[:,k] + off_k V[:,k^1] with off = -up for the
// first and +up for the second member.  The engine uploads the eigen-columns permuted so that every pair starts at an
// even index (ModelDev::Vp / Vinvp / rep / imp), which keeps a pair inside one k-block.  Singular generators keep the
// series kernel of pt_kernels.cuh.
//
// Tiling: one CTA per (matrix, column panel).  WARPS warps, warp w owns MBW 8-row blocks (all of M = S =
// WARPS*MBW*8 is covered by the CTA) times the panel's NB 8-column blocks: MBW*NB DMMA atoms per k-step
// against MBW + NB fragment loads.  V and V^-1 are streamed through shared memory in k-slabs of 8 by a
// 2-stage cp.async pipeline; the eigenvalue factor is applied to the A fragment on the way to the
// registers.  Shared-memory row strides are = 4 (mod 16) doubles so that every fragment load is
// conflict free (SURVEY/DESIGN: half-warp (g<4, q) -> bank pairs g*4+q).
#pragma once
#include "dmma.cuh"
#include "pt_kernels.cuh"

namespace bppgpu {

constexpr int kPtKT = 8;        // k-slab
constexpr int kPtAStride = 12;  // doubles per staged V row (8 + 4 pad)

// Sp = S rounded up to a multiple of 8; Vp / Vinvp / rep are the model's arrays zero-padded to Sp
// (ModelDev::Vp ...; identical to V / Vinv / re when S % 8 == 0).  Row block mb of warp w, slot i is
// w + i*WARPS (interleaved, predicated on mb < Sp/8); the CTA covers all rows and NB column blocks.
template <int NB>
__host__ __device__ constexpr int pt_bstride() { return NB * 8 + 4; }

constexpr int kPtStages = 3;    // cp.async pipeline depth
inline size_t pt_dmma_smem_bytes(int Sp, int NB) {
  return (size_t)(6 * Sp + kPtStages * Sp * kPtAStride + kPtStages * kPtKT * (NB * 8 + 4)) * sizeof(double);
}

template <int WARPS, int MBW, int NB>
__global__ void __launch_bounds__(WARPS * 32, 2) pt_dmma_kernel(PtParams p, int Sp) {
  constexpr int NT = WARPS * 32;
  constexpr int BStride = pt_bstride<NB>();
  const int S = p.S;
  const int kAStage = Sp * kPtAStride;
  constexpr int kBStage = kPtKT * BStride;
  extern __shared__ __align__(16) double sm_pt[];
  double* dtab = sm_pt;               // [3][Sp] diagonal factors
  double* otab = dtab + 3 * Sp;       // [3][Sp] partner-column factors (0 for real eigenvalues)
  double* As = otab + 3 * Sp;                // [stages][Sp][12]
  double* Bs = As + kPtStages * kAStage;     // [stages][8][BStride]

  const int m = blockIdx.x;
  const int c = m % p.C;
  const int node = (m / p.C) % p.nn;
  const int point = m / (p.C * p.nn);
  if (node == p.root) return;
  const ModelDev md = p.models[p.branch_model[point * p.nn + node]];
  if (!(md.flags & 2u)) return;  // singular generator: series kernel
  const double rc = p.rates[c];
  const double t = p.brlen[point * p.nn + node] * rc;
  const double l = md.rate * t;
  const int n0 = blockIdx.y * NB * 8;  // first column of this panel
  const int nblk = Sp >> 3;

  for (int k = threadIdx.x; k < Sp; k += NT) {
    const double bk = md.has_complex ? md.imp[k] : 0.0;
    if (bk == 0.0) {
      const double a = md.rep[k];
      const double ex = exp(a * l);
      const double ra = md.rate * a;
      dtab[k] = ex;
      dtab[Sp + k] = rc * (ra * ex);
      dtab[2 * Sp + k] = rc * rc * (ra * ra * ex);
      otab[k] = otab[Sp + k] = otab[2 * Sp + k] = 0.0;
    } else {
      const int kf = k & ~1;  // first member holds +im
      const double ar = md.rep[kf], b = md.imp[kf];
      const double ex = exp(ar * l);
      double sn, cs;
      sincos(b * l, &sn, &cs);
      const double r1 = md.rate, r2 = md.rate * md.rate;
      const double sg = (k & 1) ? 1.0 : -1.0;
      dtab[k] = ex * cs;
      dtab[Sp + k] = rc * (r1 * (ar * cs - b * sn) * ex);
      dtab[2 * Sp + k] = rc * rc * (r2 * ((ar * ar - b * b) * cs - 2.0 * ar * b * sn) * ex);
      otab[k] = sg * ex * sn;
      otab[Sp + k] = sg * rc * (r1 * (ar * sn + b * cs) * ex);
      otab[2 * Sp + k] = sg * rc * rc * (r2 * ((ar * ar - b * b) * sn + 2.0 * ar * b * cs) * ex);
    }
  }

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  const bool chr_deriv = md.flags & 8u;
  const bool clamp = md.flags & 4u;
  const size_t base = (size_t)m * S * S;

  auto stage = [&](int ks, int buf) {
    // Vp[:, ks*8 .. +8) : Sp rows x 64 B = 4 x 16 B chunks per row
    double* a_dst = As + buf * kAStage;
    for (int i = threadIdx.x; i < Sp * 4; i += NT) {
      const int row = i >> 2, ch = i & 3;
      cp_async16(a_dst + row * kPtAStride + ch * 2, md.Vp + (size_t)row * Sp + ks * kPtKT + ch * 2);
    }
    // Vinvp[ks*8 .. +8, n0 .. n0 + NB*8)
    double* b_dst = Bs + buf * kBStage;
    for (int i = threadIdx.x; i < kPtKT * NB * 4; i += NT) {
      const int row = i / (NB * 4), ch = i - row * (NB * 4);
      if (n0 + ch * 2 < Sp)
        cp_async16(b_dst + row * BStride + ch * 2, md.Vinvp + (size_t)(ks * kPtKT + row) * Sp + n0 + ch * 2);
    }
  };

  for (int tab = 0; tab < 3; ++tab) {
    const bool wanted = (p.want >> tab) & 1u;
    // Chromosome models rebuild dP/d2P from the unclamped P (pt_chr_deriv_kernel): only P (+Pun) here
    if (!wanted || (tab > 0 && chr_deriv)) continue;
    double acc[MBW][NB][2];
#pragma unroll
    for (int i = 0; i < MBW; ++i)
#pragma unroll
      for (int j = 0; j < NB; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    const double* dt = dtab + tab * Sp;
    const double* ot = otab + tab * Sp;
    const bool cplx = md.has_complex != 0;

    __syncthreads();  // dtab ready / previous table's smem reads done
    const int NKS = Sp / kPtKT;
    // multistage pipeline, one barrier per slab: slab ks+2 is requested right after the barrier that proves everybody
    // is done with slab ks-1 (whose buffer it re-uses); one commit group per iteration keeps the wait count fixed
#pragma unroll
    for (int s0 = 0; s0 < kPtStages - 1; ++s0) {
      if (s0 < NKS) stage(s0, s0);
      cp_async_commit();
    }
    for (int ks = 0; ks < NKS; ++ks) {
      const int buf = ks % kPtStages;
      cp_async_wait<kPtStages - 2>();
      __syncthreads();
      if (ks + kPtStages - 1 < NKS) stage(ks + kPtStages - 1, (ks + kPtStages - 1) % kPtStages);
      cp_async_commit();
      const double* a_src = As + buf * kAStage;
      const double* b_src = Bs + buf * kBStage;
#pragma unroll
      for (int kb = 0; kb < kPtKT / 4; ++kb) {
        const double dk = dt[ks * kPtKT + kb * 4 + q];
        const double ok = cplx ? ot[ks * kPtKT + kb * 4 + q] : 0.0;
        double a[MBW], b[NB];
#pragma unroll
        for (int i = 0; i < MBW; ++i) {
          const int mb = warp + i * WARPS;
          double av = 0.0;
          if (mb < nblk) {
            const double* ar = a_src + (mb * 8 + g) * kPtAStride + kb * 4;
            av = ar[q] * dk;
            if (cplx) av = fma(ar[q ^ 1], ok, av);
          }
          a[i] = av;
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) b[j] = b_src[(kb * 4 + q) * BStride + j * 8 + g];
#pragma unroll
        for (int i = 0; i < MBW; ++i) {
          if (warp + i * WARPS < nblk) {  // warp-uniform
#pragma unroll
            for (int j = 0; j < NB; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
          }
        }
      }
    }
    cp_async_wait<0>();

    double* out = tab == 0 ? p.P : tab == 1 ? p.dP : p.d2P;
#pragma unroll
    for (int i = 0; i < MBW; ++i) {
      const int x = (warp + i * WARPS) * 8 + g;
      if (x >= S) continue;
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const int y = n0 + j * 8 + 2 * q;
        double v[2] = {acc[i][j][0], acc[i][j][1]};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (y + h >= S) continue;
          double val = v[h];
          if (tab == 0) {
            if (t == 0.0) val = x == y + h ? 1.0 : 0.0;  // AbstractSubstitutionModel.cpp:428-431
            if (chr_deriv && p.Pun) p.Pun[base + (size_t)x * S + y + h] = val;
            if (clamp) val = val < 0.0 ? 1e-20 : (val > 1.0 ? 1.0 : val);  // ChromosomeSubstitutionModel.cpp:903-916
          }
          out[base + (size_t)x * S + y + h] = val;
        }
      }
    }
  }
}

// ---- stacked variant --------------------------------------------------------------------------------------------------
// When every branch of a point uses the same model (the usual case: homogeneous likelihoods, ChromEvol), all matrices of
// the point share V and V^-1 and differ only in the diagonal factor, so their 8-row blocks can be STACKED along M: a CTA
// takes WARPS*MBW consecutive row blocks of the point's (nn*C matrices x nblk blocks) list, whatever matrices they belong
// to.  Every warp then owns exactly MBW blocks (S = 200 has 25 blocks per matrix: the per-matrix kernel leaves 7 of 32
// block slots idle, 22 %).  A CTA touches at most kPtMaxMats matrices; their factor tables are rebuilt per table.
constexpr int kPtMaxMats = 6;

inline size_t pt_dmma_stacked_smem_bytes(int Sp, int NB) {
  return (size_t)(2 * kPtMaxMats * Sp + kPtStages * Sp * kPtAStride + kPtStages * kPtKT * (NB * 8 + 4)) * sizeof(double);
}

template <int WARPS, int MBW, int NB>
__global__ void __launch_bounds__(WARPS * 32, 2) pt_dmma_stacked_kernel(PtParams p, int Sp) {
  constexpr int NT = WARPS * 32;
  constexpr int BStride = pt_bstride<NB>();
  constexpr int RB = WARPS * MBW;  // row blocks per CTA
  const int S = p.S;
  const int kAStage = Sp * kPtAStride;
  constexpr int kBStage = kPtKT * BStride;
  extern __shared__ __align__(16) double sm_pt[];
  double* dtab = sm_pt;                        // [kPtMaxMats][Sp]
  double* otab = dtab + kPtMaxMats * Sp;       // [kPtMaxMats][Sp]
  double* As = otab + kPtMaxMats * Sp;         // [stages][Sp][12]
  double* Bs = As + kPtStages * kAStage;       // [stages][8][BStride]
  __shared__ double tau_s[kPtMaxMats];

  const int point = blockIdx.z;
  const int nblk = Sp >> 3;
  const int nmat_pt = p.nn * p.C;              // matrices of this point (the root's slot included: t = 0 -> identity)
  const long long total_rb = (long long)nmat_pt * nblk;
  const long long rb0 = (long long)blockIdx.x * RB;
  if (rb0 >= total_rb) return;
  const int mat0 = (int)(rb0 / nblk);
  const int mat_last = (int)(min(rb0 + RB, total_rb) - 1) / nblk;
  const int nm = mat_last - mat0 + 1;          // <= kPtMaxMats (checked on the host)
  const ModelDev md = p.models[p.branch_model[point * p.nn + (p.root == 0 ? 1 : 0)]];
  const int n0 = blockIdx.y * NB * 8;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  const bool chr_deriv = md.flags & 8u;
  const bool clamp = md.flags & 4u;
  const bool cplx = md.has_complex != 0;

  // this warp's row blocks
  int mloc[MBW], mb[MBW];
  bool valid[MBW];
#pragma unroll
  for (int i = 0; i < MBW; ++i) {
    const long long rb = rb0 + warp + i * WARPS;
    valid[i] = rb < total_rb;
    const long long rbc = valid[i] ? rb : rb0;
    mloc[i] = (int)(rbc / nblk) - mat0;
    mb[i] = (int)(rbc % nblk);
  }

  auto stage = [&](int ks, int buf) {
    double* a_dst = As + buf * kAStage;
    for (int i = threadIdx.x; i < Sp * 4; i += NT) {
      const int row = i >> 2, ch = i & 3;
      cp_async16(a_dst + row * kPtAStride + ch * 2, md.Vp + (size_t)row * Sp + ks * kPtKT + ch * 2);
    }
    double* b_dst = Bs + buf * kBStage;
    for (int i = threadIdx.x; i < kPtKT * NB * 4; i += NT) {
      const int row = i / (NB * 4), ch = i - row * (NB * 4);
      if (n0 + ch * 2 < Sp)
        cp_async16(b_dst + row * BStride + ch * 2, md.Vinvp + (size_t)(ks * kPtKT + row) * Sp + n0 + ch * 2);
    }
  };

  for (int tab = 0; tab < 3; ++tab) {
    const bool wanted = (p.want >> tab) & 1u;
    if (!wanted || (tab > 0 && chr_deriv)) continue;
    __syncthreads();  // previous table's smem reads done
    // factor tables of the matrices this CTA touches
    for (int e = threadIdx.x; e < nm * Sp; e += NT) {
      const int ml = e / Sp, k = e - ml * Sp;
      const int mi = mat0 + ml;
      const int c = mi % p.C, node = mi / p.C;
      const double rc = p.rates[c];
      const double t = p.brlen[point * p.nn + node] * rc;
      const double l = md.rate * t;
      if (k == 0) tau_s[ml] = t;
      const double bk = cplx ? md.imp[k] : 0.0;
      const double rpow = tab == 0 ? 1.0 : (tab == 1 ? rc : rc * rc);
      double dv, ov = 0.0;
      if (bk == 0.0) {
        const double a = md.rep[k];
        const double ex = exp(a * l);
        const double ra = md.rate * a;
        dv = tab == 0 ? ex : (tab == 1 ? ra * ex : ra * ra * ex);
      } else {
        const int kf = k & ~1;
        const double ar = md.rep[kf], b = md.imp[kf];
        const double ex = exp(ar * l);
        double sn, cs;
        sincos(b * l, &sn, &cs);
        const double r1 = md.rate, r2 = md.rate * md.rate;
        const double sg = (k & 1) ? 1.0 : -1.0;
        if (tab == 0) { dv = ex * cs; ov = sg * ex * sn; }
        else if (tab == 1) { dv = r1 * (ar * cs - b * sn) * ex; ov = sg * r1 * (ar * sn + b * cs) * ex; }
        else { dv = r2 * ((ar * ar - b * b) * cs - 2.0 * ar * b * sn) * ex; ov = sg * r2 * ((ar * ar - b * b) * sn + 2.0 * ar * b * cs) * ex; }
      }
      dtab[e] = rpow * dv;
      otab[e] = rpow * ov;
    }
    double acc[MBW][NB][2];
#pragma unroll
    for (int i = 0; i < MBW; ++i)
#pragma unroll
      for (int j = 0; j < NB; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int NKS = Sp / kPtKT;
#pragma unroll
    for (int s0 = 0; s0 < kPtStages - 1; ++s0) {
      if (s0 < NKS) stage(s0, s0);
      cp_async_commit();
    }
    for (int ks = 0; ks < NKS; ++ks) {
      const int buf = ks % kPtStages;
      cp_async_wait<kPtStages - 2>();
      __syncthreads();  // also publishes dtab / otab before their first use
      if (ks + kPtStages - 1 < NKS) stage(ks + kPtStages - 1, (ks + kPtStages - 1) % kPtStages);
      cp_async_commit();
      const double* a_src = As + buf * kAStage;
      const double* b_src = Bs + buf * kBStage;
#pragma unroll
      for (int kb = 0; kb < kPtKT / 4; ++kb) {
        const int k = ks * kPtKT + kb * 4 + q;
        double a[MBW], b[NB];
#pragma unroll
        for (int i = 0; i < MBW; ++i) {
          const double* ar = a_src + (mb[i] * 8 + g) * kPtAStride + kb * 4;
          double av = ar[q] * dtab[mloc[i] * Sp + k];
          if (cplx) av = fma(ar[q ^ 1], otab[mloc[i] * Sp + k], av);
          a[i] = av;
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) b[j] = b_src[(kb * 4 + q) * BStride + j * 8 + g];
#pragma unroll
        for (int i = 0; i < MBW; ++i)
#pragma unroll
          for (int j = 0; j < NB; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
    }
    cp_async_wait<0>();

    double* out = tab == 0 ? p.P : tab == 1 ? p.dP : p.d2P;
#pragma unroll
    for (int i = 0; i < MBW; ++i) {
      if (!valid[i]) continue;
      const int x = mb[i] * 8 + g;
      if (x >= S) continue;
      const size_t base = ((size_t)point * nmat_pt + mat0 + mloc[i]) * S * S;
      const double t = tau_s[mloc[i]];
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const int y = n0 + j * 8 + 2 * q;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (y + h >= S) continue;
          double val = acc[i][j][h];
          if (tab == 0) {
            if (t == 0.0) val = x == y + h ? 1.0 : 0.0;
            if (chr_deriv && p.Pun) p.Pun[base + (size_t)x * S + y + h] = val;
            if (clamp) val = val < 0.0 ? 1e-20 : (val > 1.0 ? 1.0 : val);
          }
          out[base + (size_t)x * S + y + h] = val;
        }
      }
    }
  }
}

}  // namespace bppgpu
