// Batched-points path WITHOUT transition-probability tables (BASELINE config 5: ChromEvol-style chromosome-number models,
// S ~ 200 states, one character, thousands of parameter points on one tree).
//
// With one (pattern, class) row per point the pruning step of a branch is a matrix-VECTOR product, P_b . CLV_son
// (RHomogeneousTreeLikelihood.cpp:851-856; DRNonHomogeneousTreeLikelihood::computeSubtreeLikelihoodPostfix), and building
// P_b = V exp(L t_b) V^-1 (ChromosomeSubstitutionModel.cpp:808-851) first costs 2 S^3 flop and S^2 doubles of HBM traffic per
// branch for the sake of 2 S^2 useful flop.  Here the product is applied in factored form,
//     term_b = V . ( T(t_b) . ( V^-1 . x_b ) ),        T = exp(L t) with the 2x2 rotation blocks of conjugate pairs
// (the reference's own block form, :821-850), and the branches of one tree LEVEL (sons of equal subtree height) are the
// columns of two skinny GEMMs per point on the FP64 tensor cores (mma.sync m8n8k4): A = V^-1 / V straight from L2 (a point's
// 2 S^2 doubles are re-used by every column tile of every level), B = the level's column tile in shared memory.
// Tip sons with an observed count k need no first GEMM: V^-1 e_k is a column of V^-1.
//   flops per point ~ 4 S^2 (#internal sons) + 2 S^2 (#tips)  vs  2 S^3 (#branches):  ~ 100x fewer at S = 200.
//
// What this route cannot reproduce is the reference's per-ENTRY clamp of P (P < 0 -> 1e-20, P > 1 -> 1, :903-916): it never
// sees the entries.  The engine therefore evaluates a guard per point (P at the shortest, the geometric-mean and the longest
// branch through the tensor-core table kernel, entries outside [-1e-8, 1 + 1e-8] counted) and sends points that fail it,
// and singular generators (Taylor route), through the table path (points_kernels.cuh).
#pragma once
#include "dmma.cuh"
#include "pt_kernels.cuh"

namespace bppgpu {

// Guard: a point takes the factored route when no entry of its probe tables leaves [0, 1] by more than this; entries that do
// are the ones the reference clamps (ChromosomeSubstitutionModel.cpp:903-916).  Observed tips are exact whatever the guard says
// (their term is a column of P and is clamped entry by entry); the guard is about internal branches, where the clamp cannot be
// applied.  Measured on the benchmark's points (S = 200, 500 taxa; tools/chr_guard_sweep.py, profiles/r2_chr_guard_sweep.json):
// with violations up to 1e-8 the factored lnL is within 1e-13 relative of the table route's, i.e. four orders below the 1e-9
// parity bar, so 1e-8 is the threshold: tol 1e-12 sends 74 % of those points to the tables, 1e-10 12 %, 1e-9 0.2 %, 1e-8 none.
constexpr double kChrGuardTol = 1e-8;

struct ChrLevelParams {
  const ModelDev* models;
  const int* branch_model;   // [npts][nn] (homogeneous points: every branch of a point has the same slot)
  const double* brlen;       // [npts][nn]
  double rate0;              // the single rate class
  int S, nn, p0;             // p0: first point of this launch (blockIdx.y is relative to it)
  int tile0;                 // first tile of the level
  const int* tile_edges;     // [ntiles][kChrCols] son node ids, -1 = unused column
  const int* tile_kind;      // [ntiles] 0: tip sons with an observed state (first GEMM skipped), 1: dense columns
  const int* child_off;      // CSR sons of every node
  const int* children;
  const int* leaf_state;     // [nn] observed state of a leaf, -1 = dense leaf (leaf_vec), -2 = internal node
  const double* leaf_vec;    // [nn][S] getInitValue row of a dense leaf (ambiguity / unknown), unused otherwise
  double* term;              // [points of this launch][nn][S] term of the branch above node n, rescaled
  int* term_exp;             // [points of this launch][nn]
  const int* skip;           // [npts] 1 = the point goes through the table route (guard), or nullptr
  const double* aslab;       // [model slot][V^-1 | V | (V^-1)^T] slab-ordered copies + transpose (chr_slab_kernel), slab-streamed kernels only
  int nst;                   // stages of the slab ring
};

// per-tile scalars behind the column tile: tl, cmax, cscale [32] doubles; cexp, cnode [32] ints; the sons of every column
// [32][4] ints (count + up to 3 ids); the point's eigenvalues re, im [K8] doubles and role [K8] ints
__host__ __device__ inline size_t chr_meta_bytes(int S) {
  const int K8 = (S + 7) & ~7;
  return 3 * kChrCols * sizeof(double) + 2 * kChrCols * sizeof(int) + 4 * kChrCols * sizeof(int) + (size_t)K8 * (2 * sizeof(double) + sizeof(int));
}
inline size_t chr_level_smem(int S) {
  const int K4 = (S + 7) & ~7;
  return (size_t)K4 * kChrLD * sizeof(double) + chr_meta_bytes(S) +
         (size_t)kChrWarps * kChrRingDoubles * sizeof(double);   // + every warp's ring of A fragments
}

// shared memory of the slab-streamed kernels: the column tile and its per-column scalars as above, then the ring's barriers and slabs
inline size_t chr_slab_smem(int S, int nst) {
  const int K8 = (S + 7) & ~7;
  return (size_t)K8 * kChrLD * sizeof(double) + chr_meta_bytes(S) + 2 * 16 * sizeof(unsigned long long) +
         (size_t)nst * chr_slab_doubles(K8) * sizeof(double);
}
constexpr int kChrMaxStages = 16;

// barrier over the threads that work on the tile: the whole CTA, or the consumer warps of a slab-streamed kernel (named barrier 1;
// the producer warp is not part of it)
template <bool SLAB>
__device__ __forceinline__ void chr_sync() {
  if (SLAB) asm volatile("bar.sync 1, %0;" ::"n"(kChrCons * 32) : "memory");
  else __syncthreads();
}

// one tile (<= 32 sons of one level) of one point; every (consumer) thread of the CTA calls it
// `cached_slot`: the model slot whose eigenvalues sit in shared memory (a chain CTA keeps them over its tiles), -1 = none
template <bool SLAB>
__device__ __forceinline__ void chr_tile(const ChrLevelParams& p, double* sm_chr, int tile, int prel, ChrRing* rg, int& cached_slot) {
  constexpr int NW = SLAB ? kChrCons : kChrWarps;
  constexpr int NT = NW * 32;
  const int S = p.S;
  const int K4 = (S + 7) & ~7;                 // rows of the shared tiles (multiple of 8: row blocks and k-steps)
  double* Xs = sm_chr;                         // [K4][LD]  the column tile: x, then V^-1 x, then T V^-1 x, then the terms (in place:
  double* Ws = Xs;                             //           every GEMM reads it completely before its result is stored back)
  double* tl = Xs + (size_t)K4 * kChrLD;       // [32] rate * t of the column's branch
  double* cmax = tl + kChrCols;                // [32] column maxima
  double* cscale = cmax + kChrCols;            // [32]
  int* cexp = reinterpret_cast<int*>(cscale + kChrCols);   // [32] exponent carried in by the column
  int* cnode = cexp + kChrCols;                // [32]
  int* cch = cnode + kChrCols;                 // [32][4] sons of the column's node: count, then up to 3 ids (more: read from the CSR lists)
  double* sre = reinterpret_cast<double*>(cch + 4 * kChrCols);   // [K4] eigenvalues of the point's model: real parts,
  double* sim = sre + K4;                                        //      imaginary parts,
  int* srole = reinterpret_cast<int*>(sim + K4);                 //      role (0 real, 1 / 2 the members of a conjugate pair)
  double* ring = SLAB ? nullptr : reinterpret_cast<double*>(reinterpret_cast<char*>(tl) + chr_meta_bytes(S)) + (size_t)(threadIdx.x >> 5) * kChrRingDoubles;   // this warp's A ring

  const int pt = p.p0 + prel;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
  const int* edges = p.tile_edges + (size_t)tile * kChrCols;
  const int kind = p.tile_kind[tile];
  const int first = edges[0];
  const int slot = p.branch_model[(size_t)pt * p.nn + first];
  const ModelDev md = p.models[slot];
  const int nrb = K4 >> 3;
  double* term_pt = p.term + (size_t)prel * p.nn * S;
  int* texp_pt = p.term_exp + (size_t)prel * p.nn;

  if (tid < kChrCols) {
    const int n = edges[tid];
    cnode[tid] = n;
    tl[tid] = n >= 0 ? md.rate * p.rate0 * p.brlen[(size_t)pt * p.nn + n] : 0.0;
    cexp[tid] = 0;
    // the sons of a dense column, once per tile (the products below then issue their term loads back to back instead of walking
    // child_off -> children -> term for every element)
    int cnt = 0;
    if (n >= 0 && kind != 0 && p.leaf_state[n] == -2) {
      const int c0 = p.child_off[n];
      cnt = p.child_off[n + 1] - c0;
      for (int c = 0; c < 3 && c < cnt; ++c) cch[4 * tid + 1 + c] = p.children[c0 + c];
    }
    cch[4 * tid] = cnt;
  }
  if (slot != cached_slot) {   // (uniform over the CTA) this point's eigenvalues: one coalesced pass, read many times below
    for (int k = tid; k < K4; k += NT) {
      sre[k] = k < S ? md.re[k] : 0.0;
      sim[k] = k < S ? md.im[k] : 0.0;
      srole[k] = k < S ? md.role[k] : 0;
    }
    cached_slot = slot;
  }
  // columns in use (a tile is filled from column 0): only their 8-column blocks go through the tensor cores, and the
  // element-wise phases stop at the last block in use.  Most levels of a tree hold a handful of branches.
  const int ncb = (__popc(__ballot_sync(0xffffffffu, edges[lane] >= 0)) + 7) >> 3;
  const int ncols = ncb * 8;
  chr_sync<SLAB>();

  if (kind == 0) {
    // observed tips: W[:, j] = V^-1[:, state_j]
    // (the tiles of observed tips are sorted by state: neighbouring columns read neighbouring entries of a row of V^-1)
    if (SLAB) {
      // from the TRANSPOSED copy kept next to the slab images: a column of V^-1 is a contiguous row there (coalesced), instead of
      // S loads 8 S bytes apart
      const double* vt = p.aslab + (size_t)slot * 3 * K4 * K4 + (size_t)2 * K4 * K4;
      for (int i = tid; i < K4 * ncols; i += NT) {
        const int j = i / K4, k = i - j * K4;
        const int n = cnode[j];
        Ws[k * kChrLD + j] = n >= 0 ? __ldg(vt + (size_t)p.leaf_state[n] * K4 + k) : 0.0;
      }
    } else {
      for (int i = tid; i < K4 * ncols; i += NT) {
        const int k = i / ncols, j = i - k * ncols;
        const int n = cnode[j];
        Ws[k * kChrLD + j] = (n >= 0 && k < S) ? __ldg(md.Vinv + (size_t)k * S + p.leaf_state[n]) : 0.0;
      }
    }
  } else {
    // dense columns: x = the son's conditional likelihoods = product of ITS sons' terms (or a dense leaf's init row)
    for (int i = tid; i < K4 * ncols; i += NT) {
      const int j = i / K4, k = i - j * K4;      // consecutive threads read one son's term contiguously
      const int n = cnode[j];
      double v = 0.0;
      if (n >= 0 && k < S) {
        const int cnt = cch[4 * j];
        if (cnt == 0) v = p.leaf_vec[(size_t)n * S + k];   // dense leaf (unknown / ambiguous count)
        else {
          v = term_pt[(size_t)cch[4 * j + 1] * S + k];    // (same order of the factors as the CSR list)
          if (cnt > 1) v *= term_pt[(size_t)cch[4 * j + 2] * S + k];
          if (cnt > 2) v *= term_pt[(size_t)cch[4 * j + 3] * S + k];
          for (int c = p.child_off[n] + 3; c < p.child_off[n] + cnt; ++c) v *= term_pt[(size_t)p.children[c] * S + k];
        }
      }
      Xs[k * kChrLD + j] = v;
    }
    if (tid < kChrCols) {
      const int n = cnode[tid];
      int e = 0;
      if (n >= 0 && p.leaf_state[n] == -2)
        for (int c = p.child_off[n]; c < p.child_off[n + 1]; ++c) e += texp_pt[p.children[c]];
      cexp[tid] = e;
    }
    chr_sync<SLAB>();
    // the product of several rescaled terms may be small again: bring every column's maximum back to [0.5, 1)
    {
      for (int j = warp; j < ncols; j += NW) {
        double m = 0.0;
        for (int k = lane; k < S; k += 32) m = fmax(m, fabs(Xs[k * kChrLD + j]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) {
          int sh = 0;
          const int hi = hi_word(m);
          if (hi < kScaleThresholdHi && hi >= (1 << 20)) sh = rescale_shift(hi);
          cscale[j] = pow2(sh);
          cexp[j] += sh;
        }
      }
    }
    chr_sync<SLAB>();
    for (int i = tid; i < S * ncols; i += NT) {
      const int k = i / ncols, j = i - k * ncols;
      Xs[k * kChrLD + j] *= cscale[j];
    }
    chr_sync<SLAB>();
    double acc[kChrMaxRB][kChrCols / 8][2];
    if (SLAB) chr_gemm_slab_ncb(ncb, *rg, K4, Xs, nrb, warp, lane, acc);
    else chr_gemm_ncb(ncb, md.Vinv, S, K4, Xs, ring, nrb, warp, lane, acc);
    chr_sync<SLAB>();
    chr_store_acc_w<NW>(Ws, nrb, warp, g, q, acc, ncb);
  }
  chr_sync<SLAB>();
  // T(t): exp(re l) on real eigenvalues, the rotation block on conjugate pairs (ChromosomeSubstitutionModel.cpp:821-850)
  for (int i = tid; i < S * ncols; i += NT) {
    const int k = i / ncols, j = i - k * ncols;
    const int role = srole[k];
    const double l = tl[j];
    if (role == 0) {
      Ws[k * kChrLD + j] *= exp(sre[k] * l);
    } else if (role == 1) {
      const double ex = exp(sre[k] * l);
      double sn, cs;
      sincos(sim[k] * l, &sn, &cs);
      const double w0 = Ws[k * kChrLD + j], w1 = Ws[(k + 1) * kChrLD + j];
      Ws[k * kChrLD + j] = ex * (cs * w0 + sn * w1);
      Ws[(k + 1) * kChrLD + j] = ex * (cs * w1 - sn * w0);
    }
  }
  chr_sync<SLAB>();
  {
    double acc[kChrMaxRB][kChrCols / 8][2];
    if (SLAB) chr_gemm_slab_ncb(ncb, *rg, K4, Ws, nrb, warp, lane, acc);
    else chr_gemm_ncb(ncb, md.V, S, K4, Ws, ring, nrb, warp, lane, acc);
    chr_sync<SLAB>();
    chr_store_acc_w<NW>(Xs, nrb, warp, g, q, acc, ncb);
  }
  chr_sync<SLAB>();
  // rescale every term to [0.5, 1) by an exact power of two and store it with its exponent
  {
    for (int j = warp; j < ncols; j += NW) {
      double m = 0.0;
      for (int k = lane; k < S; k += 32) m = fmax(m, fabs(Xs[k * kChrLD + j]));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
      if (lane == 0) {
        int sh = 0;
        const int hi = hi_word(m);
        if (hi < kScaleThresholdHi && hi >= (1 << 20)) sh = rescale_shift(hi);
        cscale[j] = pow2(sh);
        const int n = cnode[j];
        if (n >= 0) texp_pt[n] = cexp[j] + sh;
      }
    }
  }
  chr_sync<SLAB>();
  // an observed tip's term IS a column of P: the reference's per-entry clamp (ChromosomeSubstitutionModel.cpp:903-916) applies
  // to it exactly; cscale is 1 there unless the whole column is below 2^-256
  const bool clamp_col = kind == 0 && (md.flags & 4u);
  for (int i = tid; i < S * ncols; i += NT) {
    const int j = i / S, k = i - j * S;     // consecutive threads write one term contiguously
    const int n = cnode[j];
    if (n < 0) continue;
    double v = Xs[k * kChrLD + j];
    if (clamp_col) v = v < 0.0 ? 1e-20 : (v > 1.0 ? 1.0 : v);
    term_pt[(size_t)n * S + k] = v * cscale[j];
  }
}

// one launch per tree level: grid (tiles of the level, points)
__global__ void __launch_bounds__(kChrWarps * 32, 2) chr_level_kernel(ChrLevelParams p) {
  extern __shared__ __align__(16) double sm_chr[];
  if (p.skip && p.skip[p.p0 + blockIdx.y]) return;
  int cached = -1;
  chr_tile<false>(p, sm_chr, p.tile0 + blockIdx.x, blockIdx.y, nullptr, cached);
}

// The top of the tree in ONE launch: from the first level on which every later level is a single tile (a handful of branches
// each -- 43 of the 47 levels of the 500-taxon benchmark tree) a CTA keeps its point and walks the levels itself.  Level-by-level
// launches stream the V and V^-1 of ALL points (2.6 GB at 4096 points x 200 states) from HBM once per level; here the launch is
// sized to one CTA per SM (p.chain_smem), so the eigenvectors of the 148 resident points (95 MB) stay in the 126 MB L2 while
// their CTA goes up the tree, and HBM sees each point once.  The terms written by one level are read back by the same CTA after
// a block barrier.
__global__ void __launch_bounds__(kChrWarps * 32, 1) chr_chain_kernel(ChrLevelParams p, int ntiles) {
  extern __shared__ __align__(16) double sm_chr[];
  if (p.skip && p.skip[p.p0 + blockIdx.x]) return;
  int cached = -1;
  for (int t = 0; t < ntiles; ++t) {
    chr_tile<false>(p, sm_chr, p.tile0 + t, blockIdx.x, nullptr, cached);
    __syncthreads();
  }
}

// ---- slab-streamed variants (dmma.cuh: chr_gemm_slab): 7 consumer warps + 1 producer warp ------------------------------------
__device__ __forceinline__ ChrRing chr_ring_setup(const ChrLevelParams& p, double* sm_chr) {
  const int K8 = (p.S + 7) & ~7;
  ChrRing r;
  r.full = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(sm_chr) + (size_t)K8 * kChrLD * sizeof(double) + chr_meta_bytes(p.S));
  r.empty = r.full + 16;
  r.slab = reinterpret_cast<double*>(r.empty + 16);
  r.nst = p.nst;
  r.stage = 0;
  r.phase = 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.nst; ++i) {
      mbar_init(r.full + i, 1);
      mbar_init(r.empty + i, kChrCons);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();   // (the only CTA-wide barrier: the producer warp leaves the kernel on its own)
  return r;
}
// the slabs tile `tile` of point `prel` consumes, in its order: [V^-1 | V] for dense columns, V alone for observed tips
__device__ __forceinline__ void chr_produce_tile(const ChrLevelParams& p, ChrRing& r, int tile, int prel) {
  const int K8 = (p.S + 7) & ~7;
  const int first = p.tile_edges[(size_t)tile * kChrCols];
  const int slot = p.branch_model[(size_t)(p.p0 + prel) * p.nn + first];
  const bool tips = p.tile_kind[tile] == 0;
  const double* src = p.aslab + (size_t)slot * 3 * K8 * K8 + (tips ? (size_t)K8 * K8 : 0);
  chr_ring_produce(r, src, (tips ? 1 : 2) * (K8 >> 3), K8);
}
__global__ void __launch_bounds__((kChrCons + 1) * 32, 2) chr_level_slab_kernel(ChrLevelParams p) {
  extern __shared__ __align__(16) double sm_chr[];
  if (p.skip && p.skip[p.p0 + blockIdx.y]) return;
  ChrRing r = chr_ring_setup(p, sm_chr);
  if ((threadIdx.x >> 5) == kChrCons) {
    if ((threadIdx.x & 31) == 0) chr_produce_tile(p, r, p.tile0 + blockIdx.x, blockIdx.y);
    return;
  }
  int cached = -1;
  chr_tile<true>(p, sm_chr, p.tile0 + blockIdx.x, blockIdx.y, &r, cached);
}
// MINB = CTAs per SM the register budget allows: 2 interleaves two points per SM (the element-wise phases between the products
// are latency-bound: one CTA leaves the SM idle in them), 1 keeps the 148 resident points' eigenvectors inside L2
template <int MINB>
__global__ void __launch_bounds__((kChrCons + 1) * 32, MINB) chr_chain_slab_kernel(ChrLevelParams p, int ntiles) {
  extern __shared__ __align__(16) double sm_chr[];
  if (p.skip && p.skip[p.p0 + blockIdx.x]) return;
  ChrRing r = chr_ring_setup(p, sm_chr);
  if ((threadIdx.x >> 5) == kChrCons) {
    if ((threadIdx.x & 31) == 0)
      for (int t = 0; t < ntiles; ++t) chr_produce_tile(p, r, p.tile0 + t, blockIdx.x);
    return;
  }
  int cached = -1;
  for (int t = 0; t < ntiles; ++t) {
    chr_tile<true>(p, sm_chr, p.tile0 + t, blockIdx.x, &r, cached);
    chr_sync<true>();
  }
}

// slab-ordered copies of a model's V^-1 and V (layout: dmma.cuh, chr_gemm_slab): out[slot] = [V^-1 image | V image | (V^-1)^T], the
// images [K8 / 8 slabs][K8 rows][8 swizzled k-columns], zero-padded; the transpose [K8 states][K8] (row s = column s of V^-1: what an
// observed tip with state s starts from).  One CTA row per model slot, built when the models change.
__global__ void chr_slab_kernel(const ModelDev* models, int S, int K8, double* out) {
  const ModelDev md = models[blockIdx.y];
  if (md.V == nullptr || md.Vinv == nullptr) return;
  double* o = out + (size_t)blockIdx.y * 3 * K8 * K8;
  const int per = K8 * K8;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 3 * per; idx += gridDim.x * blockDim.x) {
    const int mat = idx / per, e = idx - mat * per;
    if (mat == 2) {
      const int st = e / K8, k = e - st * K8;
      o[idx] = (st < S && k < S) ? md.Vinv[(size_t)k * S + st] : 0.0;
      continue;
    }
    const int ks = e / (K8 * 8), rem = e - ks * (K8 * 8), r = rem >> 3, pos = rem & 7;
    const int k = ks * 8 + (pos ^ (4 * ((r >> 1) & 1)));
    const double* M = mat ? md.V : md.Vinv;
    o[idx] = (r < S && k < S) ? M[(size_t)r * S + k] : 0.0;
  }
}

// root of every point: CLV_root = product of the root's sons' terms; weighted root frequencies (setWeightedRootFreq,
// DRNonHomogeneousTreeLikelihood.cpp:927-962: with one site, freq_x = L_x / sum L, 1 / S when all are zero) or the given ones;
// lnL = log sum_x freq_x CLV_root[x] - E ln 2.
struct ChrRootParams {
  int S, nn, root, p0;
  unsigned flags;   // bit1: weighted root frequencies
  const int* child_off;
  const int* children;
  const double* term;
  const int* term_exp;
  const int* skip;             // as in ChrLevelParams
  const double* rootfreq_in;   // [npts][S]
  double* rootfreq_used;       // [npts][S]
  const double* weights;       // [1]
  double* site_lnl;            // [npts]
  double* out;                 // [npts][stride]
  int out_stride;
};
__global__ void chr_root_kernel(ChrRootParams p) {
  extern __shared__ double sm_root2[];   // [S] root CLV
  __shared__ double red[32];
  const int pt = p.p0 + blockIdx.x, S = p.S;
  if (p.skip && p.skip[pt]) return;
  const double* term_pt = p.term + (size_t)blockIdx.x * p.nn * S;
  int E = 0;
  for (int c = p.child_off[p.root]; c < p.child_off[p.root + 1]; ++c) E += p.term_exp[(size_t)blockIdx.x * p.nn + p.children[c]];
  double tot = 0.0;
  for (int x = threadIdx.x; x < S; x += blockDim.x) {
    double v = 1.0;
    for (int c = p.child_off[p.root]; c < p.child_off[p.root + 1]; ++c) v *= term_pt[(size_t)p.children[c] * S + x];
    sm_root2[x] = v;
    tot += v;
  }
  tot = block_sum(tot, red);
  __shared__ double tot_s;
  if (threadIdx.x == 0) tot_s = tot;
  __syncthreads();
  double acc = 0.0;
  for (int x = threadIdx.x; x < S; x += blockDim.x) {
    const double f = (p.flags & 2u) ? (tot_s == 0.0 ? 1.0 / S : sm_root2[x] / tot_s) : p.rootfreq_in[(size_t)pt * S + x];
    p.rootfreq_used[(size_t)pt * S + x] = f;
    acc += f * sm_root2[x];
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) {
    double L = acc;
    if (L < 0) L = 0.0;
    const double lnl = log(L) - (double)E * kLn2;
    p.site_lnl[pt] = lnl;
    p.out[(size_t)pt * p.out_stride] = p.weights[0] * lnl;
  }
}

// rows of per-point arrays gathered to / scattered from the compact list of the points the guard sends to the table route
template <typename T>
__global__ void gather_rows_kernel(const T* src, T* dst, const int* idx, int row_len) {
  const int r = blockIdx.x;
  for (int i = threadIdx.x; i < row_len; i += blockDim.x) dst[(size_t)r * row_len + i] = src[(size_t)idx[r] * row_len + i];
}
template <typename T>
__global__ void scatter_rows_kernel(const T* src, T* dst, const int* idx, int row_len, int copy_len) {
  const int r = blockIdx.x;
  for (int i = threadIdx.x; i < copy_len; i += blockDim.x) dst[(size_t)idx[r] * row_len + i] = src[(size_t)r * row_len + i];
}

// guard of the factored route: entries of the P tables of a few probe branches outside [-tol, 1 + tol], per point
__global__ void chr_guard_kernel(const double* P, int nprobe, int S, double tol, int* bad /*[npts]*/, int p0) {
  const int pt = blockIdx.x;
  const double* Pp = P + (size_t)pt * nprobe * S * S;
  int b = 0;
  for (size_t i = threadIdx.x; i < (size_t)nprobe * S * S; i += blockDim.x) {
    const double v = Pp[i];
    if (!(v >= -tol && v <= 1.0 + tol)) b = 1;
  }
  if (__syncthreads_or(b) && threadIdx.x == 0) bad[p0 + pt] = 1;
}

}  // namespace bppgpu
