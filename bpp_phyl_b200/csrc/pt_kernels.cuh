// K1: branch-batched transition-probability tables.
//
// One launch builds P, r_c*P', r_c^2*P'' for every (point, branch, rate class):
//   pxy_[b][c] = P(l_b*r_c), dpxy_ = r_c*P'(l_b*r_c), d2pxy_ = r_c^2*P''(l_b*r_c)
// (Likelihood/AbstractHomogeneousTreeLikelihood.cpp:354-414) from the HOST
// eigendecomposition of Q:
//   P   = V.diag(exp(lambda*rate*t)).V^-1                 Model/AbstractSubstitutionModel.cpp:436
//   P'  = V.diag(rate*lambda*exp(..)).V^-1                :505
//   P'' = V.diag((rate*lambda)^2*exp(..)).V^-1            :576
// Conjugate eigen-pairs use the real 2x2 block form (:438-468, :507-537, :578-612):
// V.T.V^-1 with T tridiagonal.  t == 0 gives the identity for P (:428-431).
//
// This file is the CUDA-core path used for small/medium S (one CTA per matrix,
// one thread per element group).  The S>=32 DMMA path is in gemm_kernels.cuh.
#pragma once
#include "common.cuh"
#include "dmma.cuh"

namespace bppgpu {

struct ModelDev {
  const double* V;     // [S][S] right eigenvectors (columns)
  const double* Vinv;  // [S][S] left eigenvectors (rows)
  const double* re;    // [S]
  const double* im;    // [S] (zeros when real)
  const int* role;     // [S] 0 real, 1 first of a conjugate pair, 2 second
  const double* Q;     // [S][S] generator or nullptr
  const double* Q2;    // [S][S] Q.Q or nullptr
  const double* Vp;    // [Sp][Sp] V zero-padded to Sp = roundup(S, 8) (== V when S % 8 == 0)
  const double* Vinvp; // [Sp][Sp]
  const double* rep;   // [Sp]
  const double* imp;   // [Sp] imaginary parts in the permuted order (pairs first, each at an even index)
  double rate;
  double eps;
  double q_l1;         // sum_ij |Q_ij| (ChromosomeSubstitutionModel::getFirstNorm)
  unsigned flags;
  int has_complex;
};

struct PtParams {
  const ModelDev* models;
  const int* branch_model;  // [npoints][nn]
  const double* brlen;      // [npoints][nn]
  const double* rates;      // [C]
  int S, C, nn, root;
  unsigned want;            // BPPGPU_WANT_*
  double* P;                // [npoints][nn][C][S][S]
  double* dP;
  double* d2P;
  double* Pun;              // [npoints][nn][C][S][S] unclamped P for CHR_DERIV models, or nullptr
  int dmma_real;            // 1: real-spectrum eigen models are built by pt_dmma_kernel, skip them here
};

// dynamic smem: 6*S doubles (dia/up for orders 0,1,2)
__global__ void pt_eigen_kernel(PtParams p) {
  extern __shared__ double sm[];
  const int S = p.S;
  double* dia0 = sm;
  double* up0 = sm + S;
  double* dia1 = sm + 2 * S;
  double* up1 = sm + 3 * S;
  double* dia2 = sm + 4 * S;
  double* up2 = sm + 5 * S;

  const int m = blockIdx.x;
  const int c = m % p.C;
  const int node = (m / p.C) % p.nn;
  const int point = m / (p.C * p.nn);
  if (node == p.root) return;
  const ModelDev md = p.models[p.branch_model[point * p.nn + node]];
  if (!((md.flags & 1u) || (md.flags & 2u))) return;  // handled by the series kernel
  if (!(md.flags & 2u)) return;                       // singular -> series kernel
  if (p.dmma_real) return;                            // built on the tensor cores (pt_dmma_kernels.cuh)
  const double rc = p.rates[c];
  const double t = p.brlen[point * p.nn + node] * rc;  // l_b * r_c
  const double l = md.rate * t;

  for (int k = threadIdx.x; k < S; k += blockDim.x) {
    const double a = md.re[k];
    const int role = md.has_complex ? md.role[k] : 0;
    if (role == 0) {
      const double ex = exp(a * l);
      const double ra = md.rate * a;
      dia0[k] = ex;
      dia1[k] = rc * (ra * ex);
      dia2[k] = rc * rc * (ra * ra * ex);
      up0[k] = up1[k] = up2[k] = 0.0;
    } else {
      const int kf = role == 1 ? k : k - 1;  // first member holds +im
      const double ar = md.re[kf], b = md.im[kf];
      const double ex = exp(ar * l);
      double s, cc;
      sincos(b * l, &s, &cc);
      const double r1 = md.rate, r2 = md.rate * md.rate;
      dia0[k] = ex * cc;
      dia1[k] = rc * (r1 * (ar * cc - b * s) * ex);
      dia2[k] = rc * rc * (r2 * ((ar * ar - b * b) * cc - 2.0 * ar * b * s) * ex);
      // T[kf][kf+1] = up, T[kf+1][kf] = -up ; stored at the FIRST index only
      up0[k] = role == 1 ? ex * s : 0.0;
      up1[k] = role == 1 ? rc * (r1 * (ar * s + b * cc) * ex) : 0.0;
      up2[k] = role == 1 ? rc * rc * (r2 * ((ar * ar - b * b) * s + 2.0 * ar * b * cc) * ex) : 0.0;
    }
  }
  __syncthreads();

  const size_t base = (size_t)m * S * S;
  const bool wP = p.want & 1u, wD = p.want & 2u, wD2 = p.want & 4u;
  const bool clamp = md.flags & 4u;
  const bool chr_deriv = md.flags & 8u;  // dP, d2P rebuilt from the unclamped P by pt_chr_deriv_kernel
  for (int e = threadIdx.x; e < S * S; e += blockDim.x) {
    const int x = e / S, y = e - x * S;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    if (!md.has_complex) {
      for (int k = 0; k < S; ++k) {
        const double vv = md.V[x * S + k] * md.Vinv[k * S + y];
        a0 = fma(vv, dia0[k], a0);
        a1 = fma(vv, dia1[k], a1);
        a2 = fma(vv, dia2[k], a2);
      }
    } else {
      for (int k = 0; k < S; ++k) {
        const int role = md.role[k];
        const double vk = md.V[x * S + k];
        double w0 = vk * dia0[k], w1 = vk * dia1[k], w2 = vk * dia2[k];
        if (role == 2) {  // T[k-1][k] = up[k-1]
          const double vm = md.V[x * S + k - 1];
          w0 = fma(vm, up0[k - 1], w0);
          w1 = fma(vm, up1[k - 1], w1);
          w2 = fma(vm, up2[k - 1], w2);
        } else if (role == 1) {  // T[k+1][k] = -up[k]
          const double vp = md.V[x * S + k + 1];
          w0 = fma(vp, -up0[k], w0);
          w1 = fma(vp, -up1[k], w1);
          w2 = fma(vp, -up2[k], w2);
        }
        const double u = md.Vinv[k * S + y];
        a0 = fma(w0, u, a0);
        a1 = fma(w1, u, a1);
        a2 = fma(w2, u, a2);
      }
    }
    if (t == 0.0) a0 = (x == y) ? 1.0 : 0.0;  // :428-431
    if (chr_deriv && p.Pun) p.Pun[base + e] = a0;
    if (wP) {
      if (clamp) {  // ChromosomeSubstitutionModel.cpp:903-916
        if (a0 < 0.0) a0 = 1e-20;
        else if (a0 > 1.0) a0 = 1.0;
      }
      p.P[base + e] = a0;
    }
    if (wD && !chr_deriv) p.dP[base + e] = a1;
    if (wD2 && !chr_deriv) p.d2P[base + e] = a2;
  }
}

// ---- tip lookup tables --------------------------------------------------------
// tiptab[point][leaf][c][code][x] = sum_y P_leaf[c][x][y] * code_table[code][y]
// i.e. the son contraction of RHomogeneousTreeLikelihood.cpp:851-856 evaluated once
// per distinct tip code instead of once per pattern (tips are never expanded to S
// doubles).
struct TipTabParams {
  const double* P;           // [npoints][nn][C][S][S]
  const double* code_table;  // [ncodes][S]
  const int* leaf_nodes;     // [nl] node id of each leaf slot
  int S, C, nn, nl, ncodes;
  double* tiptab;            // [npoints][nl][C][ncodes][S]
};

__global__ void tiptab_kernel(TipTabParams p) {
  const int m = blockIdx.x;  // (point*nl + leaf)*C + c  (P may be pxy_, dpxy_ or d2pxy_)
  const int c = m % p.C;
  const int leaf = (m / p.C) % p.nl;
  const int point = m / (p.C * p.nl);
  const int S = p.S;
  const double* Pm = p.P + ((size_t)(point * p.nn + p.leaf_nodes[leaf]) * p.C + c) * S * S;
  double* out = p.tiptab + (size_t)m * p.ncodes * S;
  for (int e = threadIdx.x; e < p.ncodes * S; e += blockDim.x) {
    const int code = e / S, x = e - code * S;
    const double* tv = p.code_table + (size_t)code * S;
    double acc = 0.0;
    for (int y = 0; y < S; ++y) acc = fma(Pm[x * S + y], tv[y], acc);
    out[e] = acc;
  }
}

}  // namespace bppgpu

namespace bppgpu {

// ---- K1c: series + scaling-and-squaring for singular / non-diagonalisable Q -----
// Generic model (AbstractSubstitutionModel.cpp:470-492): v = rate*t halved m times until
// <= 0.5, P = sum_{k<30} v^k/k! Q^k, squared m times; derivatives rate*Q.P and rate^2*Q.Q.P
// (:539-563, :614-638).  Chromosome variant (ChromosomeSubstitutionModel.cpp:852-899,
// :934-946): halve until v*sum|Q_ij| <= 0.5, add terms until every |term_ij| <= eps and
// P within [-eps, 1+eps] (at least 3 terms, at most 250), square back; then
// dP = P.Q.rate (:966-982), d2P = Q^2.P.rate^2 (:986-1001) from the UNCLAMPED P, and the
// clamp P<0 -> 1e-20, P>1 -> 1 (:903-916) on P itself.  BPPGPU_MODEL_EXACT_EXPM runs the
// series to FP64 convergence instead of the reference's truncation.
//
// One CTA per matrix, operands in global scratch ([4][S][S] per matrix, L2-resident).  For S >= 32 every matrix product
// -- the Taylor terms and, above all, the m squarings (m ~ 10 for a Chromosome generator: the reference halves until
// t * sum|Q_ij| <= 0.5) -- runs on the FP64 tensor cores: mat_mul_dmma walks the right factor in 32-column panels staged in
// shared memory and takes the left factor's fragments straight from L2 (the skinny panel GEMM of dmma.cuh).  Smaller
// alphabets keep the CUDA-core loop.
struct SeriesParams {
  PtParams pt;
  double* scratch;  // [nmat][4][S*S]
  double q_l1_unused;
  int* status;      // set to 1 if the chromosome series did not converge
  int only_chr_deriv;  // 1: matrices of eigen-path models flagged CHR_DERIV: rebuild dP/d2P from P
  int sparse_terms;    // 1: Taylor terms against the compressed columns of Q (opt-in: measured slower than the tensor-core product)
};

__device__ __forceinline__ void mat_mul(const double* A, const double* B, double* O, int S, double scale) {
  for (int e = threadIdx.x; e < S * S; e += blockDim.x) {
    const int x = e / S, y = e - x * S;
    double acc = 0.0;
    for (int k = 0; k < S; ++k) acc = fma(A[x * S + k], B[k * S + y], acc);
    O[e] = acc * scale;
  }
  __syncthreads();
}

// O = A . B . scale on the tensor cores (blockDim = 256, dynamic shared memory: roundup(S, 8) x 36 doubles); A, B, O distinct.
// A and B may have been written by this CTA: plain (coherent) loads, and every thread passes the closing barrier.
__device__ __forceinline__ void mat_mul_dmma(const double* A, const double* B, double* O, int S, double scale, double* Bs) {
  const int K8 = (S + 7) & ~7, nrb = K8 >> 3;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
  for (int n0 = 0; n0 < S; n0 += kChrCols) {
    __syncthreads();   // the previous panel's fragments have been read
    for (int i = tid; i < K8 * kChrCols; i += blockDim.x) {
      const int k = i / kChrCols, j = i - k * kChrCols;
      Bs[k * kChrLD + j] = (k < S && n0 + j < S) ? __ldcg(B + (size_t)k * S + n0 + j) : 0.0;
    }
    __syncthreads();
    double acc[kChrMaxRB][kChrCols / 8][2];
    chr_gemm<true>(A, S, K8, Bs, nrb, warp, g, q, acc);
#pragma unroll
    for (int i = 0; i < kChrMaxRB; ++i)
      if (warp + i * kChrWarps < nrb) {
        const int row = (warp + i * kChrWarps) * 8 + g;
#pragma unroll
        for (int cb = 0; cb < kChrCols / 8; ++cb) {
          const int col = n0 + cb * 8 + 2 * q;
          if (row < S && col < S) O[(size_t)row * S + col] = acc[i][cb][0] * scale;
          if (row < S && col + 1 < S) O[(size_t)row * S + col + 1] = acc[i][cb][1] * scale;
        }
      }
  }
  __syncthreads();
}

// ---- Taylor terms against a SPARSE generator ----------------------------------------------------------------------------
// term_i = term_{i-1} . Q . v / i is a product with the generator on the right, and a Chromosome generator has at most a handful of
// non-zeros per column (the states that reach state c by one gain, loss, duplication or demi-duplication, and c itself): the
// product is ~6 S^2 multiply-adds, not S^3.  The CTA compresses the columns of Q once (kSparseMax entries per column; a denser
// column stays dense, and more than S / 8 of those make the whole model take the dense tensor-core product) and the series' first phase -- 3 to 6 terms of the ~10 products of
// a matrix -- runs on them; the squarings stay dense.  Sums run over k ascending like the dense loop (zeros skipped).
// MEASURED (cfg5, 33 966 matrices of 200 x 200): 245 ms against 231 ms with every product on the tensor cores -- the gathers of the
// left factor from L2 (five dependent-latency loads per output) cost more than the 43 % of DMMAs they remove.  Opt-in
// (BPPGPU_SERIES_SPARSE=1) until the left factor's rows are staged in shared memory.
constexpr int kSparseMax = 8;
struct SparseCols {
  int* cnt;      // [S]
  int* row;      // [S][kSparseMax]
  double* val;   // [S][kSparseMax]
};
__host__ __device__ inline size_t sparse_cols_bytes(int S) {
  return (size_t)S * kSparseMax * (sizeof(double) + sizeof(int)) + (size_t)((S + 1) & ~1) * sizeof(int);
}
// returns (uniformly over the CTA) whether every column fits; all threads call
__device__ __forceinline__ bool sparse_cols_build(const double* Q, int S, SparseCols sc) {
  __shared__ int too_dense;
  if (threadIdx.x == 0) too_dense = 0;
  __syncthreads();
  for (int c = threadIdx.x; c < S; c += blockDim.x) {
    int n = 0;
    for (int k = 0; k < S; ++k) {
      const double q = Q[(size_t)k * S + c];
      if (q != 0.0) {
        if (n < kSparseMax) {
          sc.row[c * kSparseMax + n] = k;
          sc.val[c * kSparseMax + n] = q;
        }
        ++n;
      }
    }
    // a column with more entries is kept dense (cnt = -1) and costs S multiply-adds per row: the LAST state of a Chromosome model
    // collects every duplication / demi-duplication that would overshoot the maximal count, i.e. about half of the rows
    sc.cnt[c] = n > kSparseMax ? -1 : n;
    if (n > kSparseMax) atomicAdd(&too_dense, 1);
  }
  __syncthreads();
  return too_dense * 8 <= S;   // at most S / 8 dense columns: the product stays well below S^3 / 8
}
// O = A . Q . scale with Q given by its compressed columns (A, O distinct; A may have been written by this CTA)
__device__ __forceinline__ void mat_mul_sparse_right(const double* A, const double* Q, SparseCols sc, double* O, int S, double scale) {
  for (int e = threadIdx.x; e < S * S; e += blockDim.x) {
    const int x = e / S, c = e - x * S;
    const int n = sc.cnt[c];
    double acc = 0.0;
    if (n >= 0) {
      for (int j = 0; j < n; ++j) acc = fma(__ldcg(A + (size_t)x * S + sc.row[c * kSparseMax + j]), sc.val[c * kSparseMax + j], acc);
    } else {
      for (int k = 0; k < S; ++k) acc = fma(__ldcg(A + (size_t)x * S + k), Q[(size_t)k * S + c], acc);
    }
    O[e] = acc * scale;
  }
  __syncthreads();
}

// (DMMA variant: two CTAs per SM -- one matrix's copy loops between the products wait on L2 while the other's products run:
//  ncu of the one-CTA version showed the DMMA pipe 34 % active with long-scoreboard stalls on the scratch copies)
template <bool DMMA>
__global__ void __launch_bounds__(DMMA ? 256 : 1024, DMMA ? 2 : 1) pt_series_kernel(SeriesParams sp) {
  extern __shared__ __align__(16) double sm_series[];
  auto mat_mul = [&](const double* A_, const double* B_, double* O_, int S_, double scale_) {
    if (DMMA) mat_mul_dmma(A_, B_, O_, S_, scale_, sm_series);
    else bppgpu::mat_mul(A_, B_, O_, S_, scale_);
  };
  const PtParams& p = sp.pt;
  __shared__ double red_max[32];
  __shared__ int flag;
  const int S = p.S;
  const int m = blockIdx.x;
  const int c = m % p.C;
  const int node = (m / p.C) % p.nn;
  const int point = m / (p.C * p.nn);
  if (node == p.root) return;
  const ModelDev md = p.models[p.branch_model[point * p.nn + node]];
  const bool eigen_path = (md.flags & 2u) != 0;  // NONSINGULAR
  const bool chr = (md.flags & 16u) != 0;        // CHR_TAYLOR
  const bool chr_deriv = (md.flags & 8u) != 0;
  if (eigen_path) return;  // built by pt_eigen_kernel (+ pt_chr_deriv_kernel)
  const double rc = p.rates[c];
  const double t = p.brlen[point * p.nn + node] * rc;
  double* T = sp.scratch + (size_t)m * 4 * S * S;  // current term
  double* A = T + S * S;                            // running sum / result
  double* B = A + S * S;                            // temp
  const double* Q = md.Q;
  const size_t base = (size_t)m * S * S;
  const bool wP = p.want & 1u, wD = p.want & 2u, wD2 = p.want & 4u;

  double v = md.rate * t;
  int k = 0;
  if (chr) {
    double norm = v * md.q_l1;
    while (norm > 0.5) { ++k; v *= 0.5; norm *= 0.5; }
  } else {
    while (v > 0.5) { ++k; v *= 0.5; }
  }
  for (int e = threadIdx.x; e < S * S; e += blockDim.x) {
    const double id = (e / S == e % S) ? 1.0 : 0.0;
    T[e] = id;
    A[e] = id;
  }
  __syncthreads();
  // compressed columns of Q behind the tensor-core product's panel (DMMA variant only: that is where the dynamic shared memory is)
  SparseCols sc{};
  bool sparse = false;
  if (DMMA && sp.sparse_terms && t != 0.0) {
    char* sb = reinterpret_cast<char*>(sm_series) + (size_t)((S + 7) & ~7) * kChrLD * sizeof(double);
    sc.val = reinterpret_cast<double*>(sb);
    sc.row = reinterpret_cast<int*>(sc.val + (size_t)S * kSparseMax);
    sc.cnt = sc.row + (size_t)S * kSparseMax;
    sparse = sparse_cols_build(Q, S, sc);
  }
  if (t != 0.0) {
    const bool exact = (md.flags & 32u) != 0;
    const int max_terms = chr || exact ? 250 : 29;
    for (int i = 1; i <= max_terms; ++i) {
      if (sparse) mat_mul_sparse_right(T, Q, sc, B, S, v / (double)i);
      else mat_mul(T, Q, B, S, v / (double)i);  // term_i = term_{i-1}.Q.v/i
      double mx = 0.0;
      int bad = 0;
      for (int e = threadIdx.x; e < S * S; e += blockDim.x) {
        const double tv = B[e];
        T[e] = tv;
        const double a = A[e] + tv;
        A[e] = a;
        mx = fmax(mx, fabs(tv));
        if (chr && !exact && (a + md.eps < 0.0 || a > 1.0 + md.eps)) bad = 1;
      }
      if (chr || exact) {
        // block max of |term| and OR of the range test
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
          bad |= __shfl_xor_sync(0xffffffffu, bad, o);
        }
        if (threadIdx.x == 0) flag = 0;
        __syncthreads();
        if ((threadIdx.x & 31) == 0) {
          red_max[threadIdx.x >> 5] = mx;
          if (bad) atomicOr(&flag, 1);
        }
        __syncthreads();
        double bm = 0.0;
        for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) bm = fmax(bm, red_max[w]);
        const int anybad = flag;
        __syncthreads();
        const double tol = exact ? 1e-18 : md.eps;
        if (i >= 3 && bm <= tol && !anybad) break;
        if (i == max_terms && !exact && threadIdx.x == 0) *sp.status = 1;
      } else {
        __syncthreads();
      }
    }
    for (int j = 0; j < k; ++j) {  // square back
      mat_mul(A, A, B, S, 1.0);
      for (int e = threadIdx.x; e < S * S; e += blockDim.x) A[e] = B[e];
      __syncthreads();
    }
  }
  // A = unclamped P
  if (wD) {
    if (chr_deriv) mat_mul(A, Q, B, S, md.rate * rc);  // P.Q.rate
    else mat_mul(Q, A, B, S, md.rate * rc);            // rate.Q.P
    for (int e = threadIdx.x; e < S * S; e += blockDim.x) p.dP[base + e] = B[e];
    __syncthreads();
  }
  if (wD2) {
    mat_mul(md.Q2, A, B, S, md.rate * md.rate * rc * rc);  // rate^2.Q^2.P
    for (int e = threadIdx.x; e < S * S; e += blockDim.x) p.d2P[base + e] = B[e];
    __syncthreads();
  }
  if (wP) {
    const bool clamp = md.flags & 4u;
    for (int e = threadIdx.x; e < S * S; e += blockDim.x) {
      double a = A[e];
      if (clamp) {
        if (a < 0.0) a = 1e-20;
        else if (a > 1.0) a = 1.0;
      }
      p.P[base + e] = a;
    }
  }
}

// Chromosome models on the eigen path: dP = Pun.Q.rate, d2P = Q^2.Pun.rate^2 where Pun is the
// UNCLAMPED P that pt_eigen_kernel left in `scratch` ([nmat][S*S]).
__global__ void pt_chr_deriv_kernel(SeriesParams sp) {
  const PtParams& p = sp.pt;
  const int S = p.S;
  const int m = blockIdx.x;
  const int c = m % p.C;
  const int node = (m / p.C) % p.nn;
  const int point = m / (p.C * p.nn);
  if (node == p.root) return;
  const ModelDev md = p.models[p.branch_model[point * p.nn + node]];
  if (!(md.flags & 2u) || !(md.flags & 8u)) return;
  const double rc = p.rates[c];
  const double* Pun = sp.scratch + (size_t)m * S * S;
  const size_t base = (size_t)m * S * S;
  if (p.want & 2u) {
    for (int e = threadIdx.x; e < S * S; e += blockDim.x) {
      const int x = e / S, y = e - x * S;
      double acc = 0.0;
      for (int k = 0; k < S; ++k) acc = fma(Pun[x * S + k], md.Q[k * S + y], acc);
      p.dP[base + e] = acc * md.rate * rc;
    }
  }
  if (p.want & 4u) {
    for (int e = threadIdx.x; e < S * S; e += blockDim.x) {
      const int x = e / S, y = e - x * S;
      double acc = 0.0;
      for (int k = 0; k < S; ++k) acc = fma(md.Q2[x * S + k], Pun[k * S + y], acc);
      p.d2P[base + e] = acc * md.rate * md.rate * rc * rc;
    }
  }
}

}  // namespace bppgpu
