// Batched-points path (SURVEY.md 2.2 "batch over parameter points", config 5): many independent parameter points, each
// with its own model and therefore its own P(t) tables, evaluated together on one tree and one (small) pattern set --
// what ChromosomeNumberOptimizer does serially with a vector of likelihood objects
// (Likelihood/ChromosomeNumberOptimizer.cpp:141-153).  With one character per taxon the pruning step of a point is a
// chain of matrix-VECTOR products, P_son . CLV_son (RHomogeneousTreeLikelihood.cpp:851-856), so it is bound by streaming
// that point's P tables: one launch per tree node covers every point of the chunk, a CTA per (point, row tile), rows of P
// read fully coalesced by a warp per output state with a shuffle reduction.
#pragma once
#include "common.cuh"
#include "walk_kernels.cuh"

namespace bppgpu {

struct PointsNodeParams {
  const Child* childs;  // sons (kinds TIP / KEEP)
  int nchild;
  int out_idx;
  int S, C, nn, nl, ni, ncodes, code_bytes;
  long long N;
  const double* P;       // [npts][nn][C][S][S]
  const double* code_table;  // [ncodes][S] getInitValue rows
  const int* code_single;    // [ncodes] the state of an indicator row, or -1 (ambiguity / probabilities)
  const void* codes;     // [nl][N]
  double* keep;          // [npts][ni][N][C][S]
  int* keep_exp;         // [npts][ni][N][C]
};

// dynamic smem: (2*S + 32) doubles
__global__ void points_node_kernel(PointsNodeParams p) {
  extern __shared__ double sm_pts[];
  const int S = p.S, C = p.C;
  double* lrow = sm_pts;      // son's CLV row
  double* prod = sm_pts + S;  // running Hadamard product
  __shared__ int smax;
  const int pt = blockIdx.x;
  const long long rows = p.N * C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const size_t clv = (size_t)rows * S;
  double* keep_pt = p.keep + (size_t)pt * p.ni * clv;
  int* exp_pt = p.keep_exp + (size_t)pt * p.ni * rows;
  const double* P_pt = p.P + (size_t)pt * p.nn * C * S * S;
  for (long long rc = blockIdx.y; rc < rows; rc += gridDim.y) {
    const int c = (int)(rc % C);
    const long long pat = rc / C;
    int Ea = 0;
    for (int j = 0; j < p.nchild; ++j) {
      const Child ch = p.childs[j];
      __syncthreads();
      const double* Pm = P_pt + ((size_t)ch.pnode * C + c) * S * S;
      int single = -1;
      const double* l;
      if (ch.kind == CHILD_TIP) {
        const int code = load_code(p.codes, p.code_bytes, (long long)ch.idx * p.N + pat);
        single = p.code_single[code];
        l = p.code_table + (size_t)code * S;
      } else {
        l = keep_pt + ((size_t)ch.idx * rows + rc) * S;
        Ea += exp_pt[(size_t)ch.idx * rows + rc];
      }
      if (single >= 0) {
        // observed state y0: the son's term is column y0 of P (RHomogeneousTreeLikelihood.cpp:851-856 with a 0/1 tip vector)
        for (int x = threadIdx.x; x < S; x += blockDim.x) {
          const double t = Pm[(size_t)x * S + single];
          prod[x] = j == 0 ? t : prod[x] * t;
        }
      } else {
        for (int y = threadIdx.x; y < S; y += blockDim.x) lrow[y] = l[y];
        __syncthreads();
        for (int x = warp; x < S; x += nwarps) {
          const double* Pr = Pm + (size_t)x * S;
          double acc = 0.0;
          for (int y = lane; y < S; y += 32) acc = fma(Pr[y], lrow[y], acc);
          acc = warp_sum(acc);
          if (lane == 0) prod[x] = j == 0 ? acc : prod[x] * acc;
        }
      }
    }
    __syncthreads();
    // row maximum -> power-of-two rescale
    if (threadIdx.x == 0) smax = 0;
    __syncthreads();
    int m = 0;
    for (int x = threadIdx.x; x < S; x += blockDim.x) m = max(m, hi_word(prod[x]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) atomicMax(&smax, m);
    __syncthreads();
    m = smax;
    double f = 1.0;
    if (m < kScaleThresholdHi && m >= (1 << 20)) {
      const int k = rescale_shift(m);
      f = pow2(k);
      Ea += k;
    }
    double* out = keep_pt + ((size_t)p.out_idx * rows + rc) * S;
    for (int x = threadIdx.x; x < S; x += blockDim.x) out[x] = prod[x] * f;
    if (threadIdx.x == 0) exp_pt[(size_t)p.out_idx * rows + rc] = Ea;
  }
}

// Root of every point: [weighted root frequencies (DRNonHomogeneousTreeLikelihood.cpp:927-962)], SR_i, site lnL, lnL.
// One CTA per point; patterns are few on this path (1 for ChromEvol).
struct PointsRootParams {
  int root_idx, S, C, ni;
  unsigned flags;  // bit0 R semantics, bit1 weighted root
  long long N;
  const double* keep;
  const int* keep_exp;
  const double* probs;
  const double* weights;
  const double* rootfreq_in;  // [npts_total][S], already offset to the chunk
  double* rootfreq_used;      // [npts][S]
  double* site_lnl;           // [npts][N]
  double* out;                // [npts][stride]: lnL at [0]
  int out_stride;
};

__global__ void points_root_kernel(PointsRootParams p) {
  extern __shared__ double sm_root[];  // S doubles: frequencies in use
  __shared__ double red[32];
  __shared__ int emin_s;
  const int S = p.S, C = p.C;
  const int pt = blockIdx.x;
  const long long rows = p.N * C;
  const double* clv = p.keep + ((size_t)pt * p.ni + p.root_idx) * rows * S;
  const int* ex = p.keep_exp + ((size_t)pt * p.ni + p.root_idx) * rows;
  double* freq = sm_root;
  if (p.flags & 2u) {
    int em = 0x7fffffff;
    for (long long i = threadIdx.x; i < rows; i += blockDim.x) em = min(em, ex[i]);
    if (threadIdx.x == 0) emin_s = 0x7fffffff;
    __syncthreads();
    atomicMin(&emin_s, em);
    __syncthreads();
    const int emin = emin_s;
    double tot = 0.0;
    for (int x = threadIdx.x; x < S; x += blockDim.x) {
      double acc = 0.0;
      for (long long i = 0; i < rows; ++i) acc = fma(clv[(size_t)i * S + x] * align_factor(ex[i] - emin), p.probs[i % C], acc);
      freq[x] = acc;
      tot += acc;
    }
    const double t = block_sum(tot, red);
    __shared__ double tot_s;
    if (threadIdx.x == 0) tot_s = t;
    __syncthreads();
    // all-zero root arrays: 1 / nbStates, like setWeightedRootFreq (DRNonHomogeneousTreeLikelihood.cpp:950-953)
    for (int x = threadIdx.x; x < S; x += blockDim.x) freq[x] = tot_s == 0.0 ? 1.0 / S : freq[x] / tot_s;
  } else {
    for (int x = threadIdx.x; x < S; x += blockDim.x) freq[x] = p.rootfreq_in[(size_t)pt * S + x];
  }
  __syncthreads();
  for (int x = threadIdx.x; x < S; x += blockDim.x) p.rootfreq_used[(size_t)pt * S + x] = freq[x];
  const bool rsem = p.flags & 1u;
  double lnl_acc = 0.0;
  for (long long pat = 0; pat < p.N; ++pat) {
    int E = ex[pat * C];
    for (int c = 1; c < C; ++c) E = min(E, ex[pat * C + c]);
    double L = 0.0;
    for (int c = 0; c < C; ++c) {
      double s = 0.0;
      for (int x = threadIdx.x; x < S; x += blockDim.x) {
        const double t = clv[((size_t)pat * C + c) * S + x] * freq[x];
        s += rsem ? (t > 0 ? t : 0.0) : t;
      }
      __syncthreads();
      s = block_sum(s, red);
      __shared__ double s_b;
      if (threadIdx.x == 0) s_b = s;
      __syncthreads();
      const double lc = s_b * align_factor(ex[pat * C + c] - E) * p.probs[c];
      L += rsem ? (lc > 0 ? lc : 0.0) : lc;
    }
    if (!rsem && L < 0) L = 0.0;
    const double lnl = log(L) - (double)E * kLn2;
    if (threadIdx.x == 0) p.site_lnl[(size_t)pt * p.N + pat] = lnl;
    lnl_acc += p.weights[pat] * lnl;
  }
  if (threadIdx.x == 0) p.out[(size_t)pt * p.out_stride] = lnl_acc;
}

}  // namespace bppgpu
