// Generic fallback for K2/K3: any S, any C, any arity; one launch per node.
// Same arithmetic as walk_kernels.cuh (same references), no specialisation:
// thread = (pattern, class, state).  All node CLVs live in the keep buffers.
#pragma once
#include "common.cuh"
#include "walk_kernels.cuh"

namespace bppgpu {

struct GenericParams {
  const Child* childs;  // children of this op
  int nchild;
  int out_idx;          // keep buffer index of the node
  int S, C, ncodes, code_bytes;
  long long N;
  const double* P;       // [nn][C][S][S]
  const double* tiptab;  // [nl][C][ncodes][S]
  const void* codes;
  double* keep;
  int* keep_exp;
};

__global__ void generic_node_kernel(GenericParams p) {
  const long long total = p.N * p.C * p.S;
  const long long rows = p.N * p.C;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(e % p.S);
    const long long rc = e / p.S;
    const int c = (int)(rc % p.C);
    const long long pat = rc / p.C;
    double acc = 1.0;
    for (int j = 0; j < p.nchild; ++j) {
      const Child ch = p.childs[j];
      double t;
      if (ch.kind == CHILD_TIP) {
        const int code = load_code(p.codes, p.code_bytes, (long long)ch.idx * p.N + pat);
        t = p.tiptab[(((size_t)ch.idx * p.C + c) * p.ncodes + code) * p.S + x];
      } else {
        const double* Pr = p.P + (((size_t)ch.pnode * p.C + c) * p.S + x) * p.S;
        const double* l = p.keep + ((size_t)ch.idx * rows + rc) * p.S;
        t = 0.0;
        for (int y = 0; y < p.S; ++y) t = fma(Pr[y], l[y], t);
      }
      acc *= t;
    }
    p.keep[((size_t)p.out_idx * rows + rc) * p.S + x] = acc;
  }
}

// thread = (pattern, class) row: accumulate child exponents, rescale the row's S entries
__global__ void generic_scale_kernel(GenericParams p) {
  const long long rows = p.N * p.C;
  const long long rc = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (rc >= rows) return;
  int Ea = 0;
  for (int j = 0; j < p.nchild; ++j) {
    const Child ch = p.childs[j];
    if (ch.kind != CHILD_TIP) Ea += p.keep_exp[(size_t)ch.idx * rows + rc];
  }
  double* v = p.keep + ((size_t)p.out_idx * rows + rc) * p.S;
  int m = 0;
  for (int i = 0; i < p.S; ++i) m = max(m, hi_word(v[i]));
  if (m < kScaleThresholdHi && m >= (1 << 20)) {
    const int k = rescale_shift(m);
    const double f = pow2(k);
    for (int i = 0; i < p.S; ++i) v[i] *= f;
    Ea += k;
  }
  p.keep_exp[(size_t)p.out_idx * rows + rc] = Ea;
}

struct RootParams {
  const double* root_clv;  // row of (pattern i, class c) = i * prow + c * crow:  [N][C][S] (prow = C, crow = 1) or class-major
  const int* root_exp;     // same rows
  long long prow, crow;
  int S, C;
  unsigned flags;
  long long N;
  const double* rootfreq;
  const double* probs;
  const double* weights;
  double* SR;
  int* rexp;
  double* site_lnl;
  double* partials;
};

__global__ void generic_root_kernel(RootParams p) {
  __shared__ double red[32];
  const long long pat = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double contrib = 0.0;
  if (pat < p.N) {
    const bool rsem = p.flags & 1u;
    const double* v = p.root_clv + (size_t)pat * p.prow * p.S;
    const int* ex = p.root_exp + (size_t)pat * p.prow;
    const size_t cs = (size_t)p.crow;
    int E = ex[0];
    for (int c = 1; c < p.C; ++c) E = min(E, ex[c * cs]);
    double L = 0.0;
    for (int c = 0; c < p.C; ++c) {
      double s = 0.0;
      for (int x = 0; x < p.S; ++x) {
        const double t = v[c * cs * p.S + x] * p.rootfreq[x];
        if (rsem) s += t > 0 ? t : 0.0;
        else s += t;
      }
      const double lc = s * align_factor(ex[c * cs] - E) * p.probs[c];
      if (rsem) L += lc > 0 ? lc : 0.0;
      else L += lc;
    }
    if (!rsem && L < 0) L = 0.0;
    const double lnl = log(L) - (double)E * kLn2;
    p.SR[pat] = L;
    p.rexp[pat] = E;
    p.site_lnl[pat] = lnl;
    contrib = p.weights[pat] * lnl;
  }
  const double bs = block_sum(contrib, red);
  if (threadIdx.x == 0) p.partials[blockIdx.x] = bs;
}

// K3b: weighted root frequencies (fork): pi_x = sum_i sum_c p_c L_root[i][c][x], normalised
// (DRNonHomogeneousTreeLikelihood::setWeightedRootFreq, DRNonHomogeneousTreeLikelihood.cpp:927-962).
// Stored CLVs carry per-pattern exponents, so patterns are aligned on the smallest one first.
// Single CTA: the fork only uses this with one character (N = 1).
struct WeightedRootParams {
  const double* root_clv;  // [N][C][S]
  const int* root_exp;     // [N][C]
  int S, C;
  long long N;
  const double* probs;
  double* out;  // [S]
};

// DRNonHomogeneousTreeLikelihood::setWeightedRootFreq (DRNonHomogeneousTreeLikelihood.cpp:927-962): freq_x proportional to
// sum_i sum_c p_c L_root[i][c][x]; 1 / nbStates when every term is zero.  Written as partial + combine so that pattern shards on
// several GPUs can exchange their (exponent, S sums) records in between: rec = [emin, s_0 .. s_{S-1}].
__global__ void weighted_root_partial_kernel(WeightedRootParams p, double* rec) {
  __shared__ int emin_s;
  __shared__ double red[32];
  int em = 0x7fffffff;
  for (long long i = threadIdx.x; i < p.N * p.C; i += blockDim.x) em = min(em, p.root_exp[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) em = min(em, __shfl_xor_sync(0xffffffffu, em, o));
  if (threadIdx.x == 0) emin_s = 0x7fffffff;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) atomicMin(&emin_s, em);
  __syncthreads();
  const int emin = emin_s;
  for (int x = 0; x < p.S; ++x) {
    double acc = 0.0;
    for (long long i = threadIdx.x; i < p.N; i += blockDim.x) {
      double a = 0.0;
      for (int c = 0; c < p.C; ++c)
        a = fma(p.root_clv[((size_t)i * p.C + c) * p.S + x] * align_factor(p.root_exp[i * p.C + c] - emin), p.probs[c], a);
      acc += a;
    }
    const double s = block_sum(acc, red);
    if (threadIdx.x == 0) rec[1 + x] = s;
  }
  if (threadIdx.x == 0) rec[0] = p.N > 0 ? (double)emin : 1e300;   // an empty shard contributes nothing
}
// recs [nrec][S + 1] (one record per pattern shard, rank order) -> out[S]
__global__ void weighted_root_combine_kernel(const double* recs, int nrec, int S, double* out) {
  __shared__ double tot_s;
  double em = 1e300;
  for (int r = 0; r < nrec; ++r) em = fmin(em, recs[(size_t)r * (S + 1)]);
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int x = 0; x < S; ++x) {
      double s = 0.0;
      for (int r = 0; r < nrec; ++r) {
        const double er = recs[(size_t)r * (S + 1)];
        if (er < 1e299) s += recs[(size_t)r * (S + 1) + 1 + x] * align_factor((int)(er - em));
      }
      out[x] = s;
      tot += s;
    }
    tot_s = tot;
  }
  __syncthreads();
  for (int x = threadIdx.x; x < S; x += blockDim.x) out[x] = tot_s == 0.0 ? 1.0 / S : out[x] / tot_s;
}

}  // namespace bppgpu
