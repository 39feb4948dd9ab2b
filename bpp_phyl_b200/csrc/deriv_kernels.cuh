// K4 + K5, generic form (any S, any C, any arity): one launch per node.
//
//   upper[n][i][c][x] = likelihood of everything outside n's subtree given state x at
//                       n's FATHER, root frequencies folded in below the root
//                       (DRHomogeneousTreeLikelihood::computeSubtreeLikelihoodPrefix,
//                        Likelihood/DRHomogeneousTreeLikelihood.cpp:543-649; the father-branch
//                        contraction uses P transposed, :919-945)
//   dL_i  = [sum_c p_c sum_x upper[n][i][c][x] sum_y dpxy[n][c][x][y] lower[n][i][c][y]] / SR_i
//                       (computeTreeDLikelihoodAtNode :287-326, D2 :373-411)
//   d1[n] = sum_i w_i dL_i ; d2[n] = sum_i w_i (d2L_i - dL_i^2)          (:340-368, :425-454)
// The NH form (DRNonHomogeneousTreeLikelihood.cpp:370-413, :497-541) multiplies the full
// conditional by dP.lower / P.lower with a `denominator == 0 -> 0` guard.
//
// This is the correctness/fallback path; the specialised fused down-walk for S = 4 / 20
// lives in downwalk_kernels.cuh.
#pragma once
#include "common.cuh"
#include "walk_kernels.cuh"

namespace bppgpu {

struct UpperParams {
  const Child* sibs;  // siblings of the node (kinds TIP / KEEP), pnode = sibling node id
  int nsib;
  int father;         // node id of the father (its branch carries P^T), -1 if father is the root
  int S, C, ncodes, code_bytes;
  long long N;
  const double* P;       // [nn][C][S][S]
  const double* tiptab;  // [nl][C][ncodes][S]
  const void* codes;
  const double* keep;    // lower CLVs of internal nodes
  const int* keep_exp;
  const double* upper_f;  // father's upper [N][C][S] (father != root)
  const int* uexp_f;
  const double* rootfreq;
  double* upper_out;
  int* uexp_out;
};

__global__ void upper_node_kernel(UpperParams p) {
  const long long total = p.N * p.C * p.S;
  const long long rows = p.N * p.C;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(e % p.S);
    const long long rc = e / p.S;
    const int c = (int)(rc % p.C);
    const long long pat = rc / p.C;
    double acc;
    if (p.father < 0) {
      acc = p.rootfreq[x];
    } else {
      // sum_y P_f[c][y][x] * upper_f[i][c][y]  (transposed contraction)
      const double* Pf = p.P + ((size_t)p.father * p.C + c) * p.S * p.S;
      const double* u = p.upper_f + (size_t)rc * p.S;
      acc = 0.0;
      for (int y = 0; y < p.S; ++y) acc = fma(Pf[(size_t)y * p.S + x], u[y], acc);
    }
    for (int j = 0; j < p.nsib; ++j) {
      const Child ch = p.sibs[j];
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
    p.upper_out[e] = acc;
  }
}

__global__ void upper_scale_kernel(UpperParams p) {
  const long long rows = p.N * p.C;
  const long long rc = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (rc >= rows) return;
  int Ea = p.father < 0 ? 0 : p.uexp_f[rc];
  for (int j = 0; j < p.nsib; ++j) {
    const Child ch = p.sibs[j];
    if (ch.kind != CHILD_TIP) Ea += p.keep_exp[(size_t)ch.idx * rows + rc];
  }
  double* v = p.upper_out + (size_t)rc * p.S;
  int m = 0;
  for (int i = 0; i < p.S; ++i) m = max(m, hi_word(v[i]));
  if (m < kScaleThresholdHi && m >= (1 << 20)) {
    const int k = rescale_shift(m);
    const double f = pow2(k);
    for (int i = 0; i < p.S; ++i) v[i] *= f;
    Ea += k;
  }
  p.uexp_out[rc] = Ea;
}

struct DerivParams {
  int node;
  int is_tip;
  int idx;  // leaf slot (tip) or keep index (internal)
  int S, C, code_bytes;
  int nh_form;
  unsigned want;  // bit1: d1, bit2: d2
  long long N;
  const double* P;    // [C][S][S] of this branch
  const double* dP;
  const double* d2P;
  const double* code_table;  // [ncodes][S]
  const void* codes;
  const double* keep;
  const int* keep_exp;
  const double* upper;  // [N][C][S]
  const int* uexp;
  const double* SR;
  const int* rexp;
  const double* probs;
  const double* weights;
  double* part1;  // [gridDim.x]
  double* part2;
};

__global__ void deriv_node_kernel(DerivParams p) {
  __shared__ double red[32];
  const long long pat = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double c1 = 0.0, c2 = 0.0;
  if (pat < p.N) {
    const int S = p.S;
    const double* D0;
    const int* le = nullptr;
    if (p.is_tip) {
      const int code = load_code(p.codes, p.code_bytes, (long long)p.idx * p.N + pat);
      D0 = p.code_table + (size_t)code * S;
    } else {
      D0 = p.keep + ((size_t)p.idx * p.N + pat) * p.C * S;
      le = p.keep_exp + ((size_t)p.idx * p.N + pat) * p.C;
    }
    const int re = p.rexp[pat];
    double a1 = 0.0, a2 = 0.0;
    for (int c = 0; c < p.C; ++c) {
      const double* D = p.is_tip ? D0 : D0 + (size_t)c * S;
      const double* U = p.upper + ((size_t)pat * p.C + c) * S;
      double s1c = 0.0, s2c = 0.0;
      for (int x = 0; x < S; ++x) {
        const double* r0 = p.P + ((size_t)c * S + x) * S;
        const double* r1 = p.dP + ((size_t)c * S + x) * S;
        const double* r2 = p.d2P ? p.d2P + ((size_t)c * S + x) * S : nullptr;
        double n1 = 0.0, n2 = 0.0, den = 0.0;
        for (int y = 0; y < S; ++y) {
          const double d = D[y];
          n1 = fma(r1[y], d, n1);
          if (r2) n2 = fma(r2[y], d, n2);
          if (p.nh_form) den = fma(r0[y], d, den);
        }
        const double u = U[x];
        if (p.nh_form) {
          const double full = u * den;
          s1c += den == 0.0 ? 0.0 : full * n1 / den;
          s2c += den == 0.0 ? 0.0 : full * n2 / den;
        } else {
          s1c = fma(u, n1, s1c);
          s2c = fma(u, n2, s2c);
        }
      }
      const int sh = re - p.uexp[(size_t)pat * p.C + c] - (le ? le[c] : 0);
      a1 = fma(scalbn(s1c, sh), p.probs[c], a1);
      a2 = fma(scalbn(s2c, sh), p.probs[c], a2);
    }
    const double sr = p.SR[pat];
    const double dL = a1 / sr;
    const double d2L = a2 / sr;
    const double w = p.weights[pat];
    c1 = w * dL;
    c2 = w * (d2L - dL * dL);
  }
  const double b1 = block_sum(c1, red);
  const double b2 = block_sum(c2, red);
  if (threadIdx.x == 0) {
    p.part1[blockIdx.x] = b1;
    p.part2[blockIdx.x] = b2;
  }
}


// Per-site relative derivatives of one branch from the resident DR arrays -- DRASDRTreeLikelihoodData::getDLikelihoodArray /
// getD2LikelihoodArray (DRHomogeneousTreeLikelihood::computeTreeDLikelihoodAtNode, DRHomogeneousTreeLikelihood.cpp:287-326,
// :373-411): dL_i / L_i and d2L_i / L_i for every pattern.  Accessor-grade (one thread per pattern); slabs follow prow / crow.
struct SiteDerivParams {
  int is_tip, S, C, code_bytes, nh_form;
  long long N, prow, crow;
  const double *P, *dP, *d2P;      // [C][S][S] of this branch (d2P may be null)
  const double* code_table;
  const void* codes;               // this leaf's codes
  const double* lower;             // slab of the node (internal)
  const int* lower_exp;
  const double* upper;             // slab of the node's upper array
  const int* upper_exp;
  const double* SR;
  const int* rexp;
  const double* probs;
  double *d1, *d2;                 // [N]
};
__global__ void site_deriv_kernel(SiteDerivParams p) {
  const long long pat = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pat >= p.N) return;
  const int S = p.S;
  const int re = p.rexp[pat];
  double a1 = 0.0, a2 = 0.0;
  for (int c = 0; c < p.C; ++c) {
    const long long row = pat * p.prow + c * p.crow;
    const double* D = p.is_tip ? p.code_table + (size_t)load_code(p.codes, p.code_bytes, pat) * S : p.lower + (size_t)row * S;
    const double* U = p.upper + (size_t)row * S;
    double s1c = 0.0, s2c = 0.0;
    for (int x = 0; x < S; ++x) {
      const double* r0 = p.P + ((size_t)c * S + x) * S;
      const double* r1 = p.dP + ((size_t)c * S + x) * S;
      const double* r2 = p.d2P ? p.d2P + ((size_t)c * S + x) * S : nullptr;
      double n1 = 0.0, n2 = 0.0, den = 0.0;
      for (int y = 0; y < S; ++y) {
        const double d = D[y];
        n1 = fma(r1[y], d, n1);
        if (r2) n2 = fma(r2[y], d, n2);
        if (p.nh_form) den = fma(r0[y], d, den);
      }
      const double u = U[x];
      if (p.nh_form) {
        const double full = u * den;
        s1c += den == 0.0 ? 0.0 : full * n1 / den;
        s2c += den == 0.0 ? 0.0 : full * n2 / den;
      } else {
        s1c = fma(u, n1, s1c);
        s2c = fma(u, n2, s2c);
      }
    }
    const int sh = re - p.upper_exp[row] - (p.is_tip ? 0 : p.lower_exp[row]);
    a1 = fma(scalbn(s1c, sh), p.probs[c], a1);
    a2 = fma(scalbn(s2c, sh), p.probs[c], a2);
  }
  const double sr = p.SR[pat];
  p.d1[pat] = a1 / sr;
  if (p.d2) p.d2[pat] = a2 / sr;
}

// ---- consumers of the DR arrays ---------------------------------------------------------------------------
// DRTreeLikelihood::computeLikelihoodAtNode (DRHomogeneousTreeLikelihood::computeLikelihoodAtNode_,
// Likelihood/DRHomogeneousTreeLikelihood.cpp:723-815):  full[i][c][x] = sub[i][c][x] * sum_y P_n[c][y][x] upper_n[i][c][y]
// (sub = the node's lower CLV, or its leaf likelihoods), times the root frequencies at the root; the arrays never leave
// the device.  Element = (pattern, class, state); rows of the device slabs follow prow / crow (class-major or not), the
// output is always in the reference's [pattern][class][state] order.
struct NodeFullParams {
  int is_leaf, is_root;
  int S, C, code_bytes;
  long long N;
  long long prow, crow;
  const double* P;           // [C][S][S] of this node's branch
  const double* lower;       // slab of the node (internal)
  const int* lower_exp;
  const void* codes;         // this leaf's codes [N] and the code table [ncodes][S]
  const double* code_table;
  const double* upper;       // slab of upper[node] (non-root)
  const int* upper_exp;
  const double* rootfreq;
  double* out;               // [N][C][S]
  int* out_exp;              // [N][C]
};

__global__ void node_full_kernel(NodeFullParams p) {
  const long long total = p.N * p.C * p.S;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(e % p.S);
    const long long rc = e / p.S;
    const int c = (int)(rc % p.C);
    const long long i = rc / p.C;
    const size_t row = (size_t)(i * p.prow + c * p.crow);
    double sub;
    int ex = 0;
    if (p.is_leaf) {
      const int code = p.code_bytes == 1 ? (int)((const unsigned char*)p.codes)[i] : (int)((const unsigned short*)p.codes)[i];
      sub = p.code_table[(size_t)code * p.S + x];
    } else {
      sub = p.lower[row * p.S + x];
      ex = p.lower_exp[row];
    }
    double v;
    if (p.is_root) {
      v = sub * p.rootfreq[x];
    } else {
      const double* U = p.upper + row * p.S;
      const double* Pc = p.P + (size_t)c * p.S * p.S;
      double acc = 0.0;
      for (int y = 0; y < p.S; ++y) acc = fma(Pc[(size_t)y * p.S + x], U[y], acc);
      v = sub * acc;
      ex += p.upper_exp[row];
    }
    p.out[e] = v;
    if (x == 0) p.out_exp[rc] = ex;
  }
}

// DRTreeLikelihoodTools::getPosteriorProbabilitiesForEachStateForEachRate (Likelihood/DRTreeLikelihoodTools.cpp:46-119):
// internal node: full / sum_{c,x} full (classes aligned on the pattern's smallest exponent first; like the reference, no class
// probabilities); leaf: leaf likelihoods[x] * p_c / sum_x leaf likelihoods.  thread = pattern.
__global__ void node_posterior_kernel(const double* full, const int* full_exp, int is_leaf, const void* codes, int code_bytes,
                                      const double* code_table, const double* probs, int S, int C, long long N, double* post) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  double* o = post + (size_t)i * C * S;
  if (is_leaf) {
    const int code = code_bytes == 1 ? (int)((const unsigned char*)codes)[i] : (int)((const unsigned short*)codes)[i];
    const double* t = code_table + (size_t)code * S;
    double sum = 0.0;
    for (int x = 0; x < S; ++x) sum += t[x];
    for (int c = 0; c < C; ++c)
      for (int x = 0; x < S; ++x) o[c * S + x] = t[x] * probs[c] / sum;
    return;
  }
  const double* f = full + (size_t)i * C * S;
  const int* ex = full_exp + (size_t)i * C;
  int E = ex[0];
  for (int c = 1; c < C; ++c) E = min(E, ex[c]);
  double sum = 0.0;
  for (int c = 0; c < C; ++c) {
    const double a = align_factor(ex[c] - E);
    for (int x = 0; x < S; ++x) sum += f[c * S + x] * a;
  }
  for (int c = 0; c < C; ++c) {
    const double a = align_factor(ex[c] - E);
    for (int x = 0; x < S; ++x) o[c * S + x] = f[c * S + x] * a / sum;
  }
}


// MarginalNonRevAncestralStateReconstruction (fork; Likelihood/MarginalNonRevAncestralStateReconstruction.cpp:10-136): per
// distinct site the posterior of every state x at a node and the joint posterior of (x, father state y).  The reference sums,
// over the S root states r, pi_r times a prefix pass conditional on r (DRNonHomogeneousTreeLikelihood.cpp:1026-1162); that sum
// IS the ordinary prefix array (root frequencies folded in at the root's sons), so the resident upper slab gives it in one pass:
//   joint[i][x][y] = sum_c p_c upper[i][c][y] P[c][y][x] sub[i][c][x] / L_i     post[i][x] = sum_y joint[i][x][y]
//   root:            post[i][x] = sum_c p_c pi_x lower[i][c][x] / L_i           (:104-115)
// thread = (pattern, node state); accessor-grade CUDA-core kernel.  joint may be null.
struct MarginalParams {
  int is_leaf, is_root;
  int S, C, code_bytes;
  long long N, prow, crow;
  const double* P;           // [C][S][S] of this node's branch
  const double* lower;       // slab of the node (internal) + exponents
  const int* lower_exp;
  const void* codes;         // leaf: codes [N]
  const double* code_table;
  const double* upper;       // slab of upper[node] (non-root) + exponents
  const int* upper_exp;
  const double *rootfreq, *probs, *SR;
  const int* rexp;
  double* post;              // [N][S]
  double* joint;             // [N][S][S]  (x = node state, y = father state) or null
};

__global__ void marginal_posterior_kernel(MarginalParams p) {
  const long long total = p.N * p.S;
  const int S = p.S;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(e % S);
    const long long i = e / S;
    const double sr = p.SR[i];
    const int re = p.rexp[i];
    double leafv = 0.0;
    if (p.is_leaf) {
      const int code = p.code_bytes == 1 ? (int)((const unsigned char*)p.codes)[i] : (int)((const unsigned short*)p.codes)[i];
      leafv = p.code_table[(size_t)code * S + x];
    }
    if (p.is_root) {
      double acc = 0.0;
      for (int c = 0; c < p.C; ++c) {
        const size_t row = (size_t)(i * p.prow + c * p.crow);
        const double sub = p.is_leaf ? leafv : p.lower[row * S + x];
        const int sh = re - (p.is_leaf ? 0 : p.lower_exp[row]);
        acc += scalbn(sub * p.rootfreq[x], sh) * p.probs[c];
      }
      p.post[e] = acc / sr;
      continue;
    }
    double tot = 0.0;
    for (int y = 0; y < S; ++y) {
      double acc = 0.0;
      for (int c = 0; c < p.C; ++c) {
        const size_t row = (size_t)(i * p.prow + c * p.crow);
        const double sub = p.is_leaf ? leafv : p.lower[row * S + x];
        const int sh = re - (p.is_leaf ? 0 : p.lower_exp[row]) - p.upper_exp[row];
        const double t = p.upper[row * S + y] * p.P[((size_t)c * S + y) * S + x] * sub;
        acc += scalbn(t, sh) * p.probs[c];
      }
      acc /= sr;
      if (p.joint) p.joint[((size_t)i * S + x) * S + y] = acc;
      tot += acc;
    }
    p.post[e] = tot;
  }
}


// ---- joint ML ancestral reconstruction (fork) ------------------------------------------------------------------------------
// MLAncestralStateReconstruction (Likelihood/MLAncestralStateReconstruction.cpp:6-188; leaf arrays
// DRASRTreeLikelihoodData.cpp:265-300): Pupko's max-product recursion.  Per (pattern i, class c), indexed by the FATHER's state x:
//   leaf with observed state s (the first state whose init value is 1):  L[x] = P[c][x][s],  anc[x] = s
//       (no such state -- composite / probabilistic tips: L[x] = 1, anc[x] = 0, as the reference leaves them)
//   internal node:   L[x] = max_y P[c][x][y] prod_sons L_son[y],  anc[x] = the first y reaching the (positive) maximum
//   root:            L[x] = pi_x prod_sons L_son[x]
// anc is the table of the LAST class (the reference overwrites it class after class); the root's state is the first maximum of
// class 0 (:151-160).  The reference has no rescaling and underflows to 0 on large trees; here every (i, c) row carries a
// power-of-two exponent (max-product is invariant under a per-row scale), so the tables equal the reference's wherever it is
// finite.  One warp per (pattern, class) row; accessor-grade CUDA-core kernels.
struct MLNodeParams {
  int kind;                  // 0 leaf, 1 internal, 2 root
  int S, C, code_bytes, nson;
  long long N;
  const double* P;           // [C][S][S] of this node's branch (leaf, internal)
  const void* codes;         // leaf
  const double* code_table;
  const double* son_L[4];    // [N][C][S] of every son (internal, root)
  const int* son_E[4];
  const double* rootfreq;
  double* L;                 // [N][C][S]
  int* E;                    // [N][C]
  unsigned short* anc;       // [N][S]  (leaf, internal)
};

__global__ void ml_node_kernel(MLNodeParams p) {
  extern __shared__ double ml_sm[];        // [warps per block][S]
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + wib;
  if (row >= p.N * p.C) return;
  const int S = p.S;
  const int c = (int)(row % p.C);
  const long long i = row / p.C;
  double* prod = ml_sm + (size_t)wib * S;
  double* Lrow = p.L + (size_t)row * S;
  const bool last_class = c == p.C - 1;
  if (p.kind == 0) {
    const int code = p.code_bytes == 1 ? (int)((const unsigned char*)p.codes)[i] : (int)((const unsigned short*)p.codes)[i];
    const double* t = p.code_table + (size_t)code * S;
    int s = -1;
    for (int y = 0; y < S; ++y)
      if (t[y] == 1.0) { s = y; break; }
    const double* Pc = p.P + (size_t)c * S * S;
    for (int x = lane; x < S; x += 32) {
      Lrow[x] = s >= 0 ? Pc[(size_t)x * S + s] : 1.0;
      if (last_class) p.anc[(size_t)i * S + x] = (unsigned short)(s >= 0 ? s : 0);
    }
    if (lane == 0) p.E[row] = 0;
    return;
  }
  int e = 0;
  for (int j = 0; j < p.nson; ++j) e += p.son_E[j][row];
  for (int y = lane; y < S; y += 32) {
    double v = p.son_L[0][(size_t)row * S + y];
    for (int j = 1; j < p.nson; ++j) v *= p.son_L[j][(size_t)row * S + y];
    prod[y] = v;
  }
  __syncwarp();
  double mx = 0.0;
  if (p.kind == 2) {
    for (int x = lane; x < S; x += 32) {
      const double v = p.rootfreq[x] * prod[x];
      Lrow[x] = v;
      mx = fmax(mx, v);
    }
  } else {
    const double* Pc = p.P + (size_t)c * S * S;
    for (int x = lane; x < S; x += 32) {
      double best = 0.0;
      int arg = 0;
      const double* Px = Pc + (size_t)x * S;
      for (int y = 0; y < S; ++y) {
        const double v = prod[y] * Px[y];
        if (v > best) { best = v; arg = y; }
      }
      Lrow[x] = best;
      if (last_class) p.anc[(size_t)i * S + x] = (unsigned short)arg;
      mx = fmax(mx, best);
    }
  }
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const int mh = hi_word(mx);
  if (mx > 0.0 && mh < kScaleThresholdHi && mh >= (1 << 20)) {   // same rule as the pruning kernels
    const int k = rescale_shift(mh);
    const double f = pow2(k);
    for (int x = lane; x < S; x += 32) Lrow[x] *= f;            // each lane re-reads what it wrote
    e += k;
  }
  if (lane == 0) p.E[row] = e;
}

// state of the root: first maximum of class 0 (:151-160); best_lnl = log of that joint likelihood.  thread = pattern
__global__ void ml_root_state_kernel(const double* L, const int* E, int S, int C, long long N, int* state, double* best_lnl) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const double* l = L + (size_t)i * C * S;
  double best = 0.0;
  int arg = 0;
  for (int x = 0; x < S; ++x)
    if (l[x] > best) { best = l[x]; arg = x; }
  state[i] = arg;
  if (best_lnl) best_lnl[i] = log(best) - E[(size_t)i * C] * kLn2;
}

// trace back: state[node][i] = anc[node][i][state[father][i]] (:162-167; a leaf's table is constant in the father's state)
__global__ void ml_traceback_kernel(const unsigned short* anc, const int* father_state, int S, long long N, int* state) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) state[i] = (int)anc[(size_t)i * S + father_state[i]];
}


// ---- BrLenRoot / RootPosition (re-parametrised root branches of a rooted tree) ---------------------------------------------
// DRNonHomogeneousTreeLikelihood::getFirstOrderDerivative / getSecondOrderDerivative for the two parameters that replace the
// root branches l1 = len * pos, l2 = len * (1 - pos) (Likelihood/DRNonHomogeneousTreeLikelihood.cpp:445-478, :576-867;
// AbstractNonHomogeneousTreeLikelihood.cpp:319-330, :386-389).  The second derivatives need the cross term
// (dP_1 L_1)(dP_2 L_2), which the per-branch derivative pass does not produce, so everything is rebuilt at the root from the
// two sons' lower arrays:  per root state x
//   d_len   = pos dl1 l2 + (1 - pos) dl2 l1                    d_pos   = len (dl1 l2 - dl2 l1)
//   d2_len  = pos^2 d2l1 l2 + (1 - pos)^2 d2l2 l1 + 2 pos (1 - pos) dl1 dl2        (:662-663)
//   d2_pos  = len^2 (d2l1 l2 + d2l2 l1 - 2 dl1 dl2)
// times the other root sons' P.L, pi_x and p_c, over SR_i; the four weighted sums  sum_i w_i D_i  and
// sum_i w_i (D2_i - D_i^2)  come back as per-block partials (derivatives of +lnL; the reference returns those of -lnL).
// thread = pattern; accessor-grade CUDA-core kernel (6 S^2 multiply-adds per class and pattern).
struct RootReparamSon {
  int is_leaf;
  const double* lower;    // internal: lower slab + exponents (rows follow prow / crow)
  const int* lower_exp;
  const void* codes;      // leaf: codes [N]
  const double *P, *dP, *d2P;  // [C][S][S] of the son's branch (dP / d2P unused for the "other" sons)
};
struct RootReparamParams {
  RootReparamSon sons[4];   // [0] = root1, [1] = root2, then up to two other root sons
  int nson;
  int S, C, code_bytes;
  long long N, prow, crow;
  double pos, len;
  const double* code_table;
  const double *rootfreq, *probs, *weights, *SR;
  const int* rexp;
  double* part;   // [4][gridDim.x]
};

__global__ void root_reparam_kernel(RootReparamParams p) {
  __shared__ double red[32];
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int S = p.S;
  double D[4] = {0.0, 0.0, 0.0, 0.0};  // d_len, d_pos, d2_len, d2_pos of this pattern (already / SR)
  if (i < p.N) {
    for (int c = 0; c < p.C; ++c) {
      const size_t row = (size_t)(i * p.prow + c * p.crow);
      const double* L[4];
      int e = 0;
      for (int j = 0; j < p.nson; ++j) {
        if (p.sons[j].is_leaf) {
          const int code = p.code_bytes == 1 ? (int)((const unsigned char*)p.sons[j].codes)[i] : (int)((const unsigned short*)p.sons[j].codes)[i];
          L[j] = p.code_table + (size_t)code * S;
        } else {
          L[j] = p.sons[j].lower + row * S;
          e += p.sons[j].lower_exp[row];
        }
      }
      const size_t mo = (size_t)c * S * S;
      double a[4] = {0.0, 0.0, 0.0, 0.0};
      for (int x = 0; x < S; ++x) {
        double l1 = 0, l2 = 0, dl1 = 0, dl2 = 0, d2l1 = 0, d2l2 = 0;
        const size_t ro = mo + (size_t)x * S;
        for (int y = 0; y < S; ++y) {
          const double v1 = L[0][y], v2 = L[1][y];
          l1 = fma(p.sons[0].P[ro + y], v1, l1);
          dl1 = fma(p.sons[0].dP[ro + y], v1, dl1);
          d2l1 = fma(p.sons[0].d2P[ro + y], v1, d2l1);
          l2 = fma(p.sons[1].P[ro + y], v2, l2);
          dl2 = fma(p.sons[1].dP[ro + y], v2, dl2);
          d2l2 = fma(p.sons[1].d2P[ro + y], v2, d2l2);
        }
        double o = p.rootfreq[x];
        for (int j = 2; j < p.nson; ++j) {
          double t = 0;
          for (int y = 0; y < S; ++y) t = fma(p.sons[j].P[ro + y], L[j][y], t);
          o *= t;
        }
        a[0] = fma(o, p.pos * dl1 * l2 + (1.0 - p.pos) * dl2 * l1, a[0]);
        a[1] = fma(o, p.len * (dl1 * l2 - dl2 * l1), a[1]);
        a[2] = fma(o, p.pos * p.pos * d2l1 * l2 + (1.0 - p.pos) * (1.0 - p.pos) * d2l2 * l1 + 2.0 * p.pos * (1.0 - p.pos) * dl1 * dl2, a[2]);
        a[3] = fma(o, p.len * p.len * (d2l1 * l2 + d2l2 * l1 - 2.0 * dl1 * dl2), a[3]);
      }
      const int sh = p.rexp[i] - e;
      const double f = p.probs[c] / p.SR[i];
      for (int k = 0; k < 4; ++k) D[k] += scalbn(a[k], sh) * f;
    }
  }
  const double w = i < p.N ? p.weights[i] : 0.0;
  const double c0 = w * D[0], c1 = w * D[1], c2 = w * (D[2] - D[0] * D[0]), c3 = w * (D[3] - D[1] * D[1]);
  const double b0 = block_sum(c0, red), b1 = block_sum(c1, red), b2 = block_sum(c2, red), b3 = block_sum(c3, red);
  if (threadIdx.x == 0) {
    p.part[0 * gridDim.x + blockIdx.x] = b0;
    p.part[1 * gridDim.x + blockIdx.x] = b1;
    p.part[2 * gridDim.x + blockIdx.x] = b2;
    p.part[3 * gridDim.x + blockIdx.x] = b3;
  }
}

// one block per output: out[k] = sum of part[k][0 .. n)
__global__ void root_reparam_finalize_kernel(const double* part, int n, double* out) {
  __shared__ double red[32];
  double a = 0.0;
  for (int j = threadIdx.x; j < n; j += blockDim.x) a += part[(size_t)blockIdx.x * n + j];
  const double s = block_sum(a, red);
  if (threadIdx.x == 0) out[blockIdx.x] = s;
}

}  // namespace bppgpu
