// K2 + K3 for small state counts (S = 4, 20): the "tile walk".
//
// The pruning recursion (RHomogeneousTreeLikelihood::computeSubtreeLikelihood,
// Likelihood/RHomogeneousTreeLikelihood.cpp:802-863; DR twin
// DRHomogeneousTreeLikelihood.cpp:483-539 + :819-864) is independent per
// (pattern, rate class).  So instead of one launch per tree level, ONE launch
// walks the whole post-order "program" and every thread owns one
// (pattern, class) row for the entire tree: it only ever reads CLVs it wrote
// itself, which removes every inter-thread dependency, keeps the live CLVs of
// the recursion on chip (registers + a small per-thread stack), and fuses the
// root reduction (:162-216 / DRHomogeneousTreeLikelihood.cpp:653-719) into the
// same kernel.  With BPPGPU_FLAG_KEEP_CLVS every node's CLV is additionally
// streamed out (write-only, coalesced 32 B per thread) for the derivative pass
// and for getLikelihoodData()-style consumers.
//
// Tips are 1-byte (or 2-byte) codes resolved through the per-branch tip table
// (pt_kernels.cuh), never expanded to S doubles.
#pragma once
#include "common.cuh"

namespace bppgpu {

enum ChildKind { CHILD_TIP = 0, CHILD_SLOT = 1, CHILD_REG = 2, CHILD_KEEP = 3 };

struct Child {
  int kind;   // ChildKind
  int idx;    // leaf slot (TIP), stack slot (SLOT), keep buffer index (KEEP)
  int pnode;  // node id whose branch carries this child (index into P tables)
  int pad;
};

struct Op {
  int node;         // node id
  int nchild;
  int child_begin;  // into the Child array
  int dst_slot;     // stack slot that receives the result, or -1 (stays in registers)
  int keep_idx;     // keep buffer index, or -1
  int is_root;
  int pad0, pad1;
};

struct WalkParams {
  const Op* ops;
  const Child* childs;
  int n_ops;
  int C;
  int ncodes;
  int code_bytes;
  int nslots;
  unsigned flags;           // bit0: R semantics at the root
  long long N;              // patterns
  const double* P;          // [nn][C][S][S]      (this point)
  const double* tiptab;     // [nl][C][ncodes][S] (this point)
  const void* codes;        // [nl][N]
  double* keep;             // [n_internal][N][C][S] or nullptr
  int* keep_exp;            // [n_internal][N]
  double* gstack;           // [nslots][N][C][S] (global-stack variant) or nullptr
  int* gstack_exp;          // [nslots][N]
  const double* rootfreq;   // [S]
  const double* probs;      // [C]
  const double* weights;    // [N]
  double* SR;               // [N] scaled site likelihood
  int* rexp;                // [N] its exponent
  double* site_lnl;         // [N]
  double* partials;         // [gridDim.x] weighted lnL partial sums
};

constexpr int kWalkThreads = 256;

__device__ __forceinline__ int load_code(const void* codes, int code_bytes, long long off) {
  return code_bytes == 1 ? (int)((const unsigned char*)codes)[off] : (int)((const unsigned short*)codes)[off];
}

// ----------------------------------------------------------------------------
// S = 4 (DNA): thread = (pattern, class); 4 doubles per row; stack in shared memory.
// ----------------------------------------------------------------------------
template <int C_LOG2>
__global__ void __launch_bounds__(kWalkThreads) walk4_kernel(WalkParams prm) {
  constexpr int C = 1 << C_LOG2;
  extern __shared__ __align__(32) unsigned char smem_raw[];
  double4* st = reinterpret_cast<double4*>(smem_raw);                       // [nslots][256]
  int* ste = reinterpret_cast<int*>(smem_raw + (size_t)prm.nslots * kWalkThreads * 32);  // [nslots][256]
  __shared__ double red[32];

  const int tid = threadIdx.x;
  const long long rows = prm.N << C_LOG2;
  const long long r0 = (long long)blockIdx.x * kWalkThreads + tid;
  const bool valid = r0 < rows;
  const long long r = valid ? r0 : rows - 1;
  const long long pat = r >> C_LOG2;
  const int c = (int)(r & (C - 1));

  double v0 = 1.0, v1 = 1.0, v2 = 1.0, v3 = 1.0;
  int E = 0;

  for (int o = 0; o < prm.n_ops; ++o) {
    const Op op = prm.ops[o];
    double a0 = 1.0, a1 = 1.0, a2 = 1.0, a3 = 1.0;
    int Ea = 0;
    for (int j = 0; j < op.nchild; ++j) {
      const Child ch = prm.childs[op.child_begin + j];
      double t0, t1, t2, t3;
      if (ch.kind == CHILD_TIP) {
        const int code = load_code(prm.codes, prm.code_bytes, (long long)ch.idx * prm.N + pat);
        const double* tt = prm.tiptab + (((size_t)ch.idx * C + c) * prm.ncodes + code) * 4;
        ld256nc(tt, t0, t1, t2, t3);
      } else {
        double l0, l1, l2, l3;
        int e;
        if (ch.kind == CHILD_REG) {
          l0 = v0; l1 = v1; l2 = v2; l3 = v3; e = E;
        } else if (ch.kind == CHILD_SLOT) {
          const double4 s = st[ch.idx * kWalkThreads + tid];
          l0 = s.x; l1 = s.y; l2 = s.z; l3 = s.w;
          e = ste[ch.idx * kWalkThreads + tid];
        } else {
          ld256(prm.keep + ((size_t)ch.idx * rows + r) * 4, l0, l1, l2, l3);
          e = prm.keep_exp[(size_t)ch.idx * prm.N + pat];
        }
        const double* Pm = prm.P + ((size_t)ch.pnode * C + c) * 16;
        double p0, p1, p2, p3;
        ld256nc(Pm, p0, p1, p2, p3);
        t0 = fma(p3, l3, fma(p2, l2, fma(p1, l1, p0 * l0)));
        ld256nc(Pm + 4, p0, p1, p2, p3);
        t1 = fma(p3, l3, fma(p2, l2, fma(p1, l1, p0 * l0)));
        ld256nc(Pm + 8, p0, p1, p2, p3);
        t2 = fma(p3, l3, fma(p2, l2, fma(p1, l1, p0 * l0)));
        ld256nc(Pm + 12, p0, p1, p2, p3);
        t3 = fma(p3, l3, fma(p2, l2, fma(p1, l1, p0 * l0)));
        Ea += e;
      }
      a0 *= t0; a1 *= t1; a2 *= t2; a3 *= t3;
    }
    // per-pattern power-of-two rescale (max over states and classes)
    int m = max(max(hi_word(a0), hi_word(a1)), max(hi_word(a2), hi_word(a3)));
#pragma unroll
    for (int off = 1; off < C; off <<= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, off));
    if (m < kScaleThresholdHi && m >= (1 << 20)) {
      const int k = rescale_shift(m);
      const double f = pow2(k);
      a0 *= f; a1 *= f; a2 *= f; a3 *= f;
      Ea += k;
    }
    v0 = a0; v1 = a1; v2 = a2; v3 = a3;
    E = Ea;
    if (op.dst_slot >= 0) {
      st[op.dst_slot * kWalkThreads + tid] = make_double4(v0, v1, v2, v3);
      ste[op.dst_slot * kWalkThreads + tid] = E;
    }
    if (op.keep_idx >= 0 && valid) {
      st256(prm.keep + ((size_t)op.keep_idx * rows + r) * 4, v0, v1, v2, v3);
      if (c == 0) prm.keep_exp[(size_t)op.keep_idx * prm.N + pat] = E;
    }
  }

  // ---- root reduction: L_i = sum_c p_c sum_x pi_x CLV_root[i][c][x] ----------
  const bool rsem = prm.flags & 1u;
  const double t0 = v0 * prm.rootfreq[0], t1 = v1 * prm.rootfreq[1];
  const double t2 = v2 * prm.rootfreq[2], t3 = v3 * prm.rootfreq[3];
  double s;
  if (rsem) s = (t0 > 0 ? t0 : 0.0) + (t1 > 0 ? t1 : 0.0) + (t2 > 0 ? t2 : 0.0) + (t3 > 0 ? t3 : 0.0);
  else s = ((t0 + t1) + t2) + t3;
  double L = s * prm.probs[c];
  if (rsem && !(L > 0)) L = 0.0;
#pragma unroll
  for (int off = 1; off < C; off <<= 1) L += __shfl_xor_sync(0xffffffffu, L, off);
  if (!rsem && L < 0) L = 0.0;
  double contrib = 0.0;
  if (valid && c == 0) {
    const double lnl = log(L) - (double)E * kLn2;
    prm.SR[pat] = L;
    prm.rexp[pat] = E;
    prm.site_lnl[pat] = lnl;
    contrib = prm.weights[pat] * lnl;
  }
  const double bs = block_sum(contrib, red);
  if (tid == 0) prm.partials[blockIdx.x] = bs;
}

// ----------------------------------------------------------------------------
// General small S (multiple of 4, S <= 32; used for S = 20): thread = (pattern, class),
// P^T staged in shared memory per op, stack in global memory (L2-resident slots).
// ----------------------------------------------------------------------------
template <int S>
__device__ __forceinline__ void load_row(const double* g, double (&l)[S]) {
#pragma unroll
  for (int i = 0; i < S; i += 4) ld256(g + i, l[i], l[i + 1], l[i + 2], l[i + 3]);
}
template <int S>
__device__ __forceinline__ void load_row_nc(const double* g, double (&l)[S]) {
#pragma unroll
  for (int i = 0; i < S; i += 4) ld256nc(g + i, l[i], l[i + 1], l[i + 2], l[i + 3]);
}
template <int S>
__device__ __forceinline__ void store_row(double* g, const double (&l)[S]) {
#pragma unroll
  for (int i = 0; i < S; i += 4) st256(g + i, l[i], l[i + 1], l[i + 2], l[i + 3]);
}

// shared layout of one staged matrix: PT[c][y][x] with the class stride padded so
// that the C different matrices read by the lanes of one warp fall in distinct banks
template <int S, int C>
struct StageLayout {
  static constexpr int kMat = S * S;
  static constexpr int kCStride = kMat + 2;             // +16 B
  static constexpr int kChildStride = C * kCStride;     // doubles
};

constexpr int kMaxStagedChildren = 3;

template <int S, int C_LOG2>
__global__ void __launch_bounds__(kWalkThreads) walkS_kernel(WalkParams prm) {
  constexpr int C = 1 << C_LOG2;
  using L = StageLayout<S, C>;
  extern __shared__ __align__(32) unsigned char smem_raw[];
  double* PT = reinterpret_cast<double*>(smem_raw);  // [kMaxStagedChildren][C][S*S+2]
  __shared__ double red[32];

  const int tid = threadIdx.x;
  const long long rows = prm.N << C_LOG2;
  const long long r0 = (long long)blockIdx.x * kWalkThreads + tid;
  const bool valid = r0 < rows;
  const long long r = valid ? r0 : rows - 1;
  const long long pat = r >> C_LOG2;
  const int c = (int)(r & (C - 1));

  double v[S];
#pragma unroll
  for (int x = 0; x < S; ++x) v[x] = 1.0;
  int E = 0;

  for (int o = 0; o < prm.n_ops; ++o) {
    const Op op = prm.ops[o];
    // ---- stage P^T of every non-tip child ------------------------------------
    __syncthreads();
    for (int j = 0; j < op.nchild; ++j) {
      const Child ch = prm.childs[op.child_begin + j];
      if (ch.kind == CHILD_TIP) continue;
      const double* Pg = prm.P + (size_t)ch.pnode * C * S * S;
      double* dst = PT + (size_t)(j % kMaxStagedChildren) * L::kChildStride;
      for (int e = tid; e < C * S * S; e += kWalkThreads) {
        const int cc = e / (S * S);
        const int rem = e - cc * S * S;
        const int x = rem / S, y = rem - x * S;
        dst[cc * L::kCStride + y * S + x] = Pg[e];
      }
    }
    __syncthreads();

    double a[S];
    int Ea = 0;
    for (int j = 0; j < op.nchild; ++j) {
      const Child ch = prm.childs[op.child_begin + j];
      double t[S];
      if (ch.kind == CHILD_TIP) {
        const int code = load_code(prm.codes, prm.code_bytes, (long long)ch.idx * prm.N + pat);
        load_row_nc<S>(prm.tiptab + (((size_t)ch.idx * C + c) * prm.ncodes + code) * S, t);
      } else {
        double l[S];
        int e;
        if (ch.kind == CHILD_REG) {
#pragma unroll
          for (int x = 0; x < S; ++x) l[x] = v[x];
          e = E;
        } else if (ch.kind == CHILD_SLOT) {
          load_row<S>(prm.gstack + ((size_t)ch.idx * rows + r) * S, l);
          e = prm.gstack_exp[(size_t)ch.idx * prm.N + pat];
        } else {
          load_row<S>(prm.keep + ((size_t)ch.idx * rows + r) * S, l);
          e = prm.keep_exp[(size_t)ch.idx * prm.N + pat];
        }
        const double* pt = PT + (size_t)(j % kMaxStagedChildren) * L::kChildStride + c * L::kCStride;
#pragma unroll
        for (int x = 0; x < S; ++x) t[x] = 0.0;
#pragma unroll
        for (int y = 0; y < S; ++y) {
          const double ly = l[y];
#pragma unroll
          for (int x = 0; x < S; x += 2) {
            const double2 pp = *reinterpret_cast<const double2*>(pt + y * S + x);
            t[x] = fma(pp.x, ly, t[x]);
            t[x + 1] = fma(pp.y, ly, t[x + 1]);
          }
        }
        Ea += e;
      }
      if (j == 0) {
#pragma unroll
        for (int x = 0; x < S; ++x) a[x] = t[x];
      } else {
#pragma unroll
        for (int x = 0; x < S; ++x) a[x] *= t[x];
      }
    }
    int m = 0;
#pragma unroll
    for (int x = 0; x < S; ++x) m = max(m, hi_word(a[x]));
#pragma unroll
    for (int off = 1; off < C; off <<= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, off));
    if (m < kScaleThresholdHi && m >= (1 << 20)) {
      const int k = rescale_shift(m);
      const double f = pow2(k);
#pragma unroll
      for (int x = 0; x < S; ++x) a[x] *= f;
      Ea += k;
    }
#pragma unroll
    for (int x = 0; x < S; ++x) v[x] = a[x];
    E = Ea;
    if (op.dst_slot >= 0 && valid) {
      store_row<S>(prm.gstack + ((size_t)op.dst_slot * rows + r) * S, v);
      if (c == 0) prm.gstack_exp[(size_t)op.dst_slot * prm.N + pat] = E;
    }
    if (op.keep_idx >= 0 && valid) {
      store_row<S>(prm.keep + ((size_t)op.keep_idx * rows + r) * S, v);
      if (c == 0) prm.keep_exp[(size_t)op.keep_idx * prm.N + pat] = E;
    }
  }

  const bool rsem = prm.flags & 1u;
  double s = 0.0;
#pragma unroll
  for (int x = 0; x < S; ++x) {
    const double tx = v[x] * prm.rootfreq[x];
    if (rsem) s += tx > 0 ? tx : 0.0;
    else s += tx;
  }
  double Lk = s * prm.probs[c];
  if (rsem && !(Lk > 0)) Lk = 0.0;
#pragma unroll
  for (int off = 1; off < C; off <<= 1) Lk += __shfl_xor_sync(0xffffffffu, Lk, off);
  if (!rsem && Lk < 0) Lk = 0.0;
  double contrib = 0.0;
  if (valid && c == 0) {
    const double lnl = log(Lk) - (double)E * kLn2;
    prm.SR[pat] = Lk;
    prm.rexp[pat] = E;
    prm.site_lnl[pat] = lnl;
    contrib = prm.weights[pat] * lnl;
  }
  const double bs = block_sum(contrib, red);
  if (tid == 0) prm.partials[blockIdx.x] = bs;
}

// final deterministic reduction of the per-block partial sums (single block)
__global__ void finalize_sum_kernel(const double* partials, int n, double* out) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partials[i];
  const double s = block_sum(acc, red);
  if (threadIdx.x == 0) *out = s;
}

}  // namespace bppgpu
