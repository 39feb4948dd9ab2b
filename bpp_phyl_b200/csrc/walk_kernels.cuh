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

enum ChildKind { CHILD_TIP = 0, CHILD_SLOT = 1, CHILD_REG = 2, CHILD_KEEP = 3, CHILD_RSLOT = 4 /* register slot (walk4c) */ };

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
  int* keep_exp;            // [n_internal][N][C]
  double* gstack;           // [nslots][N][C][S] (global-stack variant) or nullptr
  int* gstack_exp;          // [nslots][N][C]
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
// S = 4 (DNA): thread = (pattern, class) row, 4 doubles per row, stack in shared memory.
//
// Everything the walk touches is laid out in WALK ORDER so that no address depends on a
// dependent load:
//   * desc[o]   one packed 64-bit descriptor per op (byte 0 nchild, byte 1 dst_slot+1,
//               bytes 2..7 child tokens kind<<6|slot), prefetched one op ahead;
//   * stream    the P / tip tables of every child, concatenated in consumption order
//               (pack_stream_kernel):  internal child -> [row x][class][4]  (a warp reads
//               ONE 128-byte line per row),  tip child -> [code][class][4];
//   * codesT    tip codes transposed to [pattern][tip in consumption order]; each thread
//               streams its own row 8 bytes at a time, one 8-byte word ahead of use, so the
//               only DRAM-latency load of the kernel is always in flight early.
// Rescaling is per row (pattern, class): max over 4 high words on the integer pipe, no
// shuffles inside the walk; classes are re-aligned once at the root.
// ----------------------------------------------------------------------------
struct Walk4Params {
  const unsigned long long* desc;  // [n_ops]
  int n_ops;
  int nslots;
  int ncodes;
  int tstride;                     // bytes per codesT row (multiple of 8, >= n_tips + 16)
  unsigned flags;                  // bit0: R semantics at the root
  long long N;
  const double* stream;            // this point's packed tables
  const unsigned char* codesT;     // [N][tstride]
  double* keep;                    // [n_ops][N*C][4] or nullptr
  int* keep_exp;                   // [n_ops][N*C]
  const double* rootfreq;          // [4]
  const double* probs;             // [C]
  const double* weights;           // [N]
  double* SR;                      // [N]
  int* rexp;                       // [N]
  double* site_lnl;                // [N]
  double* partials;                // [gridDim.x]
};

__device__ __forceinline__ unsigned long long ldg_u64_nc(const unsigned char* p) {
  unsigned long long v;
  asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}

// PT = patterns per thread.  The L1 data pipe delivers 128 B/clk/SM to the register file, and a thread needs
// the whole 128-byte P block of its class for every internal child: with one pattern per thread the kernel is
// bound by that pipe (ncu r1a: l1tex data-pipe wavefronts 77 %, FP64 pipe 18 %).  Holding PT patterns of the same
// class in one thread re-uses every loaded P row PT times.
//
// Instruction overhead is the other limiter (ncu r1b: 128 issued instructions per pattern-op for 20 FP64 ones), so
// the binary ops -- 99.9 % of a bifurcating tree -- are dispatched by a warp-uniform switch on a host-computed
// SHAPE code to straight-line handlers specialised at compile time on the two child kinds; anything else (the
// 3-son root, multifurcations) takes the generic child loop.
//
// Measured on the 1024-taxon x 1M-pattern config (ms per evaluation): PT=1 30.0, PT=2 18.7, PT=4 19.3.  Tried and
// dropped because they measured slower (DESIGN.md has the numbers): a software-pipelined event loop, register
// prefetch of the next op's tables (costs the occupancy it tries to replace) and prefetch.global.L1 of the stream.
constexpr int kWalk4Threads = 128;

// shape = 1 + 3*k(child0) + k(child1) with k: TIP 0, SLOT 1, REG 2;  0 = generic loop
__host__ __device__ constexpr int walk4_kind_index(int kind) { return kind == CHILD_TIP ? 0 : (kind == CHILD_SLOT ? 1 : 2); }

template <int C_LOG2, int PT>
struct Walk4State {
  static constexpr int C = 1 << C_LOG2;
  static constexpr int NTH = kWalk4Threads;
  double v[PT][4];
  int E[PT];
  unsigned long long q[PT], qn[PT];
  const unsigned char* crow[PT];
  int tipk;
  const double* sp;
  int tip_block;
  int c, tid;
  double4* st;
  int* ste;

  // t = tip vector of the next tip in consumption order
  __device__ __forceinline__ void term_tip(double (&t)[PT][4]) {
#pragma unroll
    for (int j = 0; j < PT; ++j) {
      const int code = (int)(q[j] & 0xffu);
      q[j] >>= 8;
      ld256nc(sp + (((code << C_LOG2) + c) << 2), t[j][0], t[j][1], t[j][2], t[j][3]);
    }
    if (((++tipk) & 7) == 0) {
#pragma unroll
      for (int j = 0; j < PT; ++j) {
        q[j] = qn[j];
        qn[j] = ldg_u64_nc(crow[j] + tipk + 8);
      }
    }
    sp += tip_block;
  }
  // t = P . l for the next internal child, l = current registers (REG) or a stack slot
  template <bool FROM_SLOT>
  __device__ __forceinline__ void term_internal(int slot, double (&t)[PT][4], int (&e)[PT]) {
    const double* Pm = sp + (c << 2);
    double p[4][4];
#pragma unroll
    for (int x = 0; x < 4; ++x) ld256nc(Pm + x * 4 * C, p[x][0], p[x][1], p[x][2], p[x][3]);
    sp += 16 * C;
#pragma unroll
    for (int j = 0; j < PT; ++j) {
      double l0, l1, l2, l3;
      if (FROM_SLOT) {
        const double4 s = st[(slot * PT + j) * NTH + tid];
        l0 = s.x; l1 = s.y; l2 = s.z; l3 = s.w;
        e[j] += ste[(slot * PT + j) * NTH + tid];
      } else {
        l0 = v[j][0]; l1 = v[j][1]; l2 = v[j][2]; l3 = v[j][3];
        e[j] += E[j];
      }
#pragma unroll
      for (int x = 0; x < 4; ++x) t[j][x] = fma(p[x][3], l3, fma(p[x][2], l2, fma(p[x][1], l1, p[x][0] * l0)));
    }
  }
  template <int K>
  __device__ __forceinline__ void term(int slot, double (&t)[PT][4], int (&e)[PT]) {
    if (K == 0) term_tip(t);
    else if (K == 1) term_internal<true>(slot, t, e);
    else term_internal<false>(slot, t, e);
  }
  // v = a, rescaled per row; E = e (+ shift)
  __device__ __forceinline__ void commit(double (&a)[PT][4], int (&e)[PT]) {
    int m[PT];
    int mmin = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < PT; ++j) {
      m[j] = max(max(hi_word(a[j][0]), hi_word(a[j][1])), max(hi_word(a[j][2]), hi_word(a[j][3])));
      mmin = min(mmin, m[j]);
    }
    if (mmin < kScaleThresholdHi) {  // rare: some row of this thread is small (or identically zero)
#pragma unroll
      for (int j = 0; j < PT; ++j) {
        if (m[j] < kScaleThresholdHi && m[j] >= (1 << 20)) {
          const int k = rescale_shift(m[j]);
          const double f = pow2(k);
          a[j][0] *= f; a[j][1] *= f; a[j][2] *= f; a[j][3] *= f;
          e[j] += k;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < PT; ++j) {
      v[j][0] = a[j][0]; v[j][1] = a[j][1]; v[j][2] = a[j][2]; v[j][3] = a[j][3];
      E[j] = e[j];
    }
  }
  template <int KA, int KB>
  __device__ __forceinline__ void op2(int slotA, int slotB) {
    double ta[PT][4], tb[PT][4];
    int e[PT];
#pragma unroll
    for (int j = 0; j < PT; ++j) e[j] = 0;
    term<KA>(slotA, ta, e);
    term<KB>(slotB, tb, e);
#pragma unroll
    for (int j = 0; j < PT; ++j) {
      ta[j][0] *= tb[j][0]; ta[j][1] *= tb[j][1]; ta[j][2] *= tb[j][2]; ta[j][3] *= tb[j][3];
    }
    commit(ta, e);
  }
  __device__ __forceinline__ void op_generic(int nchild, unsigned long long toks) {
    double a[PT][4];
    int e[PT];
#pragma unroll
    for (int j = 0; j < PT; ++j) e[j] = 0;
#pragma unroll 1
    for (int ch = 0; ch < nchild; ++ch, toks >>= 8) {
      const int kind = (int)(toks >> 6) & 3;
      double t[PT][4];
      if (kind == CHILD_TIP) term_tip(t);
      else if (kind == CHILD_SLOT) term_internal<true>((int)toks & 63, t, e);
      else term_internal<false>(0, t, e);
#pragma unroll
      for (int j = 0; j < PT; ++j) {
        if (ch == 0) {
          a[j][0] = t[j][0]; a[j][1] = t[j][1]; a[j][2] = t[j][2]; a[j][3] = t[j][3];
        } else {
          a[j][0] *= t[j][0]; a[j][1] *= t[j][1]; a[j][2] *= t[j][2]; a[j][3] *= t[j][3];
        }
      }
    }
    commit(a, e);
  }
};

template <int C_LOG2, int PT, bool KEEP>
__global__ void __launch_bounds__(kWalk4Threads) walk4_kernel(Walk4Params prm) {
  constexpr int C = 1 << C_LOG2;
  constexpr int NTH = kWalk4Threads;
  constexpr int GROUPS = NTH >> C_LOG2;  // pattern groups per CTA
  extern __shared__ __align__(32) unsigned char smem_raw[];
  __shared__ double red[32];

  Walk4State<C_LOG2, PT> s;
  s.st = reinterpret_cast<double4*>(smem_raw);                                      // [nslots][PT][NTH]
  s.ste = reinterpret_cast<int*>(smem_raw + (size_t)prm.nslots * PT * NTH * 32);    // [nslots][PT][NTH]
  const int tid = threadIdx.x;
  s.tid = tid;
  const int c = tid & (C - 1);
  s.c = c;
  const long long pat0 = ((long long)blockIdx.x * GROUPS + (tid >> C_LOG2)) * PT;
  const long long rows = prm.N << C_LOG2;
#pragma unroll
  for (int j = 0; j < PT; ++j) {
    const long long pj = pat0 + j < prm.N ? pat0 + j : prm.N - 1;
    s.crow[j] = prm.codesT + (size_t)pj * prm.tstride;
    s.q[j] = ldg_u64_nc(s.crow[j]);
    s.qn[j] = ldg_u64_nc(s.crow[j] + 8);
    s.v[j][0] = s.v[j][1] = s.v[j][2] = s.v[j][3] = 1.0;
    s.E[j] = 0;
  }
  s.tipk = 0;
  s.sp = prm.stream;
  s.tip_block = (prm.ncodes << C_LOG2) * 4;

  unsigned long long d = prm.desc[0];
  for (int o = 0; o < prm.n_ops; ++o) {
    const unsigned long long dn = prm.desc[o + 1];  // desc has a zero sentinel at [n_ops]
    const int shape = (int)(d & 0xfu);
    const int dst = (int)((d >> 8) & 0xffu);
    const int slotA = (int)(d >> 16) & 63, slotB = (int)(d >> 24) & 63;
    switch (shape) {
      case 1: s.template op2<0, 0>(slotA, slotB); break;
      case 2: s.template op2<0, 1>(slotA, slotB); break;
      case 3: s.template op2<0, 2>(slotA, slotB); break;
      case 4: s.template op2<1, 0>(slotA, slotB); break;
      case 6: s.template op2<1, 2>(slotA, slotB); break;
      case 7: s.template op2<2, 0>(slotA, slotB); break;
      case 8: s.template op2<2, 1>(slotA, slotB); break;
      default: s.op_generic((int)(d >> 4) & 0xf, d >> 16); break;
    }
    if (dst) {
#pragma unroll
      for (int j = 0; j < PT; ++j) {
        s.st[((dst - 1) * PT + j) * NTH + tid] = make_double4(s.v[j][0], s.v[j][1], s.v[j][2], s.v[j][3]);
        s.ste[((dst - 1) * PT + j) * NTH + tid] = s.E[j];
      }
    }
    if (KEEP) {
#pragma unroll
      for (int j = 0; j < PT; ++j) {
        if (pat0 + j < prm.N) {
          const long long r = ((pat0 + j) << C_LOG2) + c;
          st256(prm.keep + ((size_t)o * rows + r) * 4, s.v[j][0], s.v[j][1], s.v[j][2], s.v[j][3]);
          prm.keep_exp[(size_t)o * rows + r] = s.E[j];
        }
      }
    }
    d = dn;
  }

  // ---- root reduction: L_i = sum_c p_c 2^-(E_c - Emin) sum_x pi_x CLV_root[i][c][x] ----------
  const bool rsem = prm.flags & 1u;
  const double f0 = prm.rootfreq[0], f1 = prm.rootfreq[1], f2 = prm.rootfreq[2], f3 = prm.rootfreq[3];
  const double pc = prm.probs[c];
  double contrib = 0.0;
#pragma unroll
  for (int j = 0; j < PT; ++j) {
    int Emin = s.E[j];
#pragma unroll
    for (int off = 1; off < C; off <<= 1) Emin = min(Emin, __shfl_xor_sync(0xffffffffu, Emin, off));
    const double t0 = s.v[j][0] * f0, t1 = s.v[j][1] * f1, t2 = s.v[j][2] * f2, t3 = s.v[j][3] * f3;
    double sum;
    if (rsem) sum = (t0 > 0 ? t0 : 0.0) + (t1 > 0 ? t1 : 0.0) + (t2 > 0 ? t2 : 0.0) + (t3 > 0 ? t3 : 0.0);
    else sum = ((t0 + t1) + t2) + t3;
    double L = sum * align_factor(s.E[j] - Emin) * pc;
    if (rsem && !(L > 0)) L = 0.0;
#pragma unroll
    for (int off = 1; off < C; off <<= 1) L += __shfl_xor_sync(0xffffffffu, L, off);
    if (!rsem && L < 0) L = 0.0;
    if (pat0 + j < prm.N && c == 0) {
      const long long pat = pat0 + j;
      const double lnl = log(L) - (double)Emin * kLn2;
      prm.SR[pat] = L;
      prm.rexp[pat] = Emin;
      prm.site_lnl[pat] = lnl;
      contrib += prm.weights[pat] * lnl;
    }
  }
  const double bs = block_sum(contrib, red);
  if (tid == 0) prm.partials[blockIdx.x] = bs;
}

// codes [nl][N] (leaf-slot major) -> codesT [N][tstride], column k = k-th tip the walk consumes
__global__ void transpose_codes_kernel(const unsigned char* codes, const int* tip_order, int ntips, long long N,
                                       int tstride, unsigned char* codesT) {
  const long long pat = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pat >= N) return;
  unsigned char* row = codesT + (size_t)pat * tstride;
  for (int k0 = 0; k0 < tstride; k0 += 8) {
    unsigned long long w = 0;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const int k = k0 + b;
      if (k < ntips) w |= (unsigned long long)codes[(size_t)tip_order[k] * N + pat] << (8 * b);
    }
    *reinterpret_cast<unsigned long long*>(row + k0) = w;
  }
}

// stream packing for S = 4: one CTA per child block
struct PackBlock {
  int kind;       // CHILD_TIP or internal (anything else)
  int pnode;      // node whose branch carries the child
  long long off;  // offset of the block in the stream (doubles)
};
__global__ void pack_stream4_kernel(const PackBlock* blocks, const double* P /*[nn][C][4][4]*/, const double* code_table,
                                    int C, int ncodes, double* stream) {
  const PackBlock b = blocks[blockIdx.x];
  const double* Pn = P + (size_t)b.pnode * C * 16;
  double* out = stream + b.off;
  if (b.kind == CHILD_TIP) {
    for (int e = threadIdx.x; e < ncodes * C * 4; e += blockDim.x) {
      const int x = e & 3, c = (e >> 2) % C, code = (e >> 2) / C;
      const double* tv = code_table + code * 4;
      const double* pr = Pn + c * 16 + x * 4;
      out[e] = fma(pr[3], tv[3], fma(pr[2], tv[2], fma(pr[1], tv[1], pr[0] * tv[0])));
    }
  } else {
    for (int e = threadIdx.x; e < C * 16; e += blockDim.x) {
      const int y = e & 3, c = (e >> 2) % C, x = (e >> 2) / C;
      out[e] = Pn[c * 16 + x * 4 + y];
    }
  }
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
          e = prm.gstack_exp[(size_t)ch.idx * rows + r];
        } else {
          load_row<S>(prm.keep + ((size_t)ch.idx * rows + r) * S, l);
          e = prm.keep_exp[(size_t)ch.idx * rows + r];
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
      prm.gstack_exp[(size_t)op.dst_slot * rows + r] = E;
    }
    if (op.keep_idx >= 0 && valid) {
      store_row<S>(prm.keep + ((size_t)op.keep_idx * rows + r) * S, v);
      prm.keep_exp[(size_t)op.keep_idx * rows + r] = E;
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
  int Emin = E;
#pragma unroll
  for (int off = 1; off < C; off <<= 1) Emin = min(Emin, __shfl_xor_sync(0xffffffffu, Emin, off));
  double Lk = s * align_factor(E - Emin) * prm.probs[c];
  if (rsem && !(Lk > 0)) Lk = 0.0;
#pragma unroll
  for (int off = 1; off < C; off <<= 1) Lk += __shfl_xor_sync(0xffffffffu, Lk, off);
  if (!rsem && Lk < 0) Lk = 0.0;
  double contrib = 0.0;
  if (valid && c == 0) {
    const double lnl = log(Lk) - (double)Emin * kLn2;
    prm.SR[pat] = Lk;
    prm.rexp[pat] = Emin;
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
