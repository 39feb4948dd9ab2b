// K2 + K3 for S = 4 (DNA), value-only evaluations: the CLASS-UNIFORM, CHUNK-STREAMED tree walk.
//
// Same idea as walk4_kernel (walk_kernels.cuh): the pruning recursion (RHomogeneousTreeLikelihood::computeSubtreeLikelihood,
// Likelihood/RHomogeneousTreeLikelihood.cpp:802-863; DR twin DRHomogeneousTreeLikelihood.cpp:483-539, :819-864) is independent
// per (pattern, rate class) row, so one launch walks the whole post-order program and every CLV stays on chip; the root
// reduction (RHomogeneousTreeLikelihood.cpp:162-216 / DRHomogeneousTreeLikelihood.cpp:653-719) is fused.
//
// What ncu showed on walk4_kernel (profiles/r1_final_walk4_pt2_ncu_summary.csv): l1tex data-pipe wavefronts 82.6 %, FP64 pipe
// 24 %, long-scoreboard stalls 4.4 per issue -- every thread fetched the 128-byte P block of its class from global memory for
// each internal child (32 wavefronts per warp for 2 patterns), tip rows and the shared-memory stack went through the same pipe.
// This kernel removes that traffic instead of hiding it:
//   * a warp is CLASS-UNIFORM (32 lanes = 32 x PT patterns of ONE rate class), so P of a child is the same 16 doubles for
//     the whole warp: 8 broadcast LDS.128 (1 wavefront each) instead of 4 LDG.256 x 8 wavefronts;
//   * the walk's tables (P blocks, per-branch tip tables, op descriptors) are cut into fixed-size CHUNKS in walk order and
//     streamed L2 -> shared memory by the TMA engine (cp.async.bulk + mbarrier full/empty ring, one elected producer thread):
//     no dependent global load is left on the critical path, the chunk after next is always in flight;
//   * a tip row is a gather from a 128-byte [code][4] table in shared memory: every bank holds one address, 2 LDS.128;
//   * the innermost level of the CLV stack lives in REGISTERS: a push whose lifetime contains no other push ("leaf push":
//     the sibling evaluated next is a caterpillar) writes its result to the second register set `w`, and the father reads it
//     from there; only deeper pushes go to the shared-memory stack (two 16-byte planes: conflict free);
//   * tip codes are read coalesced: codes8[group of 8 tips][pattern] (one 8-byte word per pattern and 8 tips).
// Arithmetic, multiplication order and rescaling rule are those of walk4_kernel: the two kernels agree bit for bit per site.
#pragma once
#include "walk_kernels.cuh"

namespace bppgpu {

constexpr int kW4cThreads = 256;
constexpr int kW4cWarps = kW4cThreads / 32;
constexpr int kW4cStages = 3;
constexpr int kW4cHeader = 128;        // bytes at the head of a chunk: u32 n_ops, u32 pad, up to 15 descriptors
constexpr int kW4cMaxOpsPerChunk = 15;
constexpr int kW4cRegSlot = 63;        // dst code of the register slot

enum W4cKind { W4C_TIP = 0, W4C_SLOT = 1, W4C_REG = 2, W4C_RSL = 3 };
// binary shapes (the only ones a bifurcating tree's planner emits); 0 = generic child loop
enum W4cShape { W4C_GENERIC = 0, W4C_TT = 1, W4C_TR = 2, W4C_RT = 3, W4C_SR = 4, W4C_RS = 5, W4C_WR = 6, W4C_RW = 7 };

// descriptor: bits 0-2 shape, bit 3 result goes to the register slot, bits 4-7 nchild, bits 8-15 dst (0 none, s+1 shared slot,
// 63 register slot), byte 2+j child token kind<<6 | slot
struct Walk4cParams {
  const unsigned char* stream;       // [nchunks][CH] this point's chunks
  int nchunks, CH, nslots, ncodes;
  unsigned flags;                    // bit0: R semantics at the root
  long long N, Npad;
  const unsigned long long* codes8;  // [ntip8 + 2][Npad]
  const double* rootfreq;            // [4]
  const double* probs;               // [C]
  const double* weights;             // [N]
  double* SR;                        // [N]
  int* rexp;                         // [N]
  double* site_lnl;                  // [N]
  double* partials;                  // [gridDim.x]
};

// ---- mbarrier / TMA bulk copy (PTX ISA 8.x, sm_90+) ------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "W4C_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra W4C_DONE_%=;\n"
      "bra W4C_WAIT_%=;\n"
      "W4C_DONE_%=:\n"
      "}\n" ::"r"(smem_u32(b)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(b))
               : "memory");
}

// explicit shared-space accesses on 32-bit shared addresses (no generic-address LD/ST, no 64-bit pointer arithmetic)
__device__ __forceinline__ double2 lds128(unsigned a) {
  double2 r;
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"(a));
  return r;
}
__device__ __forceinline__ void sts128(unsigned a, double x, double y) {
  asm volatile("st.shared.v2.f64 [%2], {%0,%1};" ::"d"(x), "d"(y), "r"(a) : "memory");
}
__device__ __forceinline__ int lds32(unsigned a) {
  int r;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(r) : "r"(a));
  return r;
}
__device__ __forceinline__ void sts32(unsigned a, int x) { asm volatile("st.shared.s32 [%1], %0;" ::"r"(x), "r"(a) : "memory"); }
__device__ __forceinline__ unsigned long long lds64u(unsigned a) {
  unsigned long long r;
  asm volatile("ld.shared.u64 %0, [%1];" : "=l"(r) : "r"(a));
  return r;
}

template <int C_LOG2, int PT>
struct W4cState {
  static constexpr int C = 1 << C_LOG2;
  static constexpr int NTH = kW4cThreads;
  double v[PT][4], w[PT][4];   // current CLV rows / register slot
  int E[PT], EW[PT];
  unsigned long long q[PT], qn[PT];
  const unsigned long long* crow[PT];
  long long Npad;
  int tipk;
  unsigned sp;                 // shared address of the next table block of the current chunk
  int tip_off;                 // c * ncodes * 32 bytes
  int tip_block;               // C * ncodes * 32 bytes
  int c, tid;
  unsigned stA, stB, ste;      // shared addresses of the stack planes, already offset by this thread's lane

  __device__ __forceinline__ void term_tip(double (&t)[PT][4]) {
    const unsigned tb = sp + tip_off;
#pragma unroll
    for (int j = 0; j < PT; ++j) {
      const unsigned code = (unsigned)(q[j] & 0xffu);
      q[j] >>= 8;
      const double2 a = lds128(tb + (code << 5));
      const double2 b = lds128(tb + (code << 5) + 16);
      t[j][0] = a.x; t[j][1] = a.y; t[j][2] = b.x; t[j][3] = b.y;
    }
    if (((++tipk) & 7) == 0) {
#pragma unroll
      for (int j = 0; j < PT; ++j) {
        q[j] = qn[j];
        qn[j] = __ldg(crow[j] + (size_t)((tipk >> 3) + 1) * Npad);
      }
    }
    sp += tip_block;
  }
  // t = P . l, l from: the current rows (W4C_REG), the register slot (W4C_RSL) or a shared-memory slot (W4C_SLOT)
  template <int SRC>
  __device__ __forceinline__ void term_internal(int slot, double (&t)[PT][4], int (&e)[PT]) {
    const unsigned Pm = sp + (c << 7);
    sp += 128 * C;
    double l[PT][4];
#pragma unroll
    for (int j = 0; j < PT; ++j) {
      if (SRC == W4C_SLOT) {
        const unsigned k = (unsigned)((slot * PT + j) * NTH);
        const double2 a = lds128(stA + k * 16), b = lds128(stB + k * 16);
        l[j][0] = a.x; l[j][1] = a.y; l[j][2] = b.x; l[j][3] = b.y;
        e[j] += lds32(ste + k * 4);
      } else if (SRC == W4C_RSL) {
        l[j][0] = w[j][0]; l[j][1] = w[j][1]; l[j][2] = w[j][2]; l[j][3] = w[j][3];
        e[j] += EW[j];
      } else {
        l[j][0] = v[j][0]; l[j][1] = v[j][1]; l[j][2] = v[j][2]; l[j][3] = v[j][3];
        e[j] += E[j];
      }
    }
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      const double2 p01 = lds128(Pm + 32 * x);        // warp-uniform address: broadcast
      const double2 p23 = lds128(Pm + 32 * x + 16);
#pragma unroll
      for (int j = 0; j < PT; ++j) t[j][x] = fma(p23.y, l[j][3], fma(p23.x, l[j][2], fma(p01.y, l[j][1], p01.x * l[j][0])));
    }
  }
  template <int K>
  __device__ __forceinline__ void term(int slot, double (&t)[PT][4], int (&e)[PT]) {
    if (K == W4C_TIP) term_tip(t);
    else term_internal<K>(slot, t, e);
  }
  // rescale per row, then store to v (or to the register slot)
  template <bool DSTW>
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
      if (DSTW) {
        w[j][0] = a[j][0]; w[j][1] = a[j][1]; w[j][2] = a[j][2]; w[j][3] = a[j][3];
        EW[j] = e[j];
      } else {
        v[j][0] = a[j][0]; v[j][1] = a[j][1]; v[j][2] = a[j][2]; v[j][3] = a[j][3];
        E[j] = e[j];
      }
    }
  }
  template <int KA, int KB, bool DSTW>
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
    commit<DSTW>(ta, e);
  }
  __device__ __forceinline__ void op_generic(int nchild, unsigned long long toks, bool dstw) {
    double a[PT][4];
    int e[PT];
#pragma unroll
    for (int j = 0; j < PT; ++j) e[j] = 0;
#pragma unroll 1
    for (int ch = 0; ch < nchild; ++ch, toks >>= 8) {
      const int kind = (int)(toks >> 6) & 3;
      double t[PT][4];
      if (kind == W4C_TIP) term_tip(t);
      else if (kind == W4C_SLOT) term_internal<W4C_SLOT>((int)toks & 63, t, e);
      else if (kind == W4C_RSL) term_internal<W4C_RSL>(0, t, e);
      else term_internal<W4C_REG>(0, t, e);
#pragma unroll
      for (int j = 0; j < PT; ++j) {
        if (ch == 0) {
          a[j][0] = t[j][0]; a[j][1] = t[j][1]; a[j][2] = t[j][2]; a[j][3] = t[j][3];
        } else {
          a[j][0] *= t[j][0]; a[j][1] *= t[j][1]; a[j][2] *= t[j][2]; a[j][3] *= t[j][3];
        }
      }
    }
    if (dstw) commit<true>(a, e);
    else commit<false>(a, e);
  }
};

// dynamic shared memory of one CTA
__host__ __device__ inline size_t walk4c_smem_bytes(int CH, int nslots, int PT) {
  return (size_t)kW4cStages * CH + (size_t)(nslots > 0 ? nslots : 0) * PT * kW4cThreads * 36 + (size_t)PT * kW4cThreads * 12 + 128;
}

template <int C_LOG2, int PT>
__global__ void __launch_bounds__(kW4cThreads, (PT <= 2 ? 2 : 1)) walk4c_kernel(Walk4cParams prm) {
  constexpr int C = 1 << C_LOG2;
  constexpr int NTH = kW4cThreads;
  constexpr int G = kW4cWarps >> C_LOG2;   // pattern groups (of 32 * PT patterns) per CTA
  constexpr int PPC = G * 32 * PT;         // patterns per CTA
  extern __shared__ __align__(32) unsigned char smem_raw[];
  __shared__ double red[32];

  const int CH = prm.CH;
  unsigned char* ring = smem_raw;
  const unsigned ring_s = smem_u32(smem_raw);
  unsigned char* q0 = smem_raw + (size_t)kW4cStages * CH;
  const size_t plane = (size_t)(prm.nslots > 0 ? prm.nslots : 0) * PT * NTH;
  W4cState<C_LOG2, PT> s;
  s.stA = smem_u32(q0) + threadIdx.x * 16;
  s.stB = smem_u32(q0 + plane * 16) + threadIdx.x * 16;
  s.ste = smem_u32(q0 + plane * 32) + threadIdx.x * 4;
  double* rsum = reinterpret_cast<double*>(q0 + plane * 36);          // [C][PPC] class terms of the root reduction
  int* rexpn = reinterpret_cast<int*>(q0 + plane * 36 + (size_t)PT * NTH * 8);   // [C][PPC]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(q0 + plane * 36 + (size_t)PT * NTH * 12);  // full[], empty[]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = warp & (C - 1), g = warp >> C_LOG2;
  s.tid = tid;
  s.c = c;
  s.Npad = prm.Npad;
  s.tip_off = c * prm.ncodes * 32;
  s.tip_block = C * prm.ncodes * 32;
  const long long pat0 = (long long)blockIdx.x * PPC + g * (32 * PT) + lane;   // pattern j: pat0 + 32 j
#pragma unroll
  for (int j = 0; j < PT; ++j) {
    long long pj = pat0 + 32 * j;
    if (pj >= prm.N) pj = prm.N - 1;
    s.crow[j] = prm.codes8 + pj;
    s.q[j] = __ldg(s.crow[j]);
    s.qn[j] = __ldg(s.crow[j] + prm.Npad);
    s.v[j][0] = s.v[j][1] = s.v[j][2] = s.v[j][3] = 1.0;
    s.w[j][0] = s.w[j][1] = s.w[j][2] = s.w[j][3] = 1.0;
    s.E[j] = 0;
    s.EW[j] = 0;
  }
  s.tipk = 0;

  unsigned long long* full = bars;
  unsigned long long* empty = bars + kW4cStages;
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < kW4cStages; ++i) {
      mbar_init(full + i, 1);
      mbar_init(empty + i, kW4cWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const int nchunks = prm.nchunks;
  if (tid == 0) {
    for (int k = 0; k < kW4cStages && k < nchunks; ++k) {
      mbar_expect_tx(full + k, (unsigned)CH);
      bulk_g2s(ring + (size_t)k * CH, prm.stream + (size_t)k * CH, (unsigned)CH, full + k);
    }
  }

  int stage = 0;
  unsigned phase = 0;
  for (int k = 0; k < nchunks; ++k) {
    mbar_wait(full + stage, phase);
    const unsigned cb = ring_s + (unsigned)stage * (unsigned)CH;
    const int n_ops = lds32(cb);
    s.sp = cb + kW4cHeader;
    for (int o = 0; o < n_ops; ++o) {
      const unsigned long long d = lds64u(cb + 8 + 8 * o);
      const int shape = (int)(d & 0xfu);   // bit 3 = result to the register slot
      const int dst = (int)((d >> 8) & 0xffu);
      const int slotA = (int)(d >> 16) & 63, slotB = (int)(d >> 24) & 63;
      switch (shape) {
        case W4C_TT: s.template op2<W4C_TIP, W4C_TIP, false>(0, 0); break;
        case W4C_TR: s.template op2<W4C_TIP, W4C_REG, false>(0, 0); break;
        case W4C_RT: s.template op2<W4C_REG, W4C_TIP, false>(0, 0); break;
        case W4C_SR: s.template op2<W4C_SLOT, W4C_REG, false>(slotA, 0); break;
        case W4C_RS: s.template op2<W4C_REG, W4C_SLOT, false>(0, slotB); break;
        case W4C_WR: s.template op2<W4C_RSL, W4C_REG, false>(0, 0); break;
        case W4C_RW: s.template op2<W4C_REG, W4C_RSL, false>(0, 0); break;
        case 8 + W4C_TT: s.template op2<W4C_TIP, W4C_TIP, true>(0, 0); break;
        case 8 + W4C_TR: s.template op2<W4C_TIP, W4C_REG, true>(0, 0); break;
        case 8 + W4C_RT: s.template op2<W4C_REG, W4C_TIP, true>(0, 0); break;
        case 8 + W4C_SR: s.template op2<W4C_SLOT, W4C_REG, true>(slotA, 0); break;
        case 8 + W4C_RS: s.template op2<W4C_REG, W4C_SLOT, true>(0, slotB); break;
        case 8 + W4C_WR: s.template op2<W4C_RSL, W4C_REG, true>(0, 0); break;
        case 8 + W4C_RW: s.template op2<W4C_REG, W4C_RSL, true>(0, 0); break;
        default: s.op_generic((int)(d >> 4) & 0xf, d >> 16, (shape & 8) != 0); break;
      }
      if (dst && dst != kW4cRegSlot) {
#pragma unroll
        for (int j = 0; j < PT; ++j) {
          const unsigned k = (unsigned)(((dst - 1) * PT + j) * NTH);
          sts128(s.stA + k * 16, s.v[j][0], s.v[j][1]);
          sts128(s.stB + k * 16, s.v[j][2], s.v[j][3]);
          sts32(s.ste + k * 4, s.E[j]);
        }
      }
    }
    // this warp is done with the stage; the producer refills the stage of the PREVIOUS chunk (one chunk of slack between
    // the slowest warp and the producer's warp), so chunks k+1 .. k+kW4cStages-2 are always loaded or in flight
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + stage);
    if (tid == 0 && k >= 1 && k - 1 + kW4cStages < nchunks) {
      const int ps = stage == 0 ? kW4cStages - 1 : stage - 1;          // stage of chunk k-1
      const unsigned pph = stage == 0 ? phase ^ 1u : phase;             // its phase
      mbar_wait(empty + ps, pph);
      mbar_expect_tx(full + ps, (unsigned)CH);
      bulk_g2s(ring + (size_t)ps * CH, prm.stream + (size_t)(k - 1 + kW4cStages) * CH, (unsigned)CH, full + ps);
    }
    if (++stage == kW4cStages) { stage = 0; phase ^= 1u; }
  }

  // ---- root reduction: L_i = sum_c p_c 2^-(E_c - Emin) sum_x pi_x CLV_root[i][c][x] (association as in walk4_kernel) ----
  const bool rsem = prm.flags & 1u;
  const double f0 = prm.rootfreq[0], f1 = prm.rootfreq[1], f2 = prm.rootfreq[2], f3 = prm.rootfreq[3];
#pragma unroll
  for (int j = 0; j < PT; ++j) {
    const double t0 = s.v[j][0] * f0, t1 = s.v[j][1] * f1, t2 = s.v[j][2] * f2, t3 = s.v[j][3] * f3;
    double sum;
    if (rsem) sum = (t0 > 0 ? t0 : 0.0) + (t1 > 0 ? t1 : 0.0) + (t2 > 0 ? t2 : 0.0) + (t3 > 0 ? t3 : 0.0);
    else sum = ((t0 + t1) + t2) + t3;
    const int lp = g * (32 * PT) + 32 * j + lane;   // pattern index inside the CTA
    rsum[c * PPC + lp] = sum;
    rexpn[c * PPC + lp] = s.E[j];
  }
  __syncthreads();
  double contrib = 0.0;
  for (int lp = tid; lp < PPC; lp += NTH) {
    const long long pat = (long long)blockIdx.x * PPC + lp;
    if (pat >= prm.N) continue;
    int Emin = rexpn[lp];
#pragma unroll
    for (int cc = 1; cc < C; ++cc) Emin = min(Emin, rexpn[cc * PPC + lp]);
    double Lc[C];
#pragma unroll
    for (int cc = 0; cc < C; ++cc) {
      double L = rsum[cc * PPC + lp] * align_factor(rexpn[cc * PPC + lp] - Emin) * prm.probs[cc];
      if (rsem && !(L > 0)) L = 0.0;
      Lc[cc] = L;
    }
#pragma unroll
    for (int off = 1; off < C; off <<= 1)      // the xor-butterfly order of walk4_kernel, as seen by class 0
#pragma unroll
      for (int cc = 0; cc < C; cc += 2 * off) Lc[cc] += Lc[cc + off];
    double L = Lc[0];
    if (!rsem && L < 0) L = 0.0;
    const double lnl = log(L) - (double)Emin * kLn2;
    prm.SR[pat] = L;
    prm.rexp[pat] = Emin;
    prm.site_lnl[pat] = lnl;
    contrib += prm.weights[pat] * lnl;
  }
  const double bs = block_sum(contrib, red);
  if (tid == 0) prm.partials[blockIdx.x] = bs;
}

// codes [nl][N] (leaf-slot major) -> codes8 [group][Npad] (one 8-byte word = the 8 tips k0..k0+7 of the consumption order)
__global__ void pack_codes8_kernel(const unsigned char* codes, const int* tip_order, int ntips, long long N, long long Npad,
                                   int ngroups, unsigned long long* codes8) {
  const long long pat = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pat >= Npad) return;
  for (int g = 0; g < ngroups; ++g) {
    unsigned long long wv = 0;
    if (pat < N) {
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const int k = g * 8 + b;
        if (k < ntips) wv |= (unsigned long long)codes[(size_t)tip_order[k] * N + pat] << (8 * b);
      }
    }
    codes8[(size_t)g * Npad + pat] = wv;
  }
}

// table blocks of the chunked stream: one CTA per child block
struct Pack4cBlock {
  int kind;       // W4C_TIP or internal
  int pnode;      // node whose branch carries the child
  long long off;  // BYTE offset of the block in the chunked stream
};
__global__ void pack_stream4c_kernel(const Pack4cBlock* blocks, const double* P /*[nn][C][4][4]*/, const double* code_table,
                                     int C, int ncodes, unsigned char* stream) {
  const Pack4cBlock b = blocks[blockIdx.x];
  const double* Pn = P + (size_t)b.pnode * C * 16;
  double* out = reinterpret_cast<double*>(stream + b.off);
  if (b.kind == W4C_TIP) {   // [class][code][x]
    for (int e = threadIdx.x; e < ncodes * C * 4; e += blockDim.x) {
      const int x = e & 3, code = (e >> 2) % ncodes, c = (e >> 2) / ncodes;
      const double* tv = code_table + code * 4;
      const double* pr = Pn + c * 16 + x * 4;
      out[e] = fma(pr[3], tv[3], fma(pr[2], tv[2], fma(pr[1], tv[1], pr[0] * tv[0])));
    }
  } else {                   // [class][x][y] = the reference's pxy_ order
    for (int e = threadIdx.x; e < C * 16; e += blockDim.x) out[e] = Pn[e];
  }
}

}  // namespace bppgpu
