// K4 + K5 for protein-sized alphabets (S = 20) on the FP64 tensor cores, fused per FATHER instead of per branch.
//
// For a father f with sons s_0 .. s_{k-1} (k <= 3) one launch produces, for every son,
//   upper[s_j][i][c][x] = A_f[i][c][x] * prod_{o != j} M_o[i][c][x]
//       A_f[x] = f == root ? pi_x : sum_y P_f[c][y][x] upper[f][i][c][y]      (P transposed,
//                DRHomogeneousTreeLikelihood::computeSubtreeLikelihoodPrefix, Likelihood/DRHomogeneousTreeLikelihood.cpp:543-649, :919-945)
//       M_o[x] = sum_y P_o[c][x][y] lower_o[i][c][y]                          (computeLikelihoodFromArrays :819-864)
//   dL_i  = sum_c p_c 2^(eR_i - eU - eL) sum_x upper[s_j][x] sum_y dpxy_j[c][x][y]  lower_j[y] / SR_i    (:287-326)
//   d2L_i = ... d2pxy_j ...                                                                             (:373-411)
// and the weighted sums  sum_i w_i dL_i,  sum_i w_i (d2L_i - dL_i^2)  (:340-368, :425-454) as per-CTA partials.
// With nh_form, states x whose M_j[x] is exactly 0 are dropped (DRNonHomogeneousTreeLikelihood.cpp:396-407).
//
// Why per father: the per-branch kernel (dmma_deriv_kernels.cuh) re-reads upper[f] and re-does the father contraction
// for every son, reads each lower CLV twice (as the branch's own and as the sibling's) and needs three launches per
// branch.  Here every CLV slab is read exactly once, A_f is contracted once, and the three matrices of a son that
// share its lower CLV as the A operand -- P, dP, d2P -- are stacked along N into ONE 60-column B operand (8 column
// blocks instead of 3 x 3), so a binary father costs 15 + 2 x 40 DMMAs per 8 rows instead of 2 x 60.
//
// Column ownership: in the m8n8k4 accumulator lane q of a quad owns columns 2q, 2q+1 of every 8-column block.  The
// staged B operand is permuted so that lane q owns the SAME five states of all three matrices -- x = 4q .. 4q+3 and
// 16 + q, i.e. exactly the states whose A fragments it loads -- slot t = 5 mat + i living in block t / 2, half t % 2.
// M, dM, d2M, A_f and the rescaled upper row are therefore element-aligned in registers, every product / dot product
// of the epilogue is lane-local, a row is stored with one 32-byte and one 8-byte access per lane (a quad writes 128
// contiguous bytes), and tip sons read their rows of the ordinary [code][state] tip tables the same way.
// One persistent CTA per SM owns a contiguous range of patterns; its warps take 8-pattern row blocks round-robin and
// run the rate classes of a block back to back.  The operands of ALL classes stay in shared memory for the whole
// launch (C <= 4: 25.6 KB per class for a binary father), so there is no barrier after the prologue, the class sums
// of dL / d2L are accumulated in registers, and the pattern-level combination w (d2L - dL^2) happens in the same
// kernel; one finaliser launch per evaluation sums the per-CTA partials of every branch.
#pragma once
#include "dmma_node_kernels.cuh"

namespace bppgpu {

constexpr int kFamMaxSons = 3;
constexpr int kFamMaxClasses = 4; // operands of every class are resident in shared memory
constexpr int kFamKB = 5;        // k blocks: S = 20
// warps per SM by kernel kind: binary fathers 12 (16 warps at 128 registers measured 8 % slower); the run-time kind (up to three sons:
// 32 more accumulator registers and 10 KB more operands per class) 6
__host__ __device__ constexpr int fam_threads_kind(int KIND) { return KIND == 4 ? 192 : 384; }

// Everything the kernel dereferences is a ready-made pointer (the host folds slab indices in): the first version
// rebuilt 64-bit slab offsets per access and spent more issue slots on IMAD than on DMMA.
struct FamilySon {
  int kind;              // CHILD_TIP / CHILD_KEEP
  int node;              // node id (branch index into the packed operands and outputs)
  const double* clv;     // lower CLV slab [N][C][S]            (internal son)
  const int* exp;        // its exponents  [N][C]
  const void* codes;     // tip codes      [N]                  (tip son)
  const double* tpack;   // this leaf's tip tables of P | dP | d2P in lane order [C][ncodes][4][4][4]   (tip son)
  double* up;            // slab that receives upper[son], or null
  int* upexp;
};

struct DmmaFamilyParams {
  FamilySon sons[kFamMaxSons];
  int nson;
  int father;            // node id of f, or -1 when f is the root
  const double* fup;     // upper[f] slab
  const int* fupexp;
  int S, C, ncodes, code_bytes;
  int nh_form;
  int ppc;               // patterns per CTA (multiple of 8)
  int prow, crow;        // CLV row of (pattern i, class c) = i * prow + c * crow  ([i][c]: C, 1;  class-major [c][i]: 1, N)
  long long N;
  const double* packA;   // [nn][C][kFamPackA]  P^T of every branch in fragment order
  const double* packS;   // [nn][C][kFamPackS]  stacked P | dP | d2P in fragment order
  const double *rootfreq, *probs, *SR, *weights;
  const int* rexp;
  double* part;          // [nn][2][part_stride]
  int part_stride;       // slots per (branch, sum) of `part` (>= the CTAs that work on one father)
};

// The B operands in FRAGMENT ORDER.  Column n = 8 nb + 2 q' + h of the stacked operand holds slot t = 2 nb + h of lane
// q': mat = t / 5, i = t % 5, state x = 4 i + q'.  The value lane (g, q) feeds to the DMMA of (kb, nb) is
// B[k = ymap(kb, q)][n = 8 nb + g]; the two column blocks nb = 2 pi, 2 pi + 1 of a lane are stored side by side,
//   pack[((kb * NP + pi) * 32 + lane) * 2 + h],   NP = 4 (stacked P | dP | d2P) or 2 (P^T),
// so one conflict-free 16-byte shared-memory load feeds two DMMAs.  Packed ONCE per evaluation by family_pack_kernel
// (block = (branch, class)) and staged with straight cp.async copies: gathering them from the row-major tables inside
// the likelihood kernel cost more than the contractions (serialised L2 latency, 34 dependent iterations per class).
constexpr int kFamPackA = kFamKB * 2 * 64;  // doubles per (branch, class)
constexpr int kFamPackS = kFamKB * 4 * 64;

// state owned by lane q in slot i (0..4)
__host__ __device__ __forceinline__ int fam_state(int i, int q) { return i < 4 ? 4 * q + i : 16 + q; }

struct FamilyPackParams {
  const double *P, *dP, *d2P;  // [nn][C][S][S]; d2P may be null
  double *packA, *packS;
  double* packL;                  // [nn][C][kFamPackA]  P alone (pruning pass); packA/packS are skipped when dP is null
  // tip tables [nl][C][ncodes][S] (dtt / d2tt may be null) -> packT [nl][C][ncodes][k = 4][q = 4][4]: the 16 slots of lane q
  // (t = 4k + e = 5 mat + i <-> state fam_state(i, q)) in four 32-byte pieces, piece k of the four lanes of a quad
  // contiguous, so that one 256-bit load per lane fetches a full 128-byte line per row
  const double *tt, *dtt, *d2tt;
  double* packT;
  int S, nbc, ntc;  // nbc = nn * C operand blocks, then ntc = nl * C * ncodes tip rows (0: none)
};

__global__ void family_pack_kernel(FamilyPackParams p) {
  const int S = p.S;
  if ((int)blockIdx.x >= p.nbc) {
    const size_t row = (size_t)(blockIdx.x - p.nbc) * (blockDim.x / 64) + threadIdx.x / 64;
    if (row >= (size_t)p.ntc) return;
    const int e = threadIdx.x & 63, k = e >> 4, q = (e >> 2) & 3, t = 4 * k + (e & 3), mat = t / 5, i = t - 5 * mat;
    const double* tab = mat == 0 ? p.tt : (mat == 1 ? p.dtt : (mat == 2 ? p.d2tt : nullptr));
    p.packT[row * 64 + e] = tab != nullptr ? tab[row * S + fam_state(i, q)] : 0.0;
    return;
  }
  const size_t SS = (size_t)S * S;
  const size_t bc = blockIdx.x;  // branch * C + class
  const double* m[3] = {p.P + bc * SS, p.dP ? p.dP + bc * SS : nullptr, p.d2P ? p.d2P + bc * SS : nullptr};
  const int total = p.dP ? 2 * kFamPackA + kFamPackS : kFamPackA;  // [packL][packA][packS]
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    const int which = e < kFamPackA ? 0 : (e < 2 * kFamPackA ? 1 : 2);  // L, A, S
    const int ee = which == 0 ? e : (which == 1 ? e - kFamPackA : e - 2 * kFamPackA);
    const int NP = which == 2 ? 4 : 2;
    const int h = ee & 1, lane = (ee >> 1) & 31, kp = ee >> 6, pi = kp % NP, kb = kp / NP;
    const int g = lane >> 2, q = lane & 3;
    const int y = dmma_ymap<kFamKB>(kb, q);  // k index
    const int nb = 2 * pi + h, qq = g >> 1, hh = g & 1, t = 2 * nb + hh;  // column n = 8 nb + g
    const int mat = t / 5, i = t - 5 * mat, x = fam_state(i, qq);
    double v = 0.0;
    if (x < S && y < S) {
      if (which == 2) {
        if (mat < 3 && m[mat]) v = m[mat][(size_t)x * S + y];
      } else if (mat == 0) {
        v = which == 1 ? m[0][(size_t)y * S + x] : m[0][(size_t)x * S + y];
      }
    }
    (which == 0 ? p.packL + bc * kFamPackA : (which == 1 ? p.packA + bc * kFamPackA : p.packS + bc * kFamPackS))[ee] = v;
  }
}

// n doubles (multiple of 2), both sides 16-byte aligned
template <int NT>
__device__ __forceinline__ void family_stage_async(double* dst, const double* src, int n) {
  for (int e = 2 * threadIdx.x; e < n; e += 2 * NT) cp_async16(dst + e, src + e);
}

// acc[nb] = A-fragments x packed operand (NP pairs per k block)
template <int NB, int NP>
__device__ __forceinline__ void family_contract(double (&acc)[NB][2], const double (&a)[kFamKB], const double* Ms, int lane) {
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) acc[nb][0] = acc[nb][1] = 0.0;
#pragma unroll
  for (int kb = 0; kb < kFamKB; ++kb)
#pragma unroll
    for (int pi = 0; pi < (NB + 1) / 2; ++pi) {
      const double2 b = *reinterpret_cast<const double2*>(Ms + ((kb * NP + pi) * 32 + lane) * 2);
      dmma884(acc[2 * pi][0], acc[2 * pi][1], a[kb], b.x);
      if (2 * pi + 1 < NB) dmma884(acc[2 * pi + 1][0], acc[2 * pi + 1][1], a[kb], b.y);
    }
}

// s * 2^k without the library scalbn: two exact power-of-two factors cover |k| <= 2000
__device__ __forceinline__ double family_scale2(double s, int k) {
  const int k1 = max(-1000, min(1000, k));
  s *= pow2(k1);
  const int k2 = max(-1000, min(1000, k - k1));
  return k2 ? s * pow2(k2) : s;
}

// Per-warp row staging: the CLV rows of the NEXT row block (upper[f] and every internal son: up to 4 x 8 rows x 160 B)
// travel HBM -> shared memory with cp.async while the current block is contracted, so no register is spent on
// prefetching and DRAM latency never sits in front of a DMMA.  Row stride 22 doubles: the 16-byte A-fragment reads of a
// quarter warp then hit 32 distinct banks.
constexpr int kFamRowStride = 22;
constexpr int kFamRowArr = 8 * kFamRowStride;  // doubles per staged array (8 rows)

// internal sons of a KIND (they alone need operands and staged rows; a small shared-memory footprint leaves the rest of the
// 256 KB to L1, which is what serves the tip-table gathers)
__host__ __device__ constexpr int fam_msi(int KIND) { return KIND == 4 ? 3 : 2 - (KIND & 1) - ((KIND >> 1) & 1); }
// Depth of the per-warp row ring: the fewer arrays an item stages, the shorter it computes and the further ahead its
// rows must be requested to keep HBM busy (binary father with two internal sons: 2 stages; one: 3; none: 6 -- while leaving L1 room for the tip tables).
__host__ __device__ constexpr int fam_stages(int KIND) { return KIND == 4 ? 2 : (KIND == 0 ? 2 : (KIND == 3 ? 6 : 3)); }
// one stage of one warp: the staged arrays, then 8 exponents per array
__host__ __device__ constexpr int fam_rowstage(int KIND) { return (1 + fam_msi(KIND)) * (kFamRowArr + 4); }
template <int KIND>
constexpr size_t dmma_family_smem(int C) {
  return (size_t)(C * (kFamPackA + fam_msi(KIND) * kFamPackS) +
                  (fam_threads_kind(KIND) / 32) * fam_stages(KIND) * fam_rowstage(KIND)) * sizeof(double);
}

// NB = 5 (d1) or 8 (d1, d2) column blocks per son.
// KIND 0..3: a binary father whose son j is a tip iff bit j is set (straight-line code, 12 warps);  KIND 4: one to three
// sons of any kind decided at run time (the unrooted root, unary nodes; 6 warps).
template <int NB, int KIND>
__device__ __forceinline__ void dmma_family_body(const DmmaFamilyParams& p, const int cta_index) {
  constexpr bool GEN = KIND == 4;
  constexpr int MS = GEN ? 3 : 2;
  constexpr int NMAT = NB == 5 ? 2 : 3;
  constexpr int NT = fam_threads_kind(KIND);
  constexpr int NW = NT / 32;
  constexpr int MSI = fam_msi(KIND);                     // sons that can be internal: operand / row slots
  constexpr int MATS = kFamPackA + MSI * kFamPackS;      // doubles per class
  constexpr int NST = fam_stages(KIND);                  // row ring depth
  constexpr int ROWSTAGE = fam_rowstage(KIND);           // doubles per row stage of one warp
  constexpr int EXPOFF = (1 + MSI) * kFamRowArr;         // the stage's exponents: int [1 + MSI][8]
  extern __shared__ __align__(16) double sm_fam[];       // [C][MATS] operands, then [warp][NST][ROWSTAGE] rows
  __shared__ double red[32];
  __shared__ double sprobs[kFamMaxClasses];
  if (threadIdx.x < p.C) sprobs[threadIdx.x] = p.probs[threadIdx.x];  // (visible after the prologue's barrier)

  auto has = [&](int j) { return GEN ? j < p.nson : true; };
  auto tip = [&](int j) { return GEN ? p.sons[j].kind == CHILD_TIP : ((KIND >> j) & 1) != 0; };
  auto slot = [&](int j) { return GEN ? j : (j == 1 && !(KIND & 1) ? 1 : 0); };  // operand / row slot of internal son j

  const int S = p.S, C = p.C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  double* rows = sm_fam + (size_t)C * MATS + (size_t)warp * NST * ROWSTAGE;
  // pattern indices below are relative to the CTA's first pattern (32-bit); the CTA's base offsets are folded in once
  const long long cta0 = (long long)cta_index * p.ppc;
  const int ncta = (int)(cta0 + p.ppc < p.N ? p.ppc : p.N - cta0);
  const bool root = p.father < 0;
  const double* fup = p.fup;
  const int* fupexp = p.fupexp;
  const int prow = p.prow, crow = p.crow;
  const int rowbase = (int)cta0 * prow;  // CLV row of (CTA pattern r, class c) = rowbase + r * prow + c * crow  (< 2^31 / S, host-checked)

  // per-lane constants of the row copy: piece e = lane + 32 it of the 8 x 10 16-byte pieces of an array
  int cp_r[3], cp_src[3], cp_dst[3];
#pragma unroll
  for (int it = 0; it < 3; ++it) {
    const int e = lane + 32 * it;
    cp_r[it] = e / 10;
    cp_src[it] = 2 * (e - cp_r[it] * 10);
    cp_dst[it] = cp_r[it] * kFamRowStride + cp_src[it];
  }
  // which exponent this lane copies: array lane / 8 (0 = upper[f], 1 + slot = internal son), row lane % 8
  const int* exp_src = nullptr;
  if (lane < 8) {
    if (!root) exp_src = fupexp;
  } else {
#pragma unroll
    for (int j = 0; j < MS; ++j)
      if (has(j) && !tip(j) && (lane >> 3) == 1 + slot(j)) exp_src = p.sons[j].exp;
  }

  // rows [r0, r0 + 8) (relative to the CTA) of class c and their exponents -> stage st of this warp (asynchronous)
  auto fetch_rows = [&](int r0, int c, int st) {
    double* dst = rows + (size_t)st * ROWSTAGE;
    const int rmax = ncta - 1 - r0;  // rows past the CTA's range re-read its last row
    const int row0 = rowbase + r0 * prow + c * crow;
#pragma unroll
    for (int it = 0; it < 3; ++it) {
      if (it < 2 || lane < 16) {
        const int off = (row0 + min(cp_r[it], rmax) * prow) * S + cp_src[it];
        double* d = dst + cp_dst[it];
        if (!root) cp_async16(d, fup + off);
#pragma unroll
        for (int j = 0; j < MS; ++j)
          if (has(j) && !tip(j)) cp_async16(d + (1 + slot(j)) * kFamRowArr, p.sons[j].clv + off);
      }
    }
    if (exp_src != nullptr) cp_async4(reinterpret_cast<int*>(dst + EXPOFF) + lane, exp_src + row0 + min(lane & 7, rmax) * prow);
  };
  // the fetch cursor runs NST - 1 items ahead of the compute cursor through the same (row block, class) sequence
  int fr0 = warp * 8, fc = 0;
  auto fetch_next = [&](int st) {
    if (fr0 < ncta) {
      fetch_rows(fr0, fc, st);
      if (++fc == C) {
        fc = 0;
        fr0 += NW * 8;
      }
    }
    cp_async_commit();
  };
  // the 5 A fragments of row g from a staged array: a[0..3] = row[4q .. 4q+3], a[4] = row[16 + q]
  auto load_frag = [&](const double* arr, double (&a)[kFamKB]) {
    const double* row = arr + g * kFamRowStride;
    const double2 v0 = *reinterpret_cast<const double2*>(row + 4 * q);
    const double2 v1 = *reinterpret_cast<const double2*>(row + 4 * q + 2);
    a[0] = v0.x; a[1] = v0.y; a[2] = v1.x; a[3] = v1.y;
    a[4] = row[16 + q];
  };
  // per-pattern scalars of a row block, requested one row block (C items) ahead
  struct PatScal {
    double sr, w;
    int er, code[MS];
  };
  auto load_pat = [&](int r0, PatScal& ps) {
    const long long pat = cta0 + min(r0 + g, ncta - 1);
    ps.sr = p.SR[pat];
    ps.w = p.weights[pat];
    ps.er = p.rexp[pat];
#pragma unroll
    for (int j = 0; j < MS; ++j) ps.code[j] = (has(j) && tip(j)) ? load_code(p.sons[j].codes, p.code_bytes, pat) : 0;
  };

  // ---- prologue: operands of every class (packed once per evaluation, before the first launch of the pass: independent of the
  // previous father's launch, so they are staged while that launch drains), then the first NST - 1 items of every warp ---------
  pdl_launch_dependents();
  for (int c = 0; c < C; ++c) {
    double* base = sm_fam + (size_t)c * MATS;
    if (!root) family_stage_async<NT>(base, p.packA + ((size_t)p.father * C + c) * kFamPackA, kFamPackA);
#pragma unroll
    for (int j = 0; j < MS; ++j)
      if (has(j) && !tip(j))
        family_stage_async<NT>(base + kFamPackA + (size_t)slot(j) * kFamPackS, p.packS + ((size_t)p.sons[j].node * C + c) * kFamPackS,
                               kFamPackS);
  }
  cp_async_commit();
  pdl_wait();   // upper[f] is the previous launches' output
#pragma unroll
  for (int s0 = 0; s0 < NST - 1; ++s0) fetch_next(s0);
  PatScal pcur, pnxt;
  load_pat(warp * 8, pnxt);
  cp_async_wait<0>();
  __syncthreads();

  double acc1[MS], acc2[MS];  // per-thread partial sums of w dL and w (d2L - dL^2)   (lanes q = 0 only)
#pragma unroll
  for (int j = 0; j < MS; ++j) acc1[j] = acc2[j] = 0.0;
  int st = 0;

  for (int r0 = warp * 8; r0 < ncta; r0 += NW * 8) {
    const bool valid = r0 + g < ncta;
    pcur = pnxt;
    if (r0 + NW * 8 < ncta) load_pat(r0 + NW * 8, pnxt);
    const double inv_sr = 1.0 / pcur.sr;
    double dsum[MS];  // lane q = 0: dL summed over classes, lane q = 1: d2L
#pragma unroll
    for (int j = 0; j < MS; ++j) dsum[j] = 0.0;

    for (int c = 0; c < C; ++c) {
      // ---- request the item NST - 1 ahead, wait for this one -----------------------------------------------------------
      __syncwarp();  // every lane is done reading the stage that is about to be overwritten
      fetch_next(st == 0 ? NST - 1 : st - 1);
      cp_async_wait<NST - 1>();
      __syncwarp();
      const double* rs = rows + (size_t)st * ROWSTAGE;
      const int* ex = reinterpret_cast<const int*>(rs + EXPOFF);
      st = st + 1 == NST ? 0 : st + 1;
      const double* matA = sm_fam + (size_t)c * MATS;
      const double* matS = matA + kFamPackA;

      double a[kFamKB];
      double R[MS][NB][2];
      // ---- M, dM, d2M of tip sons: table rows in lane order, requested before the contractions -------------------------
#pragma unroll
      for (int j = 0; j < MS; ++j) {
        if (has(j) && tip(j)) {
          const double* tp = p.sons[j].tpack + (size_t)(c * p.ncodes + pcur.code[j]) * 64 + 4 * q;
#pragma unroll
          for (int k = 0; k < (2 * NB + 3) / 4; ++k) {
            double v0, v1, v2, v3;
            ld256nc(tp + 16 * k, v0, v1, v2, v3);
            R[j][2 * k][0] = v0;
            R[j][2 * k][1] = v1;
            if (2 * k + 1 < NB) {
              R[j][2 * k + 1][0] = v2;
              R[j][2 * k + 1][1] = v3;
            }
          }
        }
      }
      // ---- A_f -----------------------------------------------------------------------------------------------------
      double A[3][2];
      int eU = 0, eL[MS];
      if (!root) {
        load_frag(rs, a);
        eU = ex[g];
        family_contract<3, 2>(A, a, matA, lane);
      } else {
#pragma unroll
        for (int t = 0; t < 6; ++t) A[t >> 1][t & 1] = t < 5 ? p.rootfreq[fam_state(t, q)] : 0.0;
      }
      // ---- M, dM, d2M of internal sons ---------------------------------------------------------------------------------
#pragma unroll
      for (int j = 0; j < MS; ++j) {
        eL[j] = 0;
        if (has(j) && !tip(j)) {
          load_frag(rs + (1 + slot(j)) * kFamRowArr, a);
          eL[j] = ex[(1 + slot(j)) * 8 + g];
          family_contract<NB, 4>(R[j], a, matS + (size_t)slot(j) * kFamPackS, lane);
        }
      }
      // ---- per son: upper row, rescale, store, derivative dots --------------------------------------------------------
      const double rinv = sprobs[c] * inv_sr;
      const int oexp = rowbase + (r0 + g) * prow + c * crow;  // output row (only dereferenced when valid)
      const int ooff = oexp * S;
#pragma unroll
      for (int j = 0; j < MS; ++j) {
        if (has(j)) {
          double U[5];
          int Eu = eU;
#pragma unroll
          for (int i = 0; i < 5; ++i) U[i] = A[i >> 1][i & 1];
#pragma unroll
          for (int o = 0; o < MS; ++o) {
            if (o != j && has(o)) {
              Eu += eL[o];
#pragma unroll
              for (int i = 0; i < 5; ++i) U[i] *= R[o][i >> 1][i & 1];
            }
          }
          int m = 0;
#pragma unroll
          for (int i = 0; i < 5; ++i) m = max(m, hi_word(U[i]));
          m = max(m, __shfl_xor_sync(0xffffffffu, m, 1));
          m = max(m, __shfl_xor_sync(0xffffffffu, m, 2));
          {  // branch-free (the two sons' shuffle chains can then overlap): k = 0 multiplies by 1
            const int k = (m < kScaleThresholdHi && m >= (1 << 20)) ? rescale_shift(m) : 0;
            const double f = pow2(k);
#pragma unroll
            for (int i = 0; i < 5; ++i) U[i] *= f;
            Eu += k;
          }
          if (p.sons[j].up != nullptr && valid) {
            double* row = p.sons[j].up + ooff;
            st256(row + 4 * q, U[0], U[1], U[2], U[3]);
            row[16 + q] = U[4];
            if (q == 0) p.sons[j].upexp[oexp] = Eu;
          }
          double s1 = 0.0, s2 = 0.0;
#pragma unroll
          for (int i = 0; i < 5; ++i) {
            const double u = (p.nh_form && R[j][i >> 1][i & 1] == 0.0) ? 0.0 : U[i];
            s1 = fma(u, R[j][(5 + i) >> 1][(5 + i) & 1], s1);
            if (NMAT > 2) s2 = fma(u, R[j][(10 + i) >> 1][(10 + i) & 1], s2);
          }
          // lanes q = 0 / 1 of a row finish d1 / d2: quad sums, exponent alignment, class weight over site likelihood
          double sv = (q & 1) ? s2 : s1;
          const double so = (q & 1) ? s1 : s2;
          sv += __shfl_xor_sync(0xffffffffu, so, 1);  // lane q^1 sends what this lane keeps
          sv += __shfl_xor_sync(0xffffffffu, sv, 2);
          dsum[j] += family_scale2(sv, pcur.er - Eu - eL[j]) * rinv;
        }
      }
    }  // classes
    // pattern level (DRHomogeneousTreeLikelihood.cpp:340-368, :425-454): lane q = 0 takes d2L from lane q = 1
#pragma unroll
    for (int j = 0; j < MS; ++j) {
      if (has(j)) {
        const double d2 = __shfl_xor_sync(0xffffffffu, dsum[j], 1);
        if (q == 0 && valid) {
          acc1[j] += pcur.w * dsum[j];
          acc2[j] += pcur.w * ((NMAT > 2 ? d2 : 0.0) - dsum[j] * dsum[j]);
        }
      }
    }
  }  // row blocks

#pragma unroll
  for (int j = 0; j < MS; ++j) {
    if (has(j)) {
      const double b1 = block_sum(acc1[j], red);
      const double b2 = block_sum(acc2[j], red);
      if (threadIdx.x == 0) {
        p.part[((size_t)p.sons[j].node * 2 + 0) * p.part_stride + cta_index] = b1;
        p.part[((size_t)p.sons[j].node * 2 + 1) * p.part_stride + cta_index] = b2;
      }
    }
  }
}

// one father per launch
template <int NB, int KIND>
__global__ void __launch_bounds__(fam_threads_kind(KIND), 1) dmma_family_kernel(DmmaFamilyParams p) {
  dmma_family_body<NB, KIND>(p, (int)blockIdx.x);
}
// every father of one DEPTH (their upper arrays are complete: written by the launches of the depth above) with the same kind of
// sons in one launch, as dmma_prune_level_kernel does for the pruning pass
template <int NB, int KIND>
__global__ void __launch_bounds__(fam_threads_kind(KIND), 1) dmma_family_level_kernel(const DmmaFamilyParams* __restrict__ fathers,
                                                                                       int ctas_per_father) {
  // the descriptor in shared memory: its fields are then read like kernel parameters (no registers held for them)
  __shared__ DmmaFamilyParams sp;
  const int f = (int)blockIdx.x / ctas_per_father;
  if (threadIdx.x < sizeof(DmmaFamilyParams) / 4)
    reinterpret_cast<int*>(&sp)[threadIdx.x] = reinterpret_cast<const int*>(fathers + f)[threadIdx.x];
  __syncthreads();
  dmma_family_body<NB, KIND>(sp, (int)blockIdx.x - f * ctas_per_father);
}

// block = branch: sums the per-CTA partials of the family kernel into out[1 + n] and out[1 + nn + n]
__global__ void finalize_family_kernel(const double* part, const int* mask, int G, int nn, double* out, unsigned want) {
  __shared__ double red[32];
  const int n = blockIdx.x;
  if (!mask[n]) return;
  double a1 = 0.0, a2 = 0.0;
  for (int i = threadIdx.x; i < G; i += blockDim.x) {
    a1 += part[((size_t)n * 2 + 0) * G + i];
    a2 += part[((size_t)n * 2 + 1) * G + i];
  }
  const double s1 = block_sum(a1, red);
  const double s2 = block_sum(a2, red);
  if (threadIdx.x == 0) {
    out[1 + n] = s1;
    if (want & 4u) out[1 + nn + n] = s2;
  }
}


// ----------------------------------------------------------------------------------------------------------------------
// K2 with the same machinery, for S = 20 (protein) and S = 64 (codon): one pruning step
// CLV_n = prod_sons (P_son . CLV_son)  per launch (RHomogeneousTreeLikelihood::computeSubtreeLikelihood,
// Likelihood/RHomogeneousTreeLikelihood.cpp:802-863; DRHomogeneousTreeLikelihood::computeLikelihoodFromArrays :819-864),
// per-row power-of-two rescaling fused.  Lane q of a quad owns the S/4 states dmma_ymap(t, q), t = 0 .. S/4 - 1, in its
// accumulator columns -- the states whose A fragments it loads -- so rows are stored 32 bytes per lane and tip sons read
// their table rows the same way.  S = 20: an item is light (15 DMMAs per internal son), the kernel lives on HBM: 16 warps
// per SM with a 3-deep cp.async ring.  S = 64: 128 DMMAs per internal son and item, FP64-tensor bound: 8 warps, 2-deep ring,
// the 32 KB operand of each son resident.
struct PruneSon {
  int kind, node;
  const double* clv;   // internal son: lower CLV slab, exponents
  const int* exp;
  const void* codes;   // tip son: codes, this leaf's tip table [C][ncodes][S]
  const double* tt;
};
struct DmmaPruneParams {
  PruneSon sons[kFamMaxSons];
  int nson;
  int S, C, ncodes, code_bytes;
  int ppc;
  int prow, crow;       // CLV row of (pattern i, class c) = i * prow + c * crow
  long long N;
  const double* packL;  // [nn][C][prune_pack(S)]
  double* out;          // CLV slab of the node
  int* out_exp;
};

__host__ __device__ constexpr int prune_kb(int S) { return S / 4; }                       // k blocks = states per lane
__host__ __device__ constexpr int prune_nb(int S) { return (S + 7) / 8; }                 // column blocks
__host__ __device__ constexpr int prune_np(int S) { return (prune_nb(S) + 1) / 2; }       // 16-byte operand pairs per k block
__host__ __device__ constexpr int prune_pack(int S) { return prune_kb(S) * prune_np(S) * 64; }  // operand doubles per (branch, class)
__host__ __device__ constexpr int prune_rowstride(int S) { return S + 2; }                // = 2 (mod 4): conflict-free 16-byte fragment reads
__host__ __device__ constexpr int prune_rowarr(int S) { return 8 * prune_rowstride(S); }

// P of every (branch, class) in fragment order for the pruning kernels (block = branch * C + class):
//   pack[((kb * NP + pi) * 32 + lane) * 2 + h] = P[x][y],  y = ymap(kb, q), column n = 8 (2 pi + h) + g  <->  slot t = 2 nb + (g & 1)
//   of lane q' = g >> 1, x = ymap(t, q')
template <int S_>
__global__ void prune_pack_kernel(const double* P, double* packL) {
  constexpr int KB = prune_kb(S_), NP = prune_np(S_), PACK = prune_pack(S_);
  const size_t bc = blockIdx.x;
  const double* m = P + bc * S_ * S_;
  for (int e = threadIdx.x; e < PACK; e += blockDim.x) {
    const int h = e & 1, lane = (e >> 1) & 31, kp = e >> 6, pi = kp % NP, kb = kp / NP;
    const int g = lane >> 2, q = lane & 3;
    const int y = dmma_ymap<KB>(kb, q);
    const int nb = 2 * pi + h, t = 2 * nb + (g & 1);
    double v = 0.0;
    if (t < KB) {
      const int x = dmma_ymap<KB>(t, g >> 1);
      if (x < S_ && y < S_) v = m[(size_t)x * S_ + y];
    }
    packL[bc * PACK + e] = v;
  }
}

// acc[nb] = A-fragments x packed operand, KB k blocks
template <int KB, int NB, int NP>
__device__ __forceinline__ void prune_contract(double (&acc)[NB][2], const double (&a)[KB], const double* Ms, int lane) {
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) acc[nb][0] = acc[nb][1] = 0.0;
#pragma unroll
  for (int kb = 0; kb < KB; ++kb)
#pragma unroll
    for (int pi = 0; pi < (NB + 1) / 2; ++pi) {
      const double2 b = *reinterpret_cast<const double2*>(Ms + ((kb * NP + pi) * 32 + lane) * 2);
      dmma884(acc[2 * pi][0], acc[2 * pi][1], a[kb], b.x);
      if (2 * pi + 1 < NB) dmma884(acc[2 * pi + 1][0], acc[2 * pi + 1][1], a[kb], b.y);
    }
}

// the S/4 values of a row owned by lane q: v[t] <-> state ymap(t, q); 32-byte accesses for the full 16-state chunks
template <int S_, bool NC>
__device__ __forceinline__ void prune_load_states(const double* row, int q, double (&v)[S_ / 4]) {
  constexpr int KB = S_ / 4, J = S_ / 16;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    if (NC) ld256nc(row + 16 * j + 4 * q, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    else ld256(row + 16 * j + 4 * q, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  }
#pragma unroll
  for (int t = 4 * J; t < KB; ++t) v[t] = NC ? __ldg(row + 16 * J + 4 * (t - 4 * J) + q) : row[16 * J + 4 * (t - 4 * J) + q];
}

// CFG picks warps x ring depth of the binary S = 20 kinds: 0 = 16 x 3 (default), 1 = 12 x 4, 2 = 8 x 6 (BPPGPU_PRUNE_CFG, A/B
// runs); the run-time kind (three sons: 50 % more shared memory per warp) is 8 x 3; S = 64 is always 8 x 2 (run-time kind 4 x 2)
__host__ __device__ constexpr int prune_threads(int S, int KIND, int CFG) {
  return S > 32 ? (KIND == 4 ? 128 : 256) : (KIND == 4 ? 256 : (CFG == 0 ? 512 : (CFG == 1 ? 384 : 256)));
}
__host__ __device__ constexpr int prune_stages(int S, int KIND, int CFG) {
  return S > 32 ? 2 : (KIND == 4 ? 3 : (CFG == 0 ? 3 : (CFG == 1 ? 4 : 6)));
}
__host__ __device__ constexpr int prune_rowstage(int S, int KIND) { return fam_msi(KIND) * (prune_rowarr(S) + 4); }
template <int S_, int KIND, int CFG>
constexpr size_t dmma_prune_smem(int C) {
  return (size_t)(C * fam_msi(KIND) * prune_pack(S_) +
                  (prune_threads(S_, KIND, CFG) / 32) * prune_stages(S_, KIND, CFG) * prune_rowstage(S_, KIND)) * sizeof(double);
}

template <int S_, int KIND, int CFG>
__device__ __forceinline__ void dmma_prune_body(const DmmaPruneParams& p, const int cta_index) {
  constexpr bool GEN = KIND == 4;
  constexpr int MS = GEN ? 3 : 2;
  constexpr int KB = prune_kb(S_), NB = prune_nb(S_), NP = prune_np(S_), PACK = prune_pack(S_);
  constexpr int RSTRIDE = prune_rowstride(S_), RARR = prune_rowarr(S_);
  constexpr int NT = prune_threads(S_, KIND, CFG);
  constexpr int NW = NT / 32;
  constexpr int MSI = fam_msi(KIND);
  constexpr int MATS = MSI * PACK;
  constexpr int NST = prune_stages(S_, KIND, CFG);
  constexpr int ROWSTAGE = prune_rowstage(S_, KIND);
  constexpr int EXPOFF = MSI * RARR;
  constexpr int PIECES = S_ / 2;                    // 16-byte pieces of a row
  constexpr int NIT = (8 * PIECES + 31) / 32;       // copy instructions per lane and array
  extern __shared__ __align__(16) double sm_pr[];   // [C][MATS] operands, then [warp][NST][ROWSTAGE]

  auto has = [&](int j) { return GEN ? j < p.nson : true; };
  auto tip = [&](int j) { return GEN ? p.sons[j].kind == CHILD_TIP : ((KIND >> j) & 1) != 0; };
  auto slot = [&](int j) { return GEN ? j : (j == 1 && !(KIND & 1) ? 1 : 0); };

  const int C = p.C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  double* rows = sm_pr + (size_t)C * MATS + (size_t)warp * NST * ROWSTAGE;
  const long long cta0 = (long long)cta_index * p.ppc;
  const int ncta = (int)(cta0 + p.ppc < p.N ? p.ppc : p.N - cta0);
  const int prow = p.prow, crow = p.crow;
  const int rowbase = (int)cta0 * prow;

  int cp_r[NIT], cp_src[NIT], cp_dst[NIT];
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int e = lane + 32 * it;
    cp_r[it] = e / PIECES;
    cp_src[it] = 2 * (e - cp_r[it] * PIECES);
    cp_dst[it] = cp_r[it] * RSTRIDE + cp_src[it];
  }
  const int* exp_src = nullptr;
#pragma unroll
  for (int j = 0; j < MS; ++j)
    if (has(j) && !tip(j) && (lane >> 3) == slot(j)) exp_src = p.sons[j].exp;

  auto fetch_rows = [&](int r0, int c, int st) {
    double* dst = rows + (size_t)st * ROWSTAGE;
    const int rmax = ncta - 1 - r0;
    const int row0 = rowbase + r0 * prow + c * crow;
    if (MSI > 0) {
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        if (lane + 32 * it < 8 * PIECES) {
          const int off = (row0 + min(cp_r[it], rmax) * prow) * S_ + cp_src[it];
          double* d = dst + cp_dst[it];
#pragma unroll
          for (int j = 0; j < MS; ++j)
            if (has(j) && !tip(j)) cp_async16(d + slot(j) * RARR, p.sons[j].clv + off);
        }
      }
      if (exp_src != nullptr) cp_async4(reinterpret_cast<int*>(dst + EXPOFF) + lane, exp_src + row0 + min(lane & 7, rmax) * prow);
    }
  };
  int fr0 = warp * 8, fc = 0;
  auto fetch_next = [&](int st) {
    if (fr0 < ncta) {
      fetch_rows(fr0, fc, st);
      if (++fc == C) {
        fc = 0;
        fr0 += NW * 8;
      }
    }
    cp_async_commit();
  };
  // the A fragments of row g from a staged array: a[kb] = row[ymap(kb, q)]
  auto load_frag = [&](const double* arr, double (&a)[KB]) {
    const double* row = arr + g * RSTRIDE;
    constexpr int J = S_ / 16;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const double2 v0 = *reinterpret_cast<const double2*>(row + 16 * j + 4 * q);
      const double2 v1 = *reinterpret_cast<const double2*>(row + 16 * j + 4 * q + 2);
      a[4 * j] = v0.x; a[4 * j + 1] = v0.y; a[4 * j + 2] = v1.x; a[4 * j + 3] = v1.y;
    }
#pragma unroll
    for (int kb = 4 * J; kb < KB; ++kb) a[kb] = row[16 * J + 4 * (kb - 4 * J) + q];
  };
  auto load_codes = [&](int r0, int (&code)[MS]) {
    const long long pat = cta0 + min(r0 + g, ncta - 1);
#pragma unroll
    for (int j = 0; j < MS; ++j) code[j] = (has(j) && tip(j)) ? load_code(p.sons[j].codes, p.code_bytes, pat) : 0;
  };

  // operands first (independent of the previous node's launch: staged while it drains), then the sons' rows
  pdl_launch_dependents();
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int j = 0; j < MS; ++j)
      if (has(j) && !tip(j))
        family_stage_async<NT>(sm_pr + (size_t)c * MATS + (size_t)slot(j) * PACK, p.packL + ((size_t)p.sons[j].node * C + c) * PACK, PACK);
  cp_async_commit();
  pdl_wait();   // the sons' CLVs are earlier launches' output
#pragma unroll
  for (int s0 = 0; s0 < NST - 1; ++s0) fetch_next(s0);
  int ccur[MS], cnxt[MS];
  load_codes(warp * 8, cnxt);
  cp_async_wait<0>();
  __syncthreads();

  int st = 0;
  for (int r0 = warp * 8; r0 < ncta; r0 += NW * 8) {
    const bool valid = r0 + g < ncta;
#pragma unroll
    for (int j = 0; j < MS; ++j) ccur[j] = cnxt[j];
    if (r0 + NW * 8 < ncta) load_codes(r0 + NW * 8, cnxt);
    for (int c = 0; c < C; ++c) {
      __syncwarp();
      fetch_next(st == 0 ? NST - 1 : st - 1);
      cp_async_wait<NST - 1>();
      __syncwarp();
      const double* rs = rows + (size_t)st * ROWSTAGE;
      const int* ex = reinterpret_cast<const int*>(rs + EXPOFF);
      st = st + 1 == NST ? 0 : st + 1;
      const double* mats = sm_pr + (size_t)c * MATS;

      double prod[KB];
      int Ea = 0;
      bool firstson = true;
#pragma unroll
      for (int j = 0; j < MS; ++j) {
        if (has(j)) {
          if (tip(j)) {
            double v[KB];
            prune_load_states<S_, true>(p.sons[j].tt + (size_t)(c * p.ncodes + ccur[j]) * S_, q, v);
#pragma unroll
            for (int i = 0; i < KB; ++i) prod[i] = firstson ? v[i] : prod[i] * v[i];
          } else {
            double a[KB], acc[NB][2];
            load_frag(rs + slot(j) * RARR, a);
            Ea += ex[slot(j) * 8 + g];
            prune_contract<KB, NB, NP>(acc, a, mats + (size_t)slot(j) * PACK, lane);
#pragma unroll
            for (int i = 0; i < KB; ++i) prod[i] = firstson ? acc[i >> 1][i & 1] : prod[i] * acc[i >> 1][i & 1];
          }
          firstson = false;
        }
      }
      int m = 0;
#pragma unroll
      for (int i = 0; i < KB; ++i) m = max(m, hi_word(prod[i]));
      m = max(m, __shfl_xor_sync(0xffffffffu, m, 1));
      m = max(m, __shfl_xor_sync(0xffffffffu, m, 2));
      const int k = (m < kScaleThresholdHi && m >= (1 << 20)) ? rescale_shift(m) : 0;
      const double f = pow2(k);
      Ea += k;
      if (valid) {
        const int orow = rowbase + (r0 + g) * prow + c * crow;
        double* row = p.out + (size_t)orow * S_;
        constexpr int J = S_ / 16;
#pragma unroll
        for (int j = 0; j < J; ++j)
          st256(row + 16 * j + 4 * q, prod[4 * j] * f, prod[4 * j + 1] * f, prod[4 * j + 2] * f, prod[4 * j + 3] * f);
#pragma unroll
        for (int t = 4 * J; t < KB; ++t) row[16 * J + 4 * (t - 4 * J) + q] = prod[t] * f;
        if (q == 0) p.out_exp[orow] = Ea;
      }
    }
  }
}

// one node per launch: the CTAs share its pattern range
template <int S_, int KIND, int CFG>
__global__ void __launch_bounds__(prune_threads(S_, KIND, CFG), 1) dmma_prune_kernel(DmmaPruneParams p) {
  dmma_prune_body<S_, KIND, CFG>(p, (int)blockIdx.x);
}

// All nodes of one tree LEVEL (equal subtree height: independent of each other) with the same kind of sons in ONE launch:
// CTA b works on node b / ctas_per_node, pattern range b % ctas_per_node.  A 500-taxon tree has ~500 internal nodes on a few dozen
// levels; a launch per node pays its operand staging, pipeline fill and drain once per node on all 148 SMs, a launch per
// level once per level (a CTA keeps one node's operands for 148 / n times more rows), and the cherries -- a third of the nodes --
// are one launch.
template <int S_, int KIND, int CFG>
__global__ void __launch_bounds__(prune_threads(S_, KIND, CFG), 1) dmma_prune_level_kernel(const DmmaPruneParams* __restrict__ nodes,
                                                                                            int ctas_per_node) {
  const int node = (int)blockIdx.x / ctas_per_node;
  const DmmaPruneParams p = nodes[node];   // (a copy in registers measured faster than a copy in shared memory: 6.80 vs 6.97 ms on cfg4)
  dmma_prune_body<S_, KIND, CFG>(p, (int)blockIdx.x - node * ctas_per_node);
}

static_assert(sizeof(DmmaFamilyParams) % 4 == 0 && sizeof(DmmaFamilyParams) / 4 <= 128, "descriptor copy: one int per thread");
static_assert(sizeof(DmmaPruneParams) % 4 == 0 && sizeof(DmmaPruneParams) / 4 <= 128, "descriptor copy: one int per thread");

}  // namespace bppgpu
