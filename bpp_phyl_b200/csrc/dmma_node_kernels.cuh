// K2 on the FP64 tensor cores: one pruning step  CLV_node = prod_sons (CLV_son . P_son^T)  for medium / large
// state counts (protein S = 20, codon S = 64), CLVs resident in HBM in the reference's [pattern][class][state]
// order (RHomogeneousTreeLikelihood::computeSubtreeLikelihood, Likelihood/RHomogeneousTreeLikelihood.cpp:802-863;
// DR twin computeLikelihoodFromArrays, DRHomogeneousTreeLikelihood.cpp:819-864).
//
// GEMM view per son: T[row][x] = sum_y L[row][y] * B[y][x] with B[y][x] = P[x][y], i.e. the row-major pxy_ table
// IS the column-major B operand of mma.sync.m8n8k4.  A CTA handles one rate class and a tile of patterns; its
// 4 warps own RW 8-pattern row blocks each and all S columns.  The son's CLV rows go global -> registers as A
// fragments (each element is read exactly once, full 32-byte sectors), P_son of this class is staged once per
// CTA in shared memory with a row stride = 4 (mod 16) doubles so B-fragment loads are conflict free; tip sons are
// a gather from the per-branch tip table.  The Hadamard product over sons, the per-row power-of-two rescale
// (max over the row through two quad shuffles) and the exponent bookkeeping are fused into the epilogue.
#pragma once
#include "dmma.cuh"
#include "walk_kernels.cuh"

namespace bppgpu {

struct DmmaNodeParams {
  const Child* childs;  // sons of this node (kinds TIP / KEEP)
  int nchild;
  int out_idx;          // keep slab of the node
  int S, C, ncodes, code_bytes;
  long long N;
  const double* P;       // [nn][C][S][S]
  const double* tiptab;  // [nl][C][ncodes][S]
  const void* codes;     // [nl][N]
  double* keep;          // [ni][N][C][S]
  int* keep_exp;         // [ni][N][C]
};

constexpr int kDmmaNodeWarps = 4;

template <int KB>
__host__ __device__ constexpr int dmma_pstride() {
  // smallest stride >= 4*KB that is = 4 (mod 16) doubles
  return ((4 * KB + 11) / 16) * 16 + 4;
}
template <int KB, int NBLK>
constexpr size_t dmma_node_smem_per_child() {
  return (size_t)NBLK * 8 * dmma_pstride<KB>() * sizeof(double);
}

// The MMA sums over k, so WHICH state y sits at k-position (kb, q) is free as long as A and B agree.  The choice
// below lets a lane fetch its A elements with 32-byte vector loads (the 4 lanes of a row then read one contiguous
// 128-byte line): for the 16-state chunks j = 0 .. S/16-1, k-block kb = 4j+e of lane q holds y = 16j + 4q + e;
// the remaining S mod 16 states are taken 4 at a time, y = 16*(S/16) + 4*kb' + q.
template <int KB>
__host__ __device__ __forceinline__ int dmma_ymap(int kb, int q) {
  constexpr int J = KB / 4;  // full 16-state chunks
  return kb < 4 * J ? 16 * (kb >> 2) + 4 * q + (kb & 3) : 16 * J + 4 * (kb - 4 * J) + q;
}

// the A fragments of one CLV row (S doubles, 32-byte aligned): a[kb] = row[ymap(kb, q)]
template <int KB>
__device__ __forceinline__ void load_a_row(const double* row, int S, int q, double (&a)[KB]) {
  constexpr int J = KB / 4;
#pragma unroll
  for (int j = 0; j < J; ++j) ld256(row + 16 * j + 4 * q, a[4 * j], a[4 * j + 1], a[4 * j + 2], a[4 * j + 3]);
#pragma unroll
  for (int kb = 4 * J; kb < KB; ++kb) {
    const int y = 16 * J + 4 * (kb - 4 * J) + q;
    a[kb] = y < S ? row[y] : 0.0;
  }
}

// stage P[c] of one branch (row-major S x S) into shared memory as [x][kb*4+q] = P[x][ymap(kb,q)] (or the transpose),
// zero padded to [NBLK*8][stride]
template <int KB, int NBLK, bool TRANSPOSE>
__device__ __forceinline__ void stage_matrix(double* dst, const double* Pg, int S, int nthreads) {
  constexpr int SB = dmma_pstride<KB>();
  for (int e = threadIdx.x; e < NBLK * 8 * 4 * KB; e += nthreads) {
    const int x = e / (4 * KB), col = e - x * (4 * KB);
    const int y = dmma_ymap<KB>(col >> 2, col & 3);
    double v = 0.0;
    if (x < S && y < S) v = TRANSPOSE ? Pg[(size_t)y * S + x] : Pg[(size_t)x * S + y];
    dst[x * SB + col] = v;
  }
}

// acc[r][nb] += A(rows of block r) . B(staged matrix)   for the RW row blocks of this warp
template <int KB, int NBLK, int RW>
__device__ __forceinline__ void dmma_rows_times_matrix(double (&acc)[RW][NBLK][2], const double (&a)[RW][KB],
                                                       const double* Ps, int g, int q) {
  constexpr int SB = dmma_pstride<KB>();
#pragma unroll
  for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
    for (int nb = 0; nb < NBLK; ++nb) {
      const double b = Ps[(nb * 8 + g) * SB + kb * 4 + q];
#pragma unroll
      for (int r = 0; r < RW; ++r) dmma884(acc[r][nb][0], acc[r][nb][1], a[r][kb], b);
    }
  }
}

// two doubles of a row at an even column: one 16-byte access when the row is 16-byte aligned there (S even)
__device__ __forceinline__ void ld2(const double* row, int x, int S, bool vec, double& v0, double& v1) {
  if (vec && x + 1 < S) {
    const double2 t = *reinterpret_cast<const double2*>(row + x);
    v0 = t.x; v1 = t.y;
  } else {
    v0 = x < S ? row[x] : 0.0;
    v1 = x + 1 < S ? row[x + 1] : 0.0;
  }
}
__device__ __forceinline__ void st2(double* row, int x, int S, bool vec, double v0, double v1) {
  if (vec && x + 1 < S) {
    *reinterpret_cast<double2*>(row + x) = make_double2(v0, v1);
  } else {
    if (x < S) row[x] = v0;
    if (x + 1 < S) row[x + 1] = v1;
  }
}

constexpr int kNodePrefetch = 2;  // sons whose inputs are fetched one tile ahead (a third son, the root's, is fetched in place)

// Persistent CTA = (pattern tiles, class).  The inputs of tile t+1 (A fragments of the internal sons, tip codes,
// exponents) are requested before tile t is computed, and those of the first tile before the P matrices are staged, so
// HBM latency overlaps the tensor-core work instead of preceding it (ncu r1a: long-scoreboard on the first DMMA of every
// tile and on the staging stores dominated this kernel).
template <int KB, int NBLK, int RW>
__global__ void __launch_bounds__(kDmmaNodeWarps * 32) dmma_node_kernel(DmmaNodeParams p) {
  constexpr int NT = kDmmaNodeWarps * 32;
  constexpr int SB = dmma_pstride<KB>();
  extern __shared__ __align__(16) double sm_node[];
  const int S = p.S, C = p.C;
  const int c = blockIdx.y;
  const bool vec = (S & 1) == 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  const long long tile_rows = (long long)kDmmaNodeWarps * RW * 8;

  struct In {
    double a[kNodePrefetch][RW][KB];
    int e[kNodePrefetch][RW];
    int code[kNodePrefetch][RW];
  };
  auto pattern_of = [&](long long tile, int r) {
    const long long pp = (tile * kDmmaNodeWarps + warp) * RW * 8 + r * 8 + g;
    return pp < p.N ? pp : p.N - 1;
  };
  auto fetch = [&](long long tile, In& in) {
#pragma unroll
    for (int j = 0; j < kNodePrefetch; ++j) {
      if (j < p.nchild) {
        const Child ch = p.childs[j];
#pragma unroll
        for (int r = 0; r < RW; ++r) {
          const long long pat = pattern_of(tile, r);
          if (ch.kind == CHILD_TIP) {
            in.code[j][r] = load_code(p.codes, p.code_bytes, (long long)ch.idx * p.N + pat);
          } else {
            load_a_row<KB>(p.keep + (((size_t)ch.idx * p.N + pat) * C + c) * S, S, q, in.a[j][r]);
            in.e[j][r] = p.keep_exp[((size_t)ch.idx * p.N + pat) * C + c];
          }
        }
      }
    }
  };

  In nxt;
  long long tile = blockIdx.x;
  const bool any = tile * tile_rows < p.N;
  if (any) fetch(tile, nxt);

  int nint = 0;
  for (int j = 0; j < p.nchild; ++j) {
    const Child ch = p.childs[j];
    if (ch.kind == CHILD_TIP) continue;
    stage_matrix<KB, NBLK, false>(sm_node + (size_t)nint * NBLK * 8 * SB, p.P + ((size_t)ch.pnode * C + c) * S * S, S, NT);
    ++nint;
  }
  __syncthreads();

  for (; tile * tile_rows < p.N; tile += gridDim.x) {
    const long long pat_base = (tile * kDmmaNodeWarps + warp) * RW * 8;
    In cur = nxt;
    if ((tile + gridDim.x) * tile_rows < p.N) fetch(tile + gridDim.x, nxt);
    if (pat_base >= p.N) continue;  // this warp's rows are past the end (other warps of the CTA may still have rows)

    double prod[RW][NBLK][2];
    int Ea[RW];
#pragma unroll
    for (int r = 0; r < RW; ++r) Ea[r] = 0;
    nint = 0;
    // sons 0 .. kNodePrefetch-1: inputs already in registers (compile-time indices keep them there)
#pragma unroll
    for (int j = 0; j < kNodePrefetch; ++j) {
      if (j < p.nchild) {
        const Child ch = p.childs[j];
        double acc[RW][NBLK][2];
        if (ch.kind == CHILD_TIP) {
#pragma unroll
          for (int r = 0; r < RW; ++r) {
            const double* tt = p.tiptab + (((size_t)ch.idx * C + c) * p.ncodes + cur.code[j][r]) * S;
#pragma unroll
            for (int nb = 0; nb < NBLK; ++nb) ld2(tt, nb * 8 + 2 * q, S, vec, acc[r][nb][0], acc[r][nb][1]);
          }
        } else {
#pragma unroll
          for (int r = 0; r < RW; ++r) {
            Ea[r] += cur.e[j][r];
#pragma unroll
            for (int nb = 0; nb < NBLK; ++nb) acc[r][nb][0] = acc[r][nb][1] = 0.0;
          }
          dmma_rows_times_matrix<KB, NBLK, RW>(acc, cur.a[j], sm_node + (size_t)nint * NBLK * 8 * SB, g, q);
          ++nint;
        }
#pragma unroll
        for (int r = 0; r < RW; ++r)
#pragma unroll
          for (int nb = 0; nb < NBLK; ++nb) {
            prod[r][nb][0] = j == 0 ? acc[r][nb][0] : prod[r][nb][0] * acc[r][nb][0];
            prod[r][nb][1] = j == 0 ? acc[r][nb][1] : prod[r][nb][1] * acc[r][nb][1];
          }
      }
    }
    // further sons (multifurcations, the unrooted tree's third root son): fetched in place
    for (int j = kNodePrefetch; j < p.nchild; ++j) {
      const Child ch = p.childs[j];
      double acc[RW][NBLK][2];
      if (ch.kind == CHILD_TIP) {
#pragma unroll
        for (int r = 0; r < RW; ++r) {
          const int code = load_code(p.codes, p.code_bytes, (long long)ch.idx * p.N + pattern_of(tile, r));
          const double* tt = p.tiptab + (((size_t)ch.idx * C + c) * p.ncodes + code) * S;
#pragma unroll
          for (int nb = 0; nb < NBLK; ++nb) ld2(tt, nb * 8 + 2 * q, S, vec, acc[r][nb][0], acc[r][nb][1]);
        }
      } else {
        double a[RW][KB];
#pragma unroll
        for (int r = 0; r < RW; ++r) {
          const long long pat = pattern_of(tile, r);
          load_a_row<KB>(p.keep + (((size_t)ch.idx * p.N + pat) * C + c) * S, S, q, a[r]);
          Ea[r] += p.keep_exp[((size_t)ch.idx * p.N + pat) * C + c];
#pragma unroll
          for (int nb = 0; nb < NBLK; ++nb) acc[r][nb][0] = acc[r][nb][1] = 0.0;
        }
        dmma_rows_times_matrix<KB, NBLK, RW>(acc, a, sm_node + (size_t)nint * NBLK * 8 * SB, g, q);
        ++nint;
      }
#pragma unroll
      for (int r = 0; r < RW; ++r)
#pragma unroll
        for (int nb = 0; nb < NBLK; ++nb) {
          prod[r][nb][0] *= acc[r][nb][0];
          prod[r][nb][1] *= acc[r][nb][1];
        }
    }

#pragma unroll
    for (int r = 0; r < RW; ++r) {
      int m = 0;
#pragma unroll
      for (int nb = 0; nb < NBLK; ++nb) m = max(m, max(hi_word(prod[r][nb][0]), hi_word(prod[r][nb][1])));
      m = max(m, __shfl_xor_sync(0xffffffffu, m, 1));
      m = max(m, __shfl_xor_sync(0xffffffffu, m, 2));
      if (m < kScaleThresholdHi && m >= (1 << 20)) {
        const int k = rescale_shift(m);
        const double f = pow2(k);
#pragma unroll
        for (int nb = 0; nb < NBLK; ++nb) {
          prod[r][nb][0] *= f;
          prod[r][nb][1] *= f;
        }
        Ea[r] += k;
      }
      if (pat_base + r * 8 + g < p.N) {
        const long long pat = pat_base + r * 8 + g;
        double* row = p.keep + (((size_t)p.out_idx * p.N + pat) * C + c) * S;
#pragma unroll
        for (int nb = 0; nb < NBLK; ++nb) st2(row, nb * 8 + 2 * q, S, vec, prod[r][nb][0], prod[r][nb][1]);
        if (q == 0) p.keep_exp[((size_t)p.out_idx * p.N + pat) * C + c] = Ea[r];
      }
    }
  }
}

}  // namespace bppgpu
