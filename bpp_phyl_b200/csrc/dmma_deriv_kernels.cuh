// K4 + K5 on the FP64 tensor cores, fused per branch: for a non-root node n with father f
//   upper[n][i][c][x] = [f == root ? pi_x : sum_y P_f[c][y][x] upper[f][i][c][y]] * prod_{siblings b} sum_y P_b[c][x][y] lower_b[i][c][y]
//       (DRHomogeneousTreeLikelihood::computeSubtreeLikelihoodPrefix, Likelihood/DRHomogeneousTreeLikelihood.cpp:543-649;
//        the father-branch contraction uses P transposed, :919-945)
//   dLc[i][c]  = p_c 2^(eR_i - eU_ic - eD_ic) sum_x upper[n][i][c][x] sum_y dpxy_n[c][x][y]  lower_n[i][c][y] / SR_i
//   d2Lc[i][c] = ... d2pxy_n ...                     (computeTreeDLikelihoodAtNode :287-326, D2 :373-411)
// and, with nh_form, terms whose P_n contraction is exactly 0 are dropped
// (DRNonHomogeneousTreeLikelihood.cpp:396-407: larray includes the son, numerator/denominator form).
// deriv_combine_kernel then sums the classes per pattern and forms  w_i dL_i  and  w_i (d2L_i - dL_i^2)
// (:340-368, :425-454).
//
// Same tiling as dmma_node_kernel: CTA = (pattern tile, class), 4 warps x RW 8-row blocks x all S columns, every
// matrix of the branch staged once per CTA in shared memory, CLV rows read once from HBM as A fragments.
#pragma once
#include "dmma_node_kernels.cuh"

namespace bppgpu {

struct DmmaUpperParams {
  const Child* sibs;  // siblings of n (kinds TIP / KEEP)
  int nsib;
  int father;         // node id of f, or -1 when f is the root
  int father_upper;   // slab of upper[f] (when father >= 0)
  int node;           // n
  int node_is_tip;
  int node_idx;       // leaf slot or keep slab of n
  int upper_out;      // slab that receives upper[n], or -1 (tips: nobody reads it)
  int S, C, ncodes, code_bytes;
  int nh_form;
  unsigned want;      // bit1 d1, bit2 d2
  long long N;
  const double *P, *dP, *d2P;                 // [nn][C][S][S]
  const double *tiptab, *dtiptab, *d2tiptab;  // [nl][C][ncodes][S]
  const void* codes;
  const double* keep;    // lower CLVs [ni][N][C][S]
  const int* keep_exp;
  double* upper;         // [n_upper_slabs][N][C][S]
  int* upper_exp;
  const double* rootfreq;
  const double* probs;
  const double* SR;
  const int* rexp;
  double* dLc;           // [N][C][2]
};

template <int KB, int NBLK, int RW>
__global__ void __launch_bounds__(kDmmaNodeWarps * 32) dmma_upper_deriv_kernel(DmmaUpperParams p) {
  constexpr int NT = kDmmaNodeWarps * 32;
  constexpr int SB = dmma_pstride<KB>();
  constexpr int MAT = NBLK * 8 * SB;  // doubles per staged matrix
  extern __shared__ __align__(16) double sm_up[];
  const int S = p.S, C = p.C;
  const int c = blockIdx.y;
  const size_t SS = (size_t)S * S;

  // ---- stage: [P_f^T] [P_sib ...] [dP_n] [d2P_n] [P_n] ---------------------------------------------------------
  int nm = 0;
  int m_father = -1, m_d1 = -1, m_d2 = -1, m_p0 = -1;
  if (p.father >= 0) {
    stage_matrix<KB, NBLK, true>(sm_up + (size_t)nm * MAT, p.P + ((size_t)p.father * C + c) * SS, S, NT);
    m_father = nm++;
  }
  const int m_sib0 = nm;
  for (int j = 0; j < p.nsib; ++j) {
    const Child ch = p.sibs[j];
    if (ch.kind == CHILD_TIP) continue;
    stage_matrix<KB, NBLK, false>(sm_up + (size_t)nm * MAT, p.P + ((size_t)ch.pnode * C + c) * SS, S, NT);
    ++nm;
  }
  if (!p.node_is_tip) {
    if (p.want & 2u) {
      stage_matrix<KB, NBLK, false>(sm_up + (size_t)nm * MAT, p.dP + ((size_t)p.node * C + c) * SS, S, NT);
      m_d1 = nm++;
    }
    if (p.want & 4u) {
      stage_matrix<KB, NBLK, false>(sm_up + (size_t)nm * MAT, p.d2P + ((size_t)p.node * C + c) * SS, S, NT);
      m_d2 = nm++;
    }
    if (p.nh_form) {
      stage_matrix<KB, NBLK, false>(sm_up + (size_t)nm * MAT, p.P + ((size_t)p.node * C + c) * SS, S, NT);
      m_p0 = nm++;
    }
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  const long long tile_rows = (long long)kDmmaNodeWarps * RW * 8;
  for (long long tile = blockIdx.x; tile * tile_rows < p.N; tile += gridDim.x) {
  const long long pat_base = (tile * kDmmaNodeWarps + warp) * RW * 8;
  if (pat_base >= p.N) break;
  long long pat[RW];
#pragma unroll
  for (int r = 0; r < RW; ++r) {
    const long long pp = pat_base + r * 8 + g;
    pat[r] = pp < p.N ? pp : p.N - 1;
  }

  // ---- upper[n] -------------------------------------------------------------------------------------------------
  double U[RW][NBLK][2];
  int Eu[RW];
  double acc[RW][NBLK][2];
  double a[RW][KB];
  if (p.father >= 0) {
#pragma unroll
    for (int r = 0; r < RW; ++r) {
      load_a_row<KB>(p.upper + (((size_t)p.father_upper * p.N + pat[r]) * C + c) * S, S, q, a[r]);
      Eu[r] = p.upper_exp[((size_t)p.father_upper * p.N + pat[r]) * C + c];
#pragma unroll
      for (int nb = 0; nb < NBLK; ++nb) U[r][nb][0] = U[r][nb][1] = 0.0;
    }
    dmma_rows_times_matrix<KB, NBLK, RW>(U, a, sm_up + (size_t)m_father * MAT, g, q);
  } else {
#pragma unroll
    for (int r = 0; r < RW; ++r) {
      Eu[r] = 0;
#pragma unroll
      for (int nb = 0; nb < NBLK; ++nb) {
        const int x = nb * 8 + 2 * q;
        U[r][nb][0] = x < S ? p.rootfreq[x] : 0.0;
        U[r][nb][1] = x + 1 < S ? p.rootfreq[x + 1] : 0.0;
      }
    }
  }
  int ms = m_sib0;
  for (int j = 0; j < p.nsib; ++j) {
    const Child ch = p.sibs[j];
    if (ch.kind == CHILD_TIP) {
#pragma unroll
      for (int r = 0; r < RW; ++r) {
        const int code = load_code(p.codes, p.code_bytes, (long long)ch.idx * p.N + pat[r]);
        const double* tt = p.tiptab + (((size_t)ch.idx * C + c) * p.ncodes + code) * S;
#pragma unroll
        for (int nb = 0; nb < NBLK; ++nb) {
          const int x = nb * 8 + 2 * q;
          acc[r][nb][0] = x < S ? tt[x] : 0.0;
          acc[r][nb][1] = x + 1 < S ? tt[x + 1] : 0.0;
        }
      }
    } else {
#pragma unroll
      for (int r = 0; r < RW; ++r) {
        load_a_row<KB>(p.keep + (((size_t)ch.idx * p.N + pat[r]) * C + c) * S, S, q, a[r]);
        Eu[r] += p.keep_exp[((size_t)ch.idx * p.N + pat[r]) * C + c];
#pragma unroll
        for (int nb = 0; nb < NBLK; ++nb) acc[r][nb][0] = acc[r][nb][1] = 0.0;
      }
      dmma_rows_times_matrix<KB, NBLK, RW>(acc, a, sm_up + (size_t)ms * MAT, g, q);
      ++ms;
    }
#pragma unroll
    for (int r = 0; r < RW; ++r)
#pragma unroll
      for (int nb = 0; nb < NBLK; ++nb) {
        U[r][nb][0] *= acc[r][nb][0];
        U[r][nb][1] *= acc[r][nb][1];
      }
  }
#pragma unroll
  for (int r = 0; r < RW; ++r) {
    int m = 0;
#pragma unroll
    for (int nb = 0; nb < NBLK; ++nb) m = max(m, max(hi_word(U[r][nb][0]), hi_word(U[r][nb][1])));
    m = max(m, __shfl_xor_sync(0xffffffffu, m, 1));
    m = max(m, __shfl_xor_sync(0xffffffffu, m, 2));
    if (m < kScaleThresholdHi && m >= (1 << 20)) {
      const int k = rescale_shift(m);
      const double f = pow2(k);
#pragma unroll
      for (int nb = 0; nb < NBLK; ++nb) {
        U[r][nb][0] *= f;
        U[r][nb][1] *= f;
      }
      Eu[r] += k;
    }
    if (p.upper_out >= 0 && pat_base + r * 8 + g < p.N) {
      double* row = p.upper + (((size_t)p.upper_out * p.N + pat[r]) * C + c) * S;
#pragma unroll
      for (int nb = 0; nb < NBLK; ++nb) st2(row, nb * 8 + 2 * q, S, (S & 1) == 0, U[r][nb][0], U[r][nb][1]);
      if (q == 0) p.upper_exp[((size_t)p.upper_out * p.N + pat[r]) * C + c] = Eu[r];
    }
  }

  // ---- derivatives of branch n -----------------------------------------------------------------------------------
  if (!(p.want & 6u)) continue;
  int El[RW];
  size_t tipoff[RW];
  if (p.node_is_tip) {
#pragma unroll
    for (int r = 0; r < RW; ++r) {
      El[r] = 0;
      const int code = load_code(p.codes, p.code_bytes, (long long)p.node_idx * p.N + pat[r]);
      tipoff[r] = (((size_t)p.node_idx * C + c) * p.ncodes + code) * S;
    }
  } else {
#pragma unroll
    for (int r = 0; r < RW; ++r) {
      tipoff[r] = 0;
      load_a_row<KB>(p.keep + (((size_t)p.node_idx * p.N + pat[r]) * C + c) * S, S, q, a[r]);
      El[r] = p.keep_exp[((size_t)p.node_idx * p.N + pat[r]) * C + c];
    }
  }
  // T0 (NH form only): which x have a zero denominator
  double zmask[RW][NBLK][2];
  if (p.nh_form) {
    if (p.node_is_tip) {
#pragma unroll
      for (int r = 0; r < RW; ++r) {
        const double* tt = p.tiptab + tipoff[r];
#pragma unroll
        for (int nb = 0; nb < NBLK; ++nb) {
          const int x = nb * 8 + 2 * q;
          zmask[r][nb][0] = (x < S && tt[x] != 0.0) ? 1.0 : 0.0;
          zmask[r][nb][1] = (x + 1 < S && tt[x + 1] != 0.0) ? 1.0 : 0.0;
        }
      }
    } else {
#pragma unroll
      for (int r = 0; r < RW; ++r)
#pragma unroll
        for (int nb = 0; nb < NBLK; ++nb) acc[r][nb][0] = acc[r][nb][1] = 0.0;
      dmma_rows_times_matrix<KB, NBLK, RW>(acc, a, sm_up + (size_t)m_p0 * MAT, g, q);
#pragma unroll
      for (int r = 0; r < RW; ++r)
#pragma unroll
        for (int nb = 0; nb < NBLK; ++nb) {
          zmask[r][nb][0] = acc[r][nb][0] != 0.0 ? 1.0 : 0.0;
          zmask[r][nb][1] = acc[r][nb][1] != 0.0 ? 1.0 : 0.0;
        }
    }
  }
  double sres[2][RW];
#pragma unroll
  for (int ord = 0; ord < 2; ++ord) {
#pragma unroll
    for (int r = 0; r < RW; ++r) sres[ord][r] = 0.0;
    if (!((p.want >> (ord + 1)) & 1u)) continue;
    if (p.node_is_tip) {
      const double* tab = ord == 0 ? p.dtiptab : p.d2tiptab;
#pragma unroll
      for (int r = 0; r < RW; ++r) {
        const double* tt = tab + tipoff[r];
#pragma unroll
        for (int nb = 0; nb < NBLK; ++nb) {
          const int x = nb * 8 + 2 * q;
          acc[r][nb][0] = x < S ? tt[x] : 0.0;
          acc[r][nb][1] = x + 1 < S ? tt[x + 1] : 0.0;
        }
      }
    } else {
#pragma unroll
      for (int r = 0; r < RW; ++r)
#pragma unroll
        for (int nb = 0; nb < NBLK; ++nb) acc[r][nb][0] = acc[r][nb][1] = 0.0;
      dmma_rows_times_matrix<KB, NBLK, RW>(acc, a, sm_up + (size_t)(ord == 0 ? m_d1 : m_d2) * MAT, g, q);
    }
#pragma unroll
    for (int r = 0; r < RW; ++r) {
      double s = 0.0;
#pragma unroll
      for (int nb = 0; nb < NBLK; ++nb) {
        if (p.nh_form) {
          s = fma(U[r][nb][0] * zmask[r][nb][0], acc[r][nb][0], s);
          s = fma(U[r][nb][1] * zmask[r][nb][1], acc[r][nb][1], s);
        } else {
          s = fma(U[r][nb][0], acc[r][nb][0], s);
          s = fma(U[r][nb][1], acc[r][nb][1], s);
        }
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      sres[ord][r] = s;
    }
  }
  if (q == 0) {
    const double pc = p.probs[c];
#pragma unroll
    for (int r = 0; r < RW; ++r) {
      if (pat_base + r * 8 + g >= p.N) continue;
      const int sh = p.rexp[pat[r]] - Eu[r] - El[r];
      const double sr = p.SR[pat[r]];
      double* o = p.dLc + ((size_t)pat[r] * C + c) * 2;
      o[0] = scalbn(sres[0][r], sh) * pc / sr;
      o[1] = scalbn(sres[1][r], sh) * pc / sr;
    }
  }
  }  // tiles
}

// thread = pattern: dL_i = sum_c dLc, d2L_i = sum_c d2Lc;  partial sums of w dL and w (d2L - dL^2)
__global__ void deriv_combine_kernel(const double* dLc, const double* weights, int C, long long N, double* part1,
                                     double* part2) {
  __shared__ double red[32];
  const long long pat = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double c1 = 0.0, c2 = 0.0;
  if (pat < N) {
    double d1 = 0.0, d2 = 0.0;
    for (int c = 0; c < C; ++c) {
      d1 += dLc[((size_t)pat * C + c) * 2];
      d2 += dLc[((size_t)pat * C + c) * 2 + 1];
    }
    const double w = weights[pat];
    c1 = w * d1;
    c2 = w * (d2 - d1 * d1);
  }
  const double b1 = block_sum(c1, red);
  const double b2 = block_sum(c2, red);
  if (threadIdx.x == 0) {
    part1[blockIdx.x] = b1;
    part2[blockIdx.x] = b2;
  }
}

// one launch finishing both sums of a branch
__global__ void finalize_sum2_kernel(const double* part1, const double* part2, int n, double* out1, double* out2) {
  __shared__ double red[32];
  double a1 = 0.0, a2 = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    a1 += part1[i];
    a2 += part2[i];
  }
  const double s1 = block_sum(a1, red);
  const double s2 = block_sum(a2, red);
  if (threadIdx.x == 0) {
    *out1 = s1;
    *out2 = s2;
  }
}

}  // namespace bppgpu
