// Oracle (TEST / BASELINE INFRASTRUCTURE, never linked into the product):
// C++ restatement of the reference's CPU likelihood path, keeping its data layout and loop nest so that
// timing it is a fair "reference CPU" figure.  The reference itself cannot be compiled here
// (needs bpp-core 2.4.1 + the fork's bpp-seq), so cpu_baseline.kind is "port".
//
// What is restated (paths relative to /root/reference/src/Bpp/Phyl/):
//   * AbstractHomogeneousTreeLikelihood::computeTransitionProbabilitiesForNode
//       Likelihood/AbstractHomogeneousTreeLikelihood.cpp:354-414  (per class: getPij_t -> RowMatrix copy -> element copy)
//   * AbstractSubstitutionModel::getPij_t/getdPij_dt/getd2Pij_dt2, diagonalisable branch
//       Model/AbstractSubstitutionModel.cpp:426-437, :499-506, :570-577 (MatrixTools::mult(V, exp(lambda*rate*t), V^-1, P))
//   * RHomogeneousTreeLikelihood::computeSubtreeLikelihood   Likelihood/RHomogeneousTreeLikelihood.cpp:802-863
//       nested vector<vector<vector<double>>> arrays [site][class][state]; reset to 1; for each son the
//       4-deep (site, class, x, y) loops with the row of pxy_ re-read per (i, c, x)
//   * leaf arrays expanded to [site][class][state] doubles   Likelihood/DRASRTreeLikelihoodData.cpp:274-308
//   * DRHomogeneousTreeLikelihood::computeRootLikelihood / getLogLikelihood   DRHomogeneousTreeLikelihood.cpp:653-719,
//       :170-186 (w_i log SR_i, std::sort, sum from the largest)
//   * DR prefix pass + computeTreeD[2]LikelihoodAtNode, :287-326, :373-411, :543-649, :723-815
//       (a fresh [N][C][S] larray per branch, like computeLikelihoodAtNode_)
// The optional power-of-two rescaling is the same rule as oracle/ref_likelihood.py::_rescale (the reference
// has none and returns -inf where the unscaled product underflows).
//
// Threads stand for "the same code run as independent processes over pattern shards" (the reference is
// single-threaded): each thread owns private arrays and recomputes P(t); only the final per-site logs are
// gathered and sorted on the caller's thread.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

typedef std::vector<double> Vdouble;
typedef std::vector<Vdouble> VVdouble;
typedef std::vector<VVdouble> VVVdouble;

namespace {

struct Problem {
  int S, C, nn, root, ncodes, code_bytes, scaled, want;  // want: 1 lnL, 2 +d1, 4 +d2
  long N;
  const int *child_off, *children;
  const void* codes;  // [nl][N], leaf slots in increasing node id
  const double* code_table;
  const unsigned* weights;
  const double *rates, *probs, *V, *Vinv, *ev, *ev_im, *brlen, *rootfreq;
  double model_rate;
  int chr_clamp, weighted_root;
};

// MatrixTools::mult(A, dia, up, lo, B, O): O = A . T . B with T tridiagonal (bpp-core); the complex-pair block form of
// AbstractSubstitutionModel::getPij_t (Model/AbstractSubstitutionModel.cpp:438-468)
void mult_tridiag(const double* A, const Vdouble& dia, const Vdouble& up, const Vdouble& lo, const double* B, int S, VVdouble& O) {
  VVdouble AT(S, Vdouble(S));
  for (int i = 0; i < S; i++)
    for (int k = 0; k < S; k++) {
      double s = A[i * S + k] * dia[k];
      if (k > 0) s += A[i * S + k - 1] * up[k - 1];
      if (k < S - 1) s += A[i * S + k + 1] * lo[k];
      AT[i][k] = s;
    }
  for (int i = 0; i < S; i++)
    for (int j = 0; j < S; j++) {
      double s = 0;
      for (int k = 0; k < S; k++) s += AT[i][k] * B[k * S + j];
      O[i][j] = s;
    }
}

// MatrixTools::mult(A, D, B, O): O = A . diag(D) . B  (bpp-core), then copied out as a RowMatrix
void mult_diag(const double* A, const Vdouble& D, const double* B, int S, VVdouble& O) {
  for (int i = 0; i < S; i++)
    for (int j = 0; j < S; j++) {
      double s = 0;
      for (int k = 0; k < S; k++) s += A[i * S + k] * B[k * S + j] * D[k];
      O[i][j] = s;
    }
}

struct Shard {
  const Problem* p;
  long i0, n;  // pattern range
  std::vector<VVVdouble> lik;      // per node [n][C][S]
  std::vector<std::vector<std::vector<int> > > ex;  // per node [n][C]
  std::vector<VVVdouble> pxy, dpxy, d2pxy;  // per node [C][S][S]
  Vdouble site_lnl, SR;
  std::vector<int> SRe;
  std::vector<Vdouble> dL, d2L;  // per node [n]
  std::vector<int> leaf_slot;

  void alloc() {
    const Problem& P = *p;
    lik.assign(P.nn, VVVdouble());
    ex.assign(P.nn, std::vector<std::vector<int> >());
    leaf_slot.assign(P.nn, -1);
    int nl = 0;
    for (int v = 0; v < P.nn; v++) {
      if (P.child_off[v + 1] == P.child_off[v]) leaf_slot[v] = nl++;
      lik[v].assign(n, VVdouble(P.C, Vdouble(P.S, 1.)));
      ex[v].assign(n, std::vector<int>(P.C, 0));
    }
    init_leaves();
    pxy.assign(P.nn, VVVdouble(P.C, VVdouble(P.S, Vdouble(P.S))));
    if (P.want & 2) dpxy = pxy;
    if (P.want & 4) d2pxy = pxy;
    site_lnl.assign(n, 0.);
    SR.assign(n, 0.);
    SRe.assign(n, 0);
  }

  // leaf initialisation (getInitValue through the code table), done once like setData(); the blocked driver
  // (refcpu_eval_blocks) re-points the same arrays at the next block of patterns with it
  void init_leaves() {
    const Problem& P = *p;
    for (int v = 0; v < P.nn; v++) {
      if (leaf_slot[v] < 0) continue;
      for (long i = 0; i < n; i++) {
        long off = (long)leaf_slot[v] * P.N + i0 + i;
        int code = P.code_bytes == 1 ? ((const unsigned char*)P.codes)[off] : ((const unsigned short*)P.codes)[off];
        for (int c = 0; c < P.C; c++)
          for (int x = 0; x < P.S; x++) lik[v][i][c][x] = P.code_table[code * P.S + x];
      }
    }
  }

  void transition_probabilities() {
    const Problem& P = *p;
    const int S = P.S;
    Vdouble D(S);
    VVdouble Q(S, Vdouble(S));
    for (int v = 0; v < P.nn; v++) {
      if (v == P.root) continue;
      double l = P.brlen[v];
      for (int c = 0; c < P.C; c++) {
        double t = l * P.rates[c];
        if (t == 0) {
          for (int x = 0; x < S; x++)
            for (int y = 0; y < S; y++) Q[x][y] = x == y ? 1. : 0.;
        } else if (P.ev_im) {
          Vdouble dia(S), up(S - 1 > 0 ? S - 1 : 1, 0.), lo(S - 1 > 0 ? S - 1 : 1, 0.);
          const double lt = P.model_rate * t;
          for (int k = 0; k < S; k++) {
            const double ex = std::exp(P.ev[k] * lt);
            if (P.ev_im[k] != 0. && k + 1 < S) {
              const double c = std::cos(P.ev_im[k] * lt), sn = std::sin(P.ev_im[k] * lt);
              dia[k] = dia[k + 1] = ex * c;
              up[k] = ex * sn;
              lo[k] = -ex * sn;
              k++;
            } else {
              dia[k] = ex;
            }
          }
          mult_tridiag(P.V, dia, up, lo, P.Vinv, S, Q);
        } else {
          for (int k = 0; k < S; k++) D[k] = std::exp(P.ev[k] * (P.model_rate * t));
          mult_diag(P.V, D, P.Vinv, S, Q);
        }
        if (P.chr_clamp)  // ChromosomeSubstitutionModel.cpp:903-916
          for (int x = 0; x < S; x++)
            for (int y = 0; y < S; y++) {
              if (Q[x][y] < 0) Q[x][y] = 1e-20;
              else if (Q[x][y] > 1) Q[x][y] = 1;
            }
        VVdouble Qc = Q;  // "RowMatrix<double> Q = model_->getPij_t(...)" copies
        for (int x = 0; x < S; x++)
          for (int y = 0; y < S; y++) pxy[v][c][x][y] = Qc[x][y];
        if (P.want & 2) {
          double rc = P.rates[c];
          for (int k = 0; k < S; k++) D[k] = P.model_rate * P.ev[k] * std::exp(P.ev[k] * (P.model_rate * t));
          mult_diag(P.V, D, P.Vinv, S, Q);
          VVdouble dQ = Q;
          for (int x = 0; x < S; x++)
            for (int y = 0; y < S; y++) dpxy[v][c][x][y] = rc * dQ[x][y];
        }
        if (P.want & 4) {
          double rc = P.rates[c];
          for (int k = 0; k < S; k++) {
            double a = P.model_rate * P.ev[k];
            D[k] = a * a * std::exp(P.ev[k] * (P.model_rate * t));
          }
          mult_diag(P.V, D, P.Vinv, S, Q);
          VVdouble d2Q = Q;
          for (int x = 0; x < S; x++)
            for (int y = 0; y < S; y++) d2pxy[v][c][x][y] = rc * rc * d2Q[x][y];
        }
      }
    }
  }

  void rescale(VVVdouble& a, std::vector<std::vector<int> >& e) {
    const Problem& P = *p;
    for (long i = 0; i < n; i++)
      for (int c = 0; c < P.C; c++) {
        double m = 0;
        for (int x = 0; x < P.S; x++) m = std::max(m, a[i][c][x]);
        if (m >= std::ldexp(1.0, -1022) && m < std::ldexp(1.0, -256)) {
          int k;
          std::frexp(m, &k);
          for (int x = 0; x < P.S; x++) a[i][c][x] = std::ldexp(a[i][c][x], -k);
          e[i][c] += -k;
        }
      }
  }

  // computeSubtreeLikelihood
  void subtree(int node) {
    const Problem& P = *p;
    int b = P.child_off[node], e = P.child_off[node + 1];
    if (b == e) return;
    VVVdouble* _likelihoods_node = &lik[node];
    for (long i = 0; i < n; i++) {
      VVdouble* _likelihoods_node_i = &(*_likelihoods_node)[i];
      for (int c = 0; c < P.C; c++) {
        Vdouble* _likelihoods_node_i_c = &(*_likelihoods_node_i)[c];
        for (int x = 0; x < P.S; x++) (*_likelihoods_node_i_c)[x] = 1.;
      }
      for (int c = 0; c < P.C; c++) ex[node][i][c] = 0;
    }
    for (int k = b; k < e; k++) {
      int son = P.children[k];
      subtree(son);
      VVVdouble* pxy__son = &pxy[son];
      VVVdouble* _likelihoods_son = &lik[son];
      for (long i = 0; i < n; i++) {
        VVdouble* _likelihoods_son_i = &(*_likelihoods_son)[i];
        VVdouble* _likelihoods_node_i = &(*_likelihoods_node)[i];
        for (int c = 0; c < P.C; c++) {
          Vdouble* _likelihoods_son_i_c = &(*_likelihoods_son_i)[c];
          Vdouble* _likelihoods_node_i_c = &(*_likelihoods_node_i)[c];
          VVdouble* pxy__son_c = &(*pxy__son)[c];
          for (int x = 0; x < P.S; x++) {
            Vdouble* pxy__son_c_x = &(*pxy__son_c)[x];
            double likelihood = 0;
            for (int y = 0; y < P.S; y++) likelihood += (*pxy__son_c_x)[y] * (*_likelihoods_son_i_c)[y];
            (*_likelihoods_node_i_c)[x] *= likelihood;
          }
        }
        for (int c = 0; c < P.C; c++) ex[node][i][c] += ex[son][i][c];
      }
    }
    if (P.scaled) rescale(lik[node], ex[node]);
  }

  void root_likelihood() {
    const Problem& P = *p;
    const VVVdouble& r = lik[P.root];
    Vdouble wfreq;
    const double* rootfreq = P.rootfreq;
    if (P.weighted_root) {  // setWeightedRootFreq (DRNonHomogeneousTreeLikelihood.cpp:927-962), one shard only
      int em = ex[P.root][0][0];
      for (long i = 0; i < n; i++)
        for (int c = 0; c < P.C; c++) em = std::min(em, ex[P.root][i][c]);
      wfreq.assign(P.S, 0.);
      double tot = 0;
      for (int x = 0; x < P.S; x++) {
        for (long i = 0; i < n; i++)
          for (int c = 0; c < P.C; c++) wfreq[x] += std::ldexp(r[i][c][x], -(ex[P.root][i][c] - em)) * P.probs[c];
        tot += wfreq[x];
      }
      for (int x = 0; x < P.S; x++) wfreq[x] /= tot;
      rootfreq = wfreq.data();
    }
    for (long i = 0; i < n; i++) {
      int emin = ex[P.root][i][0];
      for (int c = 1; c < P.C; c++) emin = std::min(emin, ex[P.root][i][c]);
      double sr = 0;
      for (int c = 0; c < P.C; c++) {
        double s = 0;
        for (int x = 0; x < P.S; x++) s += r[i][c][x] * rootfreq[x];
        sr += std::ldexp(s, -(ex[P.root][i][c] - emin)) * P.probs[c];
      }
      if (sr < 0) sr = 0;
      SR[i] = sr;
      SRe[i] = emin;
      site_lnl[i] = std::log(sr) - emin * 0.693147180559945309417232121458;
    }
  }

  // ---- DR part: prefix pass and branch derivatives -----------------------------------------------
  std::vector<VVVdouble> upper;
  std::vector<std::vector<std::vector<int> > > uex;

  void contract(const VVVdouble& pm, const VVVdouble& src, bool transposed, VVVdouble& dst, bool first) {
    const Problem& P = *p;
    for (long i = 0; i < n; i++)
      for (int c = 0; c < P.C; c++)
        for (int x = 0; x < P.S; x++) {
          double l = 0;
          if (!transposed)
            for (int y = 0; y < P.S; y++) l += pm[c][x][y] * src[i][c][y];
          else
            for (int y = 0; y < P.S; y++) l += pm[c][y][x] * src[i][c][y];
          if (first) dst[i][c][x] = l;
          else dst[i][c][x] *= l;
        }
  }

  void prefix(int node) {  // fills upper[] of node's sons, fathers first
    const Problem& P = *p;
    int b = P.child_off[node], e = P.child_off[node + 1];
    for (int k = b; k < e; k++) {
      int son = P.children[k];
      VVVdouble& u = upper[son];
      std::vector<std::vector<int> >& ue = uex[son];
      for (long i = 0; i < n; i++)
        for (int c = 0; c < P.C; c++) ue[i][c] = 0;
      bool first = true;
      for (int k2 = b; k2 < e; k2++) {
        int sib = P.children[k2];
        if (sib == son) continue;
        contract(pxy[sib], lik[sib], false, u, first);
        first = false;
        for (long i = 0; i < n; i++)
          for (int c = 0; c < P.C; c++) ue[i][c] += ex[sib][i][c];
      }
      if (node != P.root) {
        contract(pxy[node], upper[node], true, u, first);
        for (long i = 0; i < n; i++)
          for (int c = 0; c < P.C; c++) ue[i][c] += uex[node][i][c];
      } else {
        for (long i = 0; i < n; i++)
          for (int c = 0; c < P.C; c++)
            for (int x = 0; x < P.S; x++) u[i][c][x] *= P.rootfreq[x];
      }
      if (P.scaled) rescale(u, ue);
      prefix(son);
    }
  }

  void derivatives() {
    const Problem& P = *p;
    upper.assign(P.nn, VVVdouble());
    uex.assign(P.nn, std::vector<std::vector<int> >());
    for (int v = 0; v < P.nn; v++)
      if (v != P.root) {
        upper[v].assign(n, VVdouble(P.C, Vdouble(P.S, 1.)));
        uex[v].assign(n, std::vector<int>(P.C, 0));
      }
    prefix(P.root);
    dL.assign(P.nn, Vdouble());
    d2L.assign(P.nn, Vdouble());
    for (int v = 0; v < P.nn; v++) {
      if (v == P.root) continue;
      dL[v].assign(n, 0.);
      if (P.want & 4) d2L[v].assign(n, 0.);
      for (int order = 1; order <= ((P.want & 4) ? 2 : 1); order++) {
        const VVVdouble& dp = order == 1 ? dpxy[v] : d2pxy[v];
        for (long i = 0; i < n; i++) {
          double acc = 0;
          for (int c = 0; c < P.C; c++) {
            double s = 0;
            for (int x = 0; x < P.S; x++) {
              double d = 0;
              for (int y = 0; y < P.S; y++) d += dp[c][x][y] * lik[v][i][c][y];
              s += upper[v][i][c][x] * d;
            }
            acc += std::ldexp(s, SRe[i] - uex[v][i][c] - ex[v][i][c]) * P.probs[c];
          }
          double val = acc / SR[i];
          (order == 1 ? dL[v] : d2L[v])[i] = val;
        }
      }
    }
  }

  void run() {
    transition_probabilities();
    subtree(p->root);
    root_likelihood();
    if (p->want & 6) derivatives();
  }
};

}  // namespace

extern "C" int refcpu_eval(int S, int C, long N, int nn, int root, const int* child_off, const int* children,
                           const void* codes, int code_bytes, int ncodes, const double* code_table,
                           const unsigned* weights, const double* rates, const double* probs, const double* V,
                           const double* Vinv, const double* ev, const double* ev_im /*NULL = real*/, int chr_clamp,
                           int weighted_root, double model_rate, const double* brlen,
                           const double* rootfreq, int scaled, int want, int nthreads, int reps, double* lnl_out,
                           double* d1_out /*[nn] or NULL*/, double* d2_out /*[nn] or NULL*/,
                           double* site_lnl_out /*[N] or NULL*/, double* best_seconds, double* sum_seconds /*all reps, or NULL*/) {
  Problem P;
  P.S = S; P.C = C; P.N = N; P.nn = nn; P.root = root; P.child_off = child_off; P.children = children;
  P.codes = codes; P.code_bytes = code_bytes; P.ncodes = ncodes; P.code_table = code_table; P.weights = weights;
  P.rates = rates; P.probs = probs; P.V = V; P.Vinv = Vinv; P.ev = ev; P.ev_im = ev_im; P.chr_clamp = chr_clamp;
  P.weighted_root = weighted_root; P.model_rate = model_rate; P.brlen = brlen;
  P.rootfreq = rootfreq; P.scaled = scaled; P.want = want;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > N && N > 0) nthreads = (int)N;
  std::vector<Shard> sh(nthreads);
  for (int t = 0; t < nthreads; t++) {
    sh[t].p = &P;
    sh[t].i0 = N * t / nthreads;
    sh[t].n = N * (t + 1) / nthreads - sh[t].i0;
  }
  {
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) th.emplace_back([&sh, t] { sh[t].alloc(); });
    for (auto& x : th) x.join();
  }
  double best = 1e300, lnl = 0, total = 0;
  Vdouble la(N);
  for (int r = 0; r < reps; r++) {
    auto t0 = std::chrono::steady_clock::now();
    if (nthreads == 1) sh[0].run();
    else {
      std::vector<std::thread> th;
      for (int t = 0; t < nthreads; t++) th.emplace_back([&sh, t] { sh[t].run(); });
      for (auto& x : th) x.join();
    }
    // getLogLikelihood: la[i] = w_i log SR_i; sort; add from the largest
    for (int t = 0; t < nthreads; t++)
      for (long i = 0; i < sh[t].n; i++) la[sh[t].i0 + i] = weights[sh[t].i0 + i] * sh[t].site_lnl[i];
    Vdouble srt = la;
    std::sort(srt.begin(), srt.end());
    lnl = 0;
    for (long i = N; i > 0; i--) lnl += srt[i - 1];
    if (want & 6) {
      for (int v = 0; v < nn; v++) {
        if (v == root) continue;
        double d = 0, d2 = 0;
        for (int t = 0; t < nthreads; t++)
          for (long i = 0; i < sh[t].n; i++) {
            double w = weights[sh[t].i0 + i], a = sh[t].dL[v][i];
            d += w * a;
            if (want & 4) d2 += w * (sh[t].d2L[v][i] - a * a);
          }
        if (d1_out) d1_out[v] = -d;
        if (d2_out && (want & 4)) d2_out[v] = -d2;
      }
    }
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    best = std::min(best, sec);
    total += sec;
  }
  *lnl_out = lnl;
  if (site_lnl_out)
    for (int t = 0; t < nthreads; t++)
      for (long i = 0; i < sh[t].n; i++) site_lnl_out[sh[t].i0 + i] = sh[t].site_lnl[i];
  if (best_seconds) *best_seconds = best;
  if (sum_seconds) *sum_seconds = total;
  return 0;
}

// The whole of a large alignment in blocks of `block` patterns: the nested arrays are allocated ONCE (the reference allocates
// them in setData, outside any evaluation; re-allocating them per block would cost far more than the arithmetic) and re-pointed
// at the next block by the leaf initialisation; value only.  lnl_out = sum over blocks of the block's sorted sum;
// sum_seconds = evaluation time only (P(t) + pruning + root), like refcpu_eval.
extern "C" int refcpu_eval_blocks(int S, int C, long N, long block, int nn, int root, const int* child_off, const int* children,
                                  const void* codes, int code_bytes, int ncodes, const double* code_table,
                                  const unsigned* weights, const double* rates, const double* probs, const double* V,
                                  const double* Vinv, const double* ev, const double* ev_im, int chr_clamp, double model_rate,
                                  const double* brlen, const double* rootfreq, int scaled, int nthreads, double* lnl_out,
                                  double* sum_seconds) {
  Problem P;
  P.S = S; P.C = C; P.N = N; P.nn = nn; P.root = root; P.child_off = child_off; P.children = children;
  P.codes = codes; P.code_bytes = code_bytes; P.ncodes = ncodes; P.code_table = code_table; P.weights = weights;
  P.rates = rates; P.probs = probs; P.V = V; P.Vinv = Vinv; P.ev = ev; P.ev_im = ev_im; P.chr_clamp = chr_clamp;
  P.weighted_root = 0; P.model_rate = model_rate; P.brlen = brlen;
  P.rootfreq = rootfreq; P.scaled = scaled; P.want = 1;
  if (block > N) block = N;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > block && block > 0) nthreads = (int)block;
  std::vector<Shard> sh(nthreads);
  std::vector<long> cap(nthreads);
  for (int t = 0; t < nthreads; t++) {
    sh[t].p = &P;
    sh[t].i0 = block * t / nthreads;
    sh[t].n = cap[t] = block * (t + 1) / nthreads - sh[t].i0;
  }
  {
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) th.emplace_back([&sh, t] { sh[t].alloc(); });
    for (auto& x : th) x.join();
  }
  double lnl = 0, total = 0;
  for (long b0 = 0; b0 < N; b0 += block) {
    const long nb = std::min(block, N - b0);
    for (int t = 0; t < nthreads; t++) {
      const long lo = nb * t / nthreads, hi = nb * (t + 1) / nthreads;
      sh[t].i0 = b0 + lo;
      sh[t].n = std::min(hi - lo, cap[t]);
    }
    // the remainder of a short last block that does not fit a shard's capacity goes to the shards in turn (never
    // happens for block sizes that are multiples of the thread count; kept for safety)
    {
      std::vector<std::thread> th;
      for (int t = 0; t < nthreads; t++) th.emplace_back([&sh, t] { if (sh[t].n > 0) sh[t].init_leaves(); });
      for (auto& x : th) x.join();
    }
    auto t0 = std::chrono::steady_clock::now();
    {
      std::vector<std::thread> th;
      for (int t = 0; t < nthreads; t++) th.emplace_back([&sh, t] { if (sh[t].n > 0) sh[t].run(); });
      for (auto& x : th) x.join();
    }
    Vdouble la;
    la.reserve(nb);
    long covered = 0;
    for (int t = 0; t < nthreads; t++) {
      for (long i = 0; i < sh[t].n; i++) la.push_back(weights[sh[t].i0 + i] * sh[t].site_lnl[i]);
      covered += sh[t].n;
    }
    if (covered != nb) return 2;
    std::sort(la.begin(), la.end());
    double bl = 0;
    for (long i = (long)la.size(); i > 0; i--) bl += la[i - 1];
    lnl += bl;
    total += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
  *lnl_out = lnl;
  if (sum_seconds) *sum_seconds = total;
  return 0;
}
