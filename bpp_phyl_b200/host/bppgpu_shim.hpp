// bppgpu_shim.hpp -- C++ host side above the C ABI (include/bppgpu.h), mirroring the reference's class surface for
// the tree-likelihood hot path so that code written against bpp-phyl reads the same:
//
//   reference (src/Bpp/Phyl/...)                                   here (namespace bppshim)
//   Model/SubstitutionModel.h:195-525  TransitionModel / SubstitutionModel   SubstitutionModel (getPij_t ... on the GPU)
//   Model/AbstractSubstitutionModel.cpp:175-421  updateMatrices                AbstractSubstitutionModel::updateMatrices (host)
//   Model/Nucleotide/{JCnuc,K80,HKY85,T92,GTR}.cpp, Model/Protein/LG08.cpp    same names
//   Model/Codon/YN98.cpp (+ AbstractWord/Codon*SubstitutionModel)              YN98
//   Model/ChromosomeSubstitutionModel.cpp:431-802                              ChromosomeSubstitutionModel
//   Model/RateDistribution/{Constant,GammaDiscrete}RateDistribution.h          same names
//   TreeTemplate.h / TreeTemplateTools (parenthesisToTree, unroot, getNodes)   TreeTemplate<Node>, TreeTemplateTools
//   SitePatterns.cpp:52-106                                                    SitePatterns (via bppgpu_site_patterns)
//   Likelihood/{R,DR}HomogeneousTreeLikelihood, DRNonHomogeneousTreeLikelihood same names; computeTreeLikelihood,
//       fireParameterChanged, getValue, get{First,Second}OrderDerivative ... run on the device through libbppgpu
//
// Only what the hot path needs is here; bpp-core / bpp-seq types the signatures mention are
// reduced to minimal stand-ins (RowMatrix, Alphabet, VectorSiteContainer, DiscreteDistribution, ParameterList-like
// name/value access).  There is no CPU fallback: every evaluation goes through the C ABI and throws bppshim::Exception
// with the library's message when no B200 is usable.
#pragma once
#include <algorithm>
#include <cmath>
#include <limits>
#include <complex>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <random>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/bppgpu.h"

namespace bppshim {

typedef std::vector<double> Vdouble;
typedef std::vector<Vdouble> VVdouble;
typedef std::vector<VVdouble> VVVdouble;

class Exception : public std::runtime_error {
 public:
  explicit Exception(const std::string& m) : std::runtime_error(m) {}
};
class ParameterNotFoundException : public Exception {
 public:
  explicit ParameterNotFoundException(const std::string& m) : Exception(m) {}
};

inline void check(int rc, const char* where) {
  if (rc != BPPGPU_OK) throw Exception(std::string(where) + ": " + bppgpu_last_error());
}

// ---- bpp-core stand-ins ---------------------------------------------------------------------------------------
template <class T>
class RowMatrix {
 public:
  RowMatrix() : r_(0), c_(0) {}
  RowMatrix(size_t r, size_t c) : r_(r), c_(c), d_(r * c) {}
  void resize(size_t r, size_t c) { r_ = r; c_ = c; d_.assign(r * c, T()); }
  T& operator()(size_t i, size_t j) { return d_[i * c_ + j]; }
  const T& operator()(size_t i, size_t j) const { return d_[i * c_ + j]; }
  size_t getNumberOfRows() const { return r_; }
  size_t getNumberOfColumns() const { return c_; }
  T* data() { return d_.data(); }
  const T* data() const { return d_.data(); }

 private:
  size_t r_, c_;
  std::vector<T> d_;
};

namespace NumConstants {
inline double TINY() { return 1e-12; }
inline double SMALL() { return 1e-6; }
inline double VERY_TINY() { return 1e-20; }
}  // namespace NumConstants

// ---- host linear algebra (updateMatrices stays on the host, north-star (1)) --------------------------------------
namespace linalg {

// cyclic Jacobi for a symmetric matrix: A = U diag(w) U^T, columns of U are eigenvectors
inline void jacobi_symmetric(std::vector<double> a, int n, std::vector<double>& w, std::vector<double>& U) {
  U.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) U[(size_t)i * n + i] = 1.0;
  for (int sweep = 0; sweep < 100; ++sweep) {
    double off = 0.0;
    for (int i = 0; i < n; ++i)
      for (int j = i + 1; j < n; ++j) off += a[(size_t)i * n + j] * a[(size_t)i * n + j];
    if (off < 1e-300) break;
    for (int p = 0; p < n; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = a[(size_t)p * n + q];
        if (std::fabs(apq) < 1e-300) continue;
        const double theta = (a[(size_t)q * n + q] - a[(size_t)p * n + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; ++k) {
          const double akp = a[(size_t)k * n + p], akq = a[(size_t)k * n + q];
          a[(size_t)k * n + p] = c * akp - s * akq;
          a[(size_t)k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {
          const double apk = a[(size_t)p * n + k], aqk = a[(size_t)q * n + k];
          a[(size_t)p * n + k] = c * apk - s * aqk;
          a[(size_t)q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          const double ukp = U[(size_t)k * n + p], ukq = U[(size_t)k * n + q];
          U[(size_t)k * n + p] = c * ukp - s * ukq;
          U[(size_t)k * n + q] = s * ukp + c * ukq;
        }
      }
  }
  w.resize(n);
  for (int i = 0; i < n; ++i) w[i] = a[(size_t)i * n + i];
}

// LU inverse with partial pivoting; returns false when singular to working precision
inline bool invert(const std::vector<double>& A, int n, std::vector<double>& inv) {
  std::vector<double> a(A);
  inv.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) inv[(size_t)i * n + i] = 1.0;
  for (int col = 0; col < n; ++col) {
    int piv = col;
    double best = std::fabs(a[(size_t)col * n + col]);
    for (int r = col + 1; r < n; ++r)
      if (std::fabs(a[(size_t)r * n + col]) > best) { best = std::fabs(a[(size_t)r * n + col]); piv = r; }
    if (!(best > 1e-300)) return false;
    if (piv != col)
      for (int k = 0; k < n; ++k) {
        std::swap(a[(size_t)piv * n + k], a[(size_t)col * n + k]);
        std::swap(inv[(size_t)piv * n + k], inv[(size_t)col * n + k]);
      }
    const double d = 1.0 / a[(size_t)col * n + col];
    for (int k = 0; k < n; ++k) { a[(size_t)col * n + k] *= d; inv[(size_t)col * n + k] *= d; }
    for (int r = 0; r < n; ++r) {
      if (r == col) continue;
      const double f = a[(size_t)r * n + col];
      if (f == 0.0) continue;
      for (int k = 0; k < n; ++k) { a[(size_t)r * n + k] -= f * a[(size_t)col * n + k]; inv[(size_t)r * n + k] -= f * inv[(size_t)col * n + k]; }
    }
  }
  for (double v : inv)
    if (!std::isfinite(v)) return false;
  // two Newton-Schulz refinement steps X <- X + X (I - A X): the elimination above loses digits on the badly
  // conditioned eigenvector bases of non-normal generators (chromosome models reach cond(V) ~ 1e5)
  for (int it = 0; it < 2; ++it) {
    std::vector<double> R((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < n; ++k) {
        const double aik = A[(size_t)i * n + k];
        if (aik == 0.0) continue;
        for (int j = 0; j < n; ++j) R[(size_t)i * n + j] -= aik * inv[(size_t)k * n + j];
      }
    for (int i = 0; i < n; ++i) R[(size_t)i * n + i] += 1.0;
    std::vector<double> X(inv);
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < n; ++k) {
        const double xik = inv[(size_t)i * n + k];
        if (xik == 0.0) continue;
        for (int j = 0; j < n; ++j) X[(size_t)i * n + j] += xik * R[(size_t)k * n + j];
      }
    for (double v : X)
      if (!std::isfinite(v)) return true;  // keep the unrefined inverse
    inv.swap(X);
  }
  return true;
}

// eigenvalues of a general real matrix: reduction to Hessenberg form by stabilised elimination, then the
// Francis double-shift QR iteration (the classical EISPACK elmhes / hqr pair)
inline bool hessenberg_qr_eigenvalues(std::vector<double> a, int n, std::vector<double>& wr, std::vector<double>& wi) {
  auto A = [&](int i, int j) -> double& { return a[(size_t)i * n + j]; };
  for (int m = 1; m < n - 1; ++m) {
    double x = 0.0;
    int i = m;
    for (int j = m; j < n; ++j)
      if (std::fabs(A(j, m - 1)) > std::fabs(x)) { x = A(j, m - 1); i = j; }
    if (i != m) {
      for (int j = m - 1; j < n; ++j) std::swap(A(i, j), A(m, j));
      for (int j = 0; j < n; ++j) std::swap(A(j, i), A(j, m));
    }
    if (x != 0.0)
      for (i = m + 1; i < n; ++i) {
        double y = A(i, m - 1);
        if (y != 0.0) {
          y /= x;
          A(i, m - 1) = y;
          for (int j = m; j < n; ++j) A(i, j) -= y * A(m, j);
          for (int j = 0; j < n; ++j) A(j, m) += y * A(j, i);
        }
      }
  }
  for (int i = 2; i < n; ++i)
    for (int j = 0; j < i - 1; ++j) A(i, j) = 0.0;
  wr.assign(n, 0.0);
  wi.assign(n, 0.0);
  double anorm = 0.0;
  for (int i = 0; i < n; ++i)
    for (int j = std::max(i - 1, 0); j < n; ++j) anorm += std::fabs(A(i, j));
  int nn = n - 1;
  double t = 0.0, p = 0, q = 0, r = 0, s = 0, w = 0, x = 0, y = 0, z = 0;
  while (nn >= 0) {
    int its = 0, l;
    do {
      for (l = nn; l >= 1; --l) {
        s = std::fabs(A(l - 1, l - 1)) + std::fabs(A(l, l));
        if (s == 0.0) s = anorm;
        if (std::fabs(A(l, l - 1)) + s == s) { A(l, l - 1) = 0.0; break; }
      }
      x = A(nn, nn);
      if (l == nn) {
        wr[nn] = x + t; wi[nn--] = 0.0;
      } else {
        y = A(nn - 1, nn - 1);
        w = A(nn, nn - 1) * A(nn - 1, nn);
        if (l == nn - 1) {
          p = 0.5 * (y - x);
          q = p * p + w;
          z = std::sqrt(std::fabs(q));
          x += t;
          if (q >= 0.0) {
            z = p + (p >= 0 ? std::fabs(z) : -std::fabs(z));
            wr[nn - 1] = wr[nn] = x + z;
            if (z != 0.0) wr[nn] = x - w / z;
            wi[nn - 1] = wi[nn] = 0.0;
          } else {
            wr[nn - 1] = wr[nn] = x + p;
            wi[nn - 1] = z;       // +im first, like JAMA's EigenValue
            wi[nn] = -z;
          }
          nn -= 2;
        } else {
          if (its == 60) return false;
          if (its == 10 || its == 20) {
            t += x;
            for (int i = 0; i <= nn; ++i) A(i, i) -= x;
            s = std::fabs(A(nn, nn - 1)) + std::fabs(A(nn - 1, nn - 2));
            y = x = 0.75 * s;
            w = -0.4375 * s * s;
          }
          ++its;
          int m;
          for (m = nn - 2; m >= l; --m) {
            z = A(m, m);
            r = x - z;
            s = y - z;
            p = (r * s - w) / A(m + 1, m) + A(m, m + 1);
            q = A(m + 1, m + 1) - z - r - s;
            r = A(m + 2, m + 1);
            s = std::fabs(p) + std::fabs(q) + std::fabs(r);
            p /= s; q /= s; r /= s;
            if (m == l) break;
            const double u = std::fabs(A(m, m - 1)) * (std::fabs(q) + std::fabs(r));
            const double v = std::fabs(p) * (std::fabs(A(m - 1, m - 1)) + std::fabs(z) + std::fabs(A(m + 1, m + 1)));
            if (u + v == v) break;
          }
          for (int i = m + 2; i <= nn; ++i) {
            A(i, i - 2) = 0.0;
            if (i != m + 2) A(i, i - 3) = 0.0;
          }
          for (int k = m; k <= nn - 1; ++k) {
            if (k != m) {
              p = A(k, k - 1);
              q = A(k + 1, k - 1);
              r = 0.0;
              if (k != nn - 1) r = A(k + 2, k - 1);
              if ((x = std::fabs(p) + std::fabs(q) + std::fabs(r)) != 0.0) { p /= x; q /= x; r /= x; }
            }
            const double sg = std::sqrt(p * p + q * q + r * r);
            s = p >= 0 ? sg : -sg;
            if (s != 0.0) {
              if (k == m) {
                if (l != m) A(k, k - 1) = -A(k, k - 1);
              } else {
                A(k, k - 1) = -s * x;
              }
              p += s;
              x = p / s; y = q / s; z = r / s;
              q /= p; r /= p;
              for (int j = k; j <= nn; ++j) {
                p = A(k, j) + q * A(k + 1, j);
                if (k != nn - 1) { p += r * A(k + 2, j); A(k + 2, j) -= p * z; }
                A(k + 1, j) -= p * y;
                A(k, j) -= p * x;
              }
              const int mmin = nn < k + 3 ? nn : k + 3;
              for (int i = l; i <= mmin; ++i) {
                p = x * A(i, k) + y * A(i, k + 1);
                if (k != nn - 1) { p += z * A(i, k + 2); A(i, k + 2) -= p * r; }
                A(i, k + 1) -= p * q;
                A(i, k) -= p;
              }
            }
          }
        }
      }
    } while (l < nn - 1);
  }
  return true;
}

// complex LU solve of (A - lambda I) x = b, used for inverse iteration
inline bool solve_shifted(const std::vector<double>& A, int n, std::complex<double> lambda, std::vector<std::complex<double>>& x) {
  typedef std::complex<double> cd;
  std::vector<cd> a((size_t)n * n);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) a[(size_t)i * n + j] = cd(A[(size_t)i * n + j]) - (i == j ? lambda : cd(0));
  for (int col = 0; col < n; ++col) {
    int piv = col;
    double best = std::abs(a[(size_t)col * n + col]);
    for (int r = col + 1; r < n; ++r)
      if (std::abs(a[(size_t)r * n + col]) > best) { best = std::abs(a[(size_t)r * n + col]); piv = r; }
    if (best < 1e-300) a[(size_t)piv * n + col] = cd(1e-300);
    if (piv != col) {
      for (int k = 0; k < n; ++k) std::swap(a[(size_t)piv * n + k], a[(size_t)col * n + k]);
      std::swap(x[piv], x[col]);
    }
    const cd d = cd(1.0) / a[(size_t)col * n + col];
    for (int r = col + 1; r < n; ++r) {
      const cd f = a[(size_t)r * n + col] * d;
      if (f == cd(0)) continue;
      for (int k = col; k < n; ++k) a[(size_t)r * n + k] -= f * a[(size_t)col * n + k];
      x[r] -= f * x[col];
    }
  }
  for (int i = n - 1; i >= 0; --i) {
    cd s = x[i];
    for (int k = i + 1; k < n; ++k) s -= a[(size_t)i * n + k] * x[k];
    x[i] = s / a[(size_t)i * n + i];
  }
  return true;
}

// Real eigen-form of a general real matrix as bpp-core's EigenValue<double> presents it: eigenvalues (re, im) and a
// REAL matrix V with A V = V D, D block diagonal ([[re, im], [-im, re]] for a conjugate pair, +im member first).
// Eigenvalues from the QR iteration above, vectors by inverse iteration.  Only V f(D) V^-1 matters for parity.
inline bool eigen_general(const std::vector<double>& A, int n, std::vector<double>& re, std::vector<double>& im, std::vector<double>& V) {
  typedef std::complex<double> cd;
  if (!hessenberg_qr_eigenvalues(A, n, re, im)) return false;
  // sort: descending real part keeps conjugates adjacent (+im first)
  std::vector<int> idx(n);
  for (int i = 0; i < n; ++i) idx[i] = i;
  std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) {
    if (re[a] != re[b]) return re[a] > re[b];
    return im[a] > im[b];
  });
  std::vector<double> r2(n), i2(n);
  for (int i = 0; i < n; ++i) { r2[i] = re[idx[i]]; i2[i] = im[idx[i]]; }
  re = r2; im = i2;
  V.assign((size_t)n * n, 0.0);
  double scale = 0.0;
  for (double v : A) scale = std::max(scale, std::fabs(v));
  if (scale == 0.0) scale = 1.0;
  for (int k = 0; k < n; ++k) {
    if (im[k] < 0.0) continue;  // second member of a pair: filled with the first
    const cd lam(re[k], im[k]);
    // inverse iteration; the eigenvalue itself is refined from the iteration (lambda = shift + <x,x>/<x,y> with
    // (A - shift I) y = x), which recovers the digits the unbalanced QR iteration loses on non-normal generators
    cd lamk = lam;
    std::vector<cd> x(n);
    for (int i = 0; i < n; ++i) x[i] = cd(1.0 + 0.37 * ((i * 7919 + k * 104729) % 101) / 101.0, 0.0);
    for (int it = 0; it < 5; ++it) {
      // a shift a few ulps off the current eigenvalue keeps (A - shift I) numerically invertible
      const double eps = std::max(std::abs(lamk), scale * 1e-3) * 4e-15 * (1.0 + (k % 7));
      const cd shifted = lamk + cd(eps, im[k] != 0.0 ? eps : 0.0);
      std::vector<cd> y = x;
      solve_shifted(A, n, shifted, y);
      cd xy(0), xx(0);
      for (int i = 0; i < n; ++i) { xy += std::conj(x[i]) * y[i]; xx += std::conj(x[i]) * x[i]; }
      double nrm = 0.0;
      for (auto& v : y) nrm = std::max(nrm, std::abs(v));
      if (!(nrm > 0.0) || !std::isfinite(nrm)) return false;
      for (int i = 0; i < n; ++i) x[i] = y[i] / nrm;
      if (it >= 1 && std::abs(xy) > 0.0) {
        cd l2 = shifted + xx / xy;
        if (im[k] == 0.0) l2 = cd(l2.real(), 0.0);
        if (std::abs(l2 - lam) <= 1e-6 * std::max(std::abs(lam), scale)) lamk = l2;  // stay on this eigenvalue
      }
    }
    re[k] = lamk.real();
    if (im[k] != 0.0) { im[k] = lamk.imag(); re[k + 1] = lamk.real(); im[k + 1] = -lamk.imag(); }
    if (im[k] == 0.0) {
      // rotate to a real vector
      cd ph(0);
      double best = 0;
      for (auto& v : x)
        if (std::abs(v) > best) { best = std::abs(v); ph = v; }
      for (int i = 0; i < n; ++i) V[(size_t)i * n + k] = (x[i] / ph).real();
    } else {
      if (k + 1 >= n) return false;
      for (int i = 0; i < n; ++i) {
        V[(size_t)i * n + k] = x[i].real();
        V[(size_t)i * n + k + 1] = x[i].imag();
      }
    }
  }
  return true;
}

inline std::vector<double> matmul(const std::vector<double>& A, const std::vector<double>& B, int n) {
  std::vector<double> C((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < n; ++k) {
      const double a = A[(size_t)i * n + k];
      if (a == 0.0) continue;
      for (int j = 0; j < n; ++j) C[(size_t)i * n + j] += a * B[(size_t)k * n + j];
    }
  return C;
}

// regularised lower incomplete gamma P(a, x) (series / continued fraction) and its inverse
inline double lgamma_(double x) { return std::lgamma(x); }
inline double gammp(double a, double x) {
  if (x <= 0) return 0.0;
  if (x < a + 1.0) {
    double ap = a, sum = 1.0 / a, del = sum;
    for (int n = 0; n < 1000; ++n) {
      ap += 1.0;
      del *= x / ap;
      sum += del;
      if (std::fabs(del) < std::fabs(sum) * 1e-17) break;
    }
    return sum * std::exp(-x + a * std::log(x) - lgamma_(a));
  }
  double b = x + 1.0 - a, c = 1.0 / 1e-300, d = 1.0 / b, h = d;
  for (int i = 1; i < 1000; ++i) {
    const double an = -i * (i - a);
    b += 2.0;
    d = an * d + b;
    if (std::fabs(d) < 1e-300) d = 1e-300;
    c = b + an / c;
    if (std::fabs(c) < 1e-300) c = 1e-300;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (std::fabs(del - 1.0) < 1e-17) break;
  }
  return 1.0 - std::exp(-x + a * std::log(x) - lgamma_(a)) * h;
}
inline double gammp_inv(double a, double p) {
  if (p <= 0) return 0.0;
  if (p >= 1) return INFINITY;
  double lo = 0.0, hi = std::max(1.0, a);
  while (gammp(a, hi) < p) hi *= 2.0;
  double x = 0.5 * (lo + hi);
  for (int it = 0; it < 200; ++it) {
    const double f = gammp(a, x) - p;
    if (f > 0) hi = x; else lo = x;
    // Newton step with the density, safeguarded by the bracket
    const double dens = std::exp(-x + (a - 1.0) * std::log(x) - lgamma_(a));
    double xn = dens > 0 ? x - f / dens : 0.5 * (lo + hi);
    if (!(xn > lo && xn < hi)) xn = 0.5 * (lo + hi);
    if (std::fabs(xn - x) <= 1e-16 * std::fabs(x)) { x = xn; break; }
    x = xn;
  }
  return x;
}

}  // namespace linalg

// ---- bpp-seq stand-ins ------------------------------------------------------------------------------------------
class Alphabet {
 public:
  virtual ~Alphabet() {}
  virtual size_t getSize() const = 0;                         // number of resolved states
  virtual unsigned getStateCodingSize() const { return 1; }   // characters per state
  // resolved states a character (string of getStateCodingSize() chars) stands for; empty = unknown character
  virtual std::vector<int> getAlias(const std::string& ch) const = 0;
};
class LetterAlphabet : public Alphabet {
 public:
  LetterAlphabet(const std::string& states, const std::map<char, std::string>& aliases) : states_(states), aliases_(aliases) {}
  size_t getSize() const { return states_.size(); }
  std::vector<int> getAlias(const std::string& ch) const {
    std::vector<int> out;
    if (ch.size() != 1) return out;
    const char c = (char)std::toupper((unsigned char)ch[0]);
    const size_t p = states_.find(c);
    if (p != std::string::npos) { out.push_back((int)p); return out; }
    std::map<char, std::string>::const_iterator it = aliases_.find(c);
    if (it != aliases_.end())
      for (char r : it->second) out.push_back((int)states_.find(r));
    return out;
  }
  const std::string& states() const { return states_; }

 private:
  std::string states_;
  std::map<char, std::string> aliases_;
};
class DNA : public LetterAlphabet {
 public:
  DNA() : LetterAlphabet("ACGT", {{'U', "T"}, {'M', "AC"}, {'R', "AG"}, {'W', "AT"}, {'S', "CG"}, {'Y', "CT"}, {'K', "GT"},
                                  {'V', "ACG"}, {'H', "ACT"}, {'D', "AGT"}, {'B', "CGT"}, {'N', "ACGT"}, {'X', "ACGT"},
                                  {'O', "ACGT"}, {'0', "ACGT"}, {'?', "ACGT"}, {'-', "ACGT"}}) {}
};
class ProteicAlphabet : public LetterAlphabet {
 public:
  ProteicAlphabet() : LetterAlphabet("ARNDCQEGHILKMFPSTWYV", {{'B', "ND"}, {'Z', "QE"}, {'J', "IL"},
                                                               {'X', "ARNDCQEGHILKMFPSTWYV"}, {'O', "ARNDCQEGHILKMFPSTWYV"},
                                                               {'0', "ARNDCQEGHILKMFPSTWYV"}, {'?', "ARNDCQEGHILKMFPSTWYV"},
                                                               {'-', "ARNDCQEGHILKMFPSTWYV"}}) {}
};
// 64 codons, index 16*n1 + 4*n2 + n3 with A,C,G,T = 0..3; standard genetic code
class CodonAlphabet : public Alphabet {
 public:
  size_t getSize() const { return 64; }
  unsigned getStateCodingSize() const { return 3; }
  std::vector<int> getAlias(const std::string& ch) const {
    std::vector<int> out;
    if (ch.size() != 3) return out;
    std::vector<int> pos[3];
    DNA dna;
    for (int k = 0; k < 3; ++k) pos[k] = dna.getAlias(std::string(1, ch[k]));
    for (int a : pos[0]) for (int b : pos[1]) for (int c : pos[2]) out.push_back(16 * a + 4 * b + c);
    return out;
  }
  static char aminoAcid(int codon) {
    static const char* tcag = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
    static const int map_[4] = {2, 1, 3, 0};  // A,C,G,T -> position in T,C,A,G
    const int a = codon / 16, b = (codon / 4) % 4, c = codon % 4;
    return tcag[16 * map_[a] + 4 * map_[b] + map_[c]];
  }
  static bool isStop(int codon) { return aminoAcid(codon) == '*'; }
};
// chromosome counts min..max, one integer per taxon written in decimal ("X" = unknown)
class ChromosomeAlphabet : public Alphabet {
 public:
  ChromosomeAlphabet(unsigned mn, unsigned mx) : min_(mn), max_(mx) {}
  size_t getSize() const { return max_ - min_ + 1; }
  unsigned getStateCodingSize() const { return 0; }  // variable width: whitespace-separated counts
  unsigned getMin() const { return min_; }
  unsigned getMax() const { return max_; }
  std::vector<int> getAlias(const std::string& ch) const {
    std::vector<int> out;
    if (ch == "X" || ch == "x" || ch == "-" || ch == "?") {
      for (unsigned s = 0; s < getSize(); ++s) out.push_back((int)s);
      return out;
    }
    const long v = std::strtol(ch.c_str(), nullptr, 10);
    if (v >= (long)min_ && v <= (long)max_) out.push_back((int)(v - min_));
    return out;
  }

 private:
  unsigned min_, max_;
};
namespace AlphabetTools {
inline const DNA& DNA_ALPHABET() { static DNA a; return a; }
inline const ProteicAlphabet& PROTEIN_ALPHABET() { static ProteicAlphabet a; return a; }
inline const CodonAlphabet& CODON_ALPHABET() { static CodonAlphabet a; return a; }
}  // namespace AlphabetTools

class BasicSequence {
 public:
  BasicSequence(const std::string& name, const std::string& content, const Alphabet* alpha) : name_(name), alpha_(alpha) {
    const unsigned w = alpha->getStateCodingSize();
    if (w == 0) {  // variable width: whitespace separated (chromosome counts)
      std::istringstream is(content);
      std::string tok;
      while (is >> tok) states_.push_back(tok);
    } else {
      for (size_t i = 0; i + w <= content.size(); i += w) states_.push_back(content.substr(i, w));
    }
  }
  BasicSequence(const std::string& name, const std::vector<std::string>& states, const Alphabet* alpha) : name_(name), states_(states), alpha_(alpha) {}
  const std::string& getName() const { return name_; }
  size_t size() const { return states_.size(); }
  const std::string& operator[](size_t i) const { return states_[i]; }
  const Alphabet* getAlphabet() const { return alpha_; }

 private:
  std::string name_;
  std::vector<std::string> states_;
  const Alphabet* alpha_;
};

class VectorSiteContainer {
 public:
  explicit VectorSiteContainer(const Alphabet* alpha) : alpha_(alpha) {}
  void addSequence(const BasicSequence& s) {
    if (!seqs_.empty() && s.size() != seqs_[0].size()) throw Exception("VectorSiteContainer::addSequence. Sequence " + s.getName() + " has a different length.");
    seqs_.push_back(s);
  }
  size_t getNumberOfSequences() const { return seqs_.size(); }
  size_t getNumberOfSites() const { return seqs_.empty() ? 0 : seqs_[0].size(); }
  const Alphabet* getAlphabet() const { return alpha_; }
  std::vector<std::string> getSequencesNames() const {
    std::vector<std::string> n;
    for (const BasicSequence& s : seqs_) n.push_back(s.getName());
    return n;
  }
  const BasicSequence& getSequence(const std::string& name) const {
    for (const BasicSequence& s : seqs_)
      if (s.getName() == name) return s;
    throw Exception("SequenceNotFoundException: " + name);
  }

 private:
  const Alphabet* alpha_;
  std::vector<BasicSequence> seqs_;
};

// ---- trees (TreeTemplate.h, TreeTemplateTools.h) ---------------------------------------------------------------------
class Node {
 public:
  Node() : id_(-1), father_(nullptr), hasLen_(false), len_(0) {}
  ~Node() { for (Node* s : sons_) delete s; }
  int getId() const { return id_; }
  void setId(int i) { id_ = i; }
  bool isLeaf() const { return sons_.empty(); }
  bool hasFather() const { return father_ != nullptr; }
  Node* getFather() const { return father_; }
  size_t getNumberOfSons() const { return sons_.size(); }
  Node* getSon(size_t i) const { return sons_[i]; }
  void addSon(Node* s) { sons_.push_back(s); s->father_ = this; }
  bool hasDistanceToFather() const { return hasLen_; }
  double getDistanceToFather() const { return len_; }
  void setDistanceToFather(double d) { len_ = d; hasLen_ = true; }
  void deleteDistanceToFather() { hasLen_ = false; len_ = 0; }
  bool hasName() const { return !name_.empty(); }
  const std::string& getName() const { return name_; }
  void setName(const std::string& n) { name_ = n; }
  void removeFather() { father_ = nullptr; }
  std::vector<Node*>& sons() { return sons_; }
  Node* cloneSubtree() const {
    Node* n = new Node();
    n->id_ = id_; n->hasLen_ = hasLen_; n->len_ = len_; n->name_ = name_;
    for (Node* s : sons_) n->addSon(s->cloneSubtree());
    return n;
  }

 private:
  int id_;
  Node* father_;
  std::vector<Node*> sons_;
  bool hasLen_;
  double len_;
  std::string name_;
};

namespace TreeTemplateTools {
// post-order, root last (TreeTemplateTools.h:354-361)
inline void getNodes(Node* n, std::vector<Node*>& out) {
  for (size_t i = 0; i < n->getNumberOfSons(); ++i) getNodes(n->getSon(i), out);
  out.push_back(n);
}
// pre-order leaves (TreeTemplateTools.h:96-106)
inline void getLeaves(Node* n, std::vector<Node*>& out) {
  if (n->isLeaf()) out.push_back(n);
  for (size_t i = 0; i < n->getNumberOfSons(); ++i) getLeaves(n->getSon(i), out);
}
}  // namespace TreeTemplateTools

template <class N = Node>
class TreeTemplate {
 public:
  explicit TreeTemplate(N* root) : root_(root) { resetNodesId(); }
  TreeTemplate(const TreeTemplate& t) : root_(t.root_->cloneSubtree()) {}
  TreeTemplate& operator=(const TreeTemplate& t) { if (this != &t) { delete root_; root_ = t.root_->cloneSubtree(); } return *this; }
  ~TreeTemplate() { delete root_; }
  N* getRootNode() const { return root_; }
  bool isRooted() const { return root_->getNumberOfSons() == 2; }
  std::vector<N*> getNodes() const { std::vector<N*> v; TreeTemplateTools::getNodes(root_, v); return v; }
  std::vector<N*> getLeaves() const { std::vector<N*> v; TreeTemplateTools::getLeaves(root_, v); return v; }
  std::vector<std::string> getLeavesNames() const {
    std::vector<std::string> n;
    for (N* l : getLeaves()) n.push_back(l->getName());
    return n;
  }
  std::vector<int> getNodesId() const { std::vector<int> v; for (N* n : getNodes()) v.push_back(n->getId()); return v; }
  N* getNode(int id) const { for (N* n : getNodes()) if (n->getId() == id) return n; throw Exception("NodeNotFoundException: TreeTemplate::getNode(): Node with id not found."); }
  int getFatherId(int id) const { return getNode(id)->getFather()->getId(); }
  void resetNodesId() { int i = 0; for (N* n : getNodes()) n->setId(i++); }
  // TreeTemplate::unroot (TreeTemplate.h:244-284): keep son 0 as the new root, hang son 1 under it, sum the lengths
  bool unroot() {
    if (!isRooted()) throw Exception("UnrootedTreeException: Tree::unroot. Tree is already rooted.");
    N* s1 = root_->getSon(0);
    N* s2 = root_->getSon(1);
    if (s1->isLeaf() && s2->isLeaf()) return false;
    if (s1->isLeaf()) std::swap(s1, s2);
    if (s1->hasDistanceToFather()) {
      s2->setDistanceToFather(s2->hasDistanceToFather() ? s1->getDistanceToFather() + s2->getDistanceToFather() : s1->getDistanceToFather());
      s1->deleteDistanceToFather();
    }
    root_->sons().clear();
    delete root_;
    s1->removeFather();
    s1->addSon(s2);
    root_ = s1;
    return true;
  }

 private:
  N* root_;
};
typedef TreeTemplate<Node> Tree;

namespace TreeTemplateTools {
inline Node* parse_(const std::string& s, size_t& pos) {
  auto skip = [&]() { while (pos < s.size() && std::isspace((unsigned char)s[pos])) ++pos; };
  skip();
  Node* n = new Node();
  if (pos < s.size() && s[pos] == '(') {
    ++pos;
    for (;;) {
      n->addSon(parse_(s, pos));
      skip();
      if (pos < s.size() && s[pos] == ',') { ++pos; continue; }
      if (pos < s.size() && s[pos] == ')') { ++pos; break; }
      delete n;
      throw Exception("TreeTemplateTools::parenthesisToTree. Bad tree description: " + s);
    }
  }
  skip();
  size_t st = pos;
  while (pos < s.size() && std::string(",():;").find(s[pos]) == std::string::npos) ++pos;
  std::string nm = s.substr(st, pos - st);
  while (!nm.empty() && std::isspace((unsigned char)nm.back())) nm.pop_back();
  if (!nm.empty()) n->setName(nm);
  skip();
  if (pos < s.size() && s[pos] == ':') {
    ++pos;
    st = pos;
    while (pos < s.size() && std::string(",();").find(s[pos]) == std::string::npos) ++pos;
    n->setDistanceToFather(std::strtod(s.substr(st, pos - st).c_str(), nullptr));
  }
  return n;
}
inline TreeTemplate<Node>* parenthesisToTree(const std::string& description) {
  size_t pos = 0;
  return new TreeTemplate<Node>(parse_(description, pos));
}
}  // namespace TreeTemplateTools

// ---- rate distributions --------------------------------------------------------------------------------------------------
class DiscreteDistribution {
 public:
  virtual ~DiscreteDistribution() {}
  size_t getNumberOfCategories() const { return values_.size(); }
  double getCategory(size_t i) const { return values_[i]; }
  double getProbability(size_t i) const { return probs_[i]; }
  virtual void setParameterValue(const std::string& name, double v) { (void)name; (void)v; throw ParameterNotFoundException(name); }

 protected:
  Vdouble values_, probs_;
};
class ConstantRateDistribution : public DiscreteDistribution {
 public:
  ConstantRateDistribution() { values_.assign(1, 1.0); probs_.assign(1, 1.0); }
};
// Gamma(alpha, beta = alpha), K equiprobable classes, class value = class mean (bpp-core GammaDiscreteDistribution)
class GammaDiscreteRateDistribution : public DiscreteDistribution {
 public:
  GammaDiscreteRateDistribution(size_t n, double alpha = 1.0) : n_(n), alpha_(alpha) { discretize(); }
  void setParameterValue(const std::string& name, double v) {
    if (name != "alpha" && name != "Gamma.alpha") throw ParameterNotFoundException(name);
    alpha_ = v;
    discretize();
  }
  double getAlpha() const { return alpha_; }

 private:
  void discretize() {
    values_.assign(n_, 1.0);
    probs_.assign(n_, 1.0 / (double)n_);
    if (n_ == 1) return;
    const double beta = alpha_;
    std::vector<double> cdf1(n_ + 1, 0.0);
    cdf1[n_] = 1.0;
    for (size_t i = 1; i < n_; ++i) {
      const double q = linalg::gammp_inv(alpha_, (double)i / (double)n_) / beta;  // class bound
      cdf1[i] = linalg::gammp(alpha_ + 1.0, q * beta);
    }
    for (size_t i = 0; i < n_; ++i) values_[i] = (double)n_ * (alpha_ / beta) * (cdf1[i + 1] - cdf1[i]);
  }
  size_t n_;
  double alpha_;
};

// ---- substitution models -----------------------------------------------------------------------------------------------------
class SubstitutionModel {
 public:
  virtual ~SubstitutionModel() {}
  virtual std::string getName() const = 0;
  virtual const Alphabet* getAlphabet() const = 0;
  virtual size_t getNumberOfStates() const = 0;
  virtual const RowMatrix<double>& getPij_t(double t) const = 0;
  virtual const RowMatrix<double>& getdPij_dt(double t) const = 0;
  virtual const RowMatrix<double>& getd2Pij_dt2(double t) const = 0;
  virtual double Pij_t(size_t i, size_t j, double t) const { return getPij_t(t)(i, j); }
  virtual double dPij_dt(size_t i, size_t j, double t) const { return getdPij_dt(t)(i, j); }
  virtual double d2Pij_dt2(size_t i, size_t j, double t) const { return getd2Pij_dt2(t)(i, j); }
  virtual const RowMatrix<double>& getGenerator() const = 0;
  virtual const Vdouble& getEigenValues() const = 0;
  virtual const Vdouble& getIEigenValues() const = 0;
  virtual bool isDiagonalizable() const = 0;
  virtual bool isNonSingular() const = 0;
  virtual const RowMatrix<double>& getRowLeftEigenVectors() const = 0;
  virtual const RowMatrix<double>& getColumnRightEigenVectors() const = 0;
  virtual double getRate() const = 0;
  virtual void setRate(double r) = 0;
  virtual const Vdouble& getFrequencies() const = 0;
  virtual double getInitValue(size_t i, const std::string& ch) const = 0;
  virtual void setParameterValue(const std::string& name, double v) = 0;
  virtual std::vector<std::string> getParameterNames() const = 0;
  virtual void fillModelDesc(bppgpu_model_desc& d) const = 0;
  virtual SubstitutionModel* clone() const = 0;
  virtual double getParameterValue(const std::string& name) const { throw ParameterNotFoundException(name); }
};

class AbstractSubstitutionModel : public SubstitutionModel {
 public:
  AbstractSubstitutionModel(const Alphabet* alpha, size_t size)
      : alphabet_(alpha), size_(size), rate_(1.0), generator_(size, size), freq_(size, 1.0 / (double)size), eigenValues_(size),
        iEigenValues_(size, 0.0), isDiagonalizable_(false), isNonSingular_(false), isScalable_(true), reversible_(false),
        rightEigenVectors_(size, size), leftEigenVectors_(size, size), device_(0), extraFlags_(0), pijt_(size, size),
        dpijt_(size, size), d2pijt_(size, size) {}
  const Alphabet* getAlphabet() const { return alphabet_; }
  size_t getNumberOfStates() const { return size_; }
  const RowMatrix<double>& getGenerator() const { return generator_; }
  const Vdouble& getEigenValues() const { return eigenValues_; }
  const Vdouble& getIEigenValues() const { return iEigenValues_; }
  bool isDiagonalizable() const { return isDiagonalizable_; }
  bool isNonSingular() const { return isNonSingular_; }
  const RowMatrix<double>& getRowLeftEigenVectors() const { return leftEigenVectors_; }
  const RowMatrix<double>& getColumnRightEigenVectors() const { return rightEigenVectors_; }
  double getRate() const { return rate_; }
  void setRate(double r) { if (r <= 0) throw Exception("Bad value for rate: " + std::to_string(r)); rate_ = r; }
  const Vdouble& getFrequencies() const { return freq_; }
  void setDevice(int d) { device_ = d; }

  // AbstractTransitionModel::getInitValue (Model/AbstractSubstitutionModel.cpp:98-112)
  double getInitValue(size_t i, const std::string& ch) const {
    if (i >= size_) throw Exception("IndexOutOfBoundsException: AbstractTransitionModel::getInitValue");
    const std::vector<int> states = alphabet_->getAlias(ch);
    if (states.empty()) throw Exception("BadIntException: AbstractTransitionModel::getInitValue. Character " + ch + " is not allowed in model.");
    for (int s : states)
      if ((int)i == s) return 1.0;
    return 0.0;
  }

  // getPij_t / getdPij_dt / getd2Pij_dt2 (Model/AbstractSubstitutionModel.cpp:426-641): Interface 1 of the C ABI.
  // Like the reference the returned reference is to an internal buffer, valid until the next call.
  const RowMatrix<double>& getPij_t(double t) const { return ptable_(t, BPPGPU_WANT_P, pijt_); }
  const RowMatrix<double>& getdPij_dt(double t) const { return ptable_(t, BPPGPU_WANT_DP, dpijt_); }
  const RowMatrix<double>& getd2Pij_dt2(double t) const { return ptable_(t, BPPGPU_WANT_D2P, d2pijt_); }

  void fillModelDesc(bppgpu_model_desc& d) const {
    d.n_states = (int32_t)size_;
    d.flags = (isDiagonalizable_ ? BPPGPU_MODEL_DIAGONALIZABLE : 0u) | (isNonSingular_ ? BPPGPU_MODEL_NONSINGULAR : 0u) | extraFlags_;
    d.rate = rate_;
    d.right_eigen = rightEigenVectors_.data();
    d.left_eigen = leftEigenVectors_.data();
    d.eigen_re = eigenValues_.data();
    d.eigen_im = iEigenValues_.data();
    d.generator = generator_.data();
    d.taylor_epsilon = 1e-4;
  }

 protected:
  // AbstractSubstitutionModel::updateMatrices (Model/AbstractSubstitutionModel.cpp:175-421): null ("stop") lines are
  // stripped, the rest eigen-decomposed on the host, the ~0 eigenvalue pinned to 0 and (optionally) turned into the
  // equilibrium frequencies, then the generator is normalised to one substitution per unit time if scalable.
  void updateMatrices(bool computeFreq) {
    const int n = (int)size_;
    std::vector<char> vnull(n, 0);
    std::vector<int> ok;
    for (int i = 0; i < n; ++i) {
      bool null_ = std::fabs(generator_(i, i)) < NumConstants::TINY();
      for (int j = 0; null_ && j < n; ++j)
        if (std::fabs(generator_(j, i)) >= NumConstants::TINY()) null_ = false;
      vnull[i] = null_;
      if (!null_) ok.push_back(i);
    }
    const int m = (int)ok.size();
    std::vector<double> A((size_t)m * m), re, im, Vk;
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < m; ++j) A[(size_t)i * m + j] = generator_(ok[i], ok[j]);
    bool eig_ok;
    if (reversible_) {
      // pi^1/2 Q pi^-1/2 is symmetric for a reversible generator
      std::vector<double> B((size_t)m * m), w, U;
      std::vector<double> sq(m);
      for (int i = 0; i < m; ++i) sq[i] = std::sqrt(freq_[ok[i]]);
      for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) B[(size_t)i * m + j] = 0.5 * (sq[i] * A[(size_t)i * m + j] / sq[j] + sq[j] * A[(size_t)j * m + i] / sq[i]);
      linalg::jacobi_symmetric(B, m, w, U);
      re = w;
      im.assign(m, 0.0);
      Vk.assign((size_t)m * m, 0.0);
      for (int i = 0; i < m; ++i)
        for (int k = 0; k < m; ++k) Vk[(size_t)i * m + k] = U[(size_t)i * m + k] / sq[i];
      eig_ok = true;
    } else {
      eig_ok = linalg::eigen_general(A, m, re, im, Vk);
    }
    std::vector<double> V((size_t)n * n, 0.0), Vinv;
    eigenValues_.assign(n, 0.0);
    iEigenValues_.assign(n, 0.0);
    isNonSingular_ = false;
    isDiagonalizable_ = false;
    if (eig_ok) {
      for (int k = 0; k < m; ++k) { eigenValues_[k] = re[k]; iEigenValues_[k] = im[k]; }
      for (int i = 0; i < m; ++i)
        for (int k = 0; k < m; ++k) V[(size_t)ok[i] * n + k] = Vk[(size_t)i * m + k];
      int gi = 0;
      for (int i = 0; i < n; ++i)
        if (vnull[i]) V[(size_t)i * n + m + gi++] = 1.0;
      bool usable = linalg::invert(V, n, Vinv);
      if (usable) {
        // the eigen form must reproduce the generator; a (nearly) defective matrix does not and takes the series path,
        // as the reference does when MatrixTools::inv fails (:283-291)
        std::vector<double> D((size_t)n * n, 0.0);
        for (int k = 0; k < n; ++k) {
          D[(size_t)k * n + k] = eigenValues_[k];
          if (iEigenValues_[k] > 0 && k + 1 < n) { D[(size_t)k * n + k + 1] = iEigenValues_[k]; D[(size_t)(k + 1) * n + k] = -iEigenValues_[k]; }
        }
        const std::vector<double> R = linalg::matmul(linalg::matmul(V, D, n), Vinv, n);
        double err = 0.0, nrm = 0.0;
        for (int i = 0; i < n; ++i)
          for (int j = 0; j < n; ++j) {
            err = std::max(err, std::fabs(R[(size_t)i * n + j] - generator_(i, j)));
            nrm = std::max(nrm, std::fabs(generator_(i, j)));
          }
        if (!(err <= 1e-8 * std::max(nrm, 1e-300))) usable = false;
      }
      if (usable) {
        isDiagonalizable_ = true;
        if (!reversible_)
          for (int k = 0; k < n; ++k)
            if (std::fabs(iEigenValues_[k]) > NumConstants::TINY()) isDiagonalizable_ = false;
        // the unique ~0 eigenvalue (tolerance ladder, :306-315)
        std::vector<int> nullev;
        double fact = 0.1;
        while (nullev.empty() && fact < 1000) {
          fact *= 10;
          for (int k = 0; k < m; ++k)
            if (std::fabs(eigenValues_[k]) < fact * NumConstants::SMALL() && std::fabs(iEigenValues_[k]) < NumConstants::SMALL()) nullev.push_back(k);
        }
        int nulleigen = -1;
        if (nullev.size() == 1) nulleigen = nullev[0];
        else
          for (int cand : nullev) {  // :326-352: the one whose right vector is constant
            const double val = V[(size_t)ok[0] * n + cand];
            bool cst = val != 0.0;
            for (int i = 1; cst && i < m; ++i)
              if (std::fabs((V[(size_t)ok[i] * n + cand] - val) / val) > NumConstants::SMALL()) cst = false;
            if (cst) { nulleigen = cand; break; }
          }
        if (nulleigen >= 0) {
          isNonSingular_ = true;
          eigenValues_[nulleigen] = 0.0;
          iEigenValues_[nulleigen] = 0.0;
          if (computeFreq) {
            double sum = 0.0;
            for (int j = 0; j < n; ++j) sum += Vinv[(size_t)nulleigen * n + j];
            for (int j = 0; j < n; ++j) freq_[j] = Vinv[(size_t)nulleigen * n + j] / sum;
          }
        } else {
          isDiagonalizable_ = false;
        }
      }
    }
    if (!isNonSingular_) {
      // :386-410: rescale so the fastest state leaves at rate 1, frequencies from (I + Q)^256
      double mn = 0.0;
      for (int i = 0; i < n; ++i) mn = std::min(mn, generator_(i, i));
      if (isScalable_ && mn < 0) scaleGenerator(-1.0 / mn);
      if (computeFreq) {
        std::vector<double> T((size_t)n * n);
        for (int i = 0; i < n; ++i)
          for (int j = 0; j < n; ++j) T[(size_t)i * n + j] = generator_(i, j) + (i == j ? 1.0 : 0.0);
        for (int k = 0; k < 8; ++k) T = linalg::matmul(T, T, n);
        for (int j = 0; j < n; ++j) freq_[j] = T[j];
      }
      Vinv.assign((size_t)n * n, 0.0);
    }
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        rightEigenVectors_(i, j) = V[(size_t)i * n + j];
        leftEigenVectors_(i, j) = Vinv.empty() ? 0.0 : Vinv[(size_t)i * n + j];
      }
    if (isScalable_) {  // normalize(): -sum_i pi_i Q_ii = 1  (:645-652, :684-688)
      double sc = 0.0;
      for (int i = 0; i < n; ++i) sc -= freq_[i] * generator_(i, i);
      if (sc > 0) scaleGenerator(1.0 / sc);
    }
  }
  void scaleGenerator(double s) {
    for (size_t i = 0; i < size_; ++i) {
      for (size_t j = 0; j < size_; ++j) generator_(i, j) *= s;
      eigenValues_[i] *= s;
      iEigenValues_[i] *= s;
    }
  }
  void setDiagonal() {
    for (size_t i = 0; i < size_; ++i) {
      double s = 0.0;
      for (size_t j = 0; j < size_; ++j)
        if (j != i) s += generator_(i, j);
      generator_(i, i) = -s;
    }
  }

  const Alphabet* alphabet_;
  size_t size_;
  double rate_;
  RowMatrix<double> generator_;
  Vdouble freq_, eigenValues_, iEigenValues_;
  bool isDiagonalizable_, isNonSingular_, isScalable_, reversible_;
  RowMatrix<double> rightEigenVectors_, leftEigenVectors_;
  int device_;
  unsigned extraFlags_;

 private:
  const RowMatrix<double>& ptable_(double t, unsigned which, RowMatrix<double>& out) const {
    bppgpu_model_desc d;
    fillModelDesc(d);
    check(bppgpu_pt_batch(device_, &d, 1, &t, which, which == BPPGPU_WANT_P ? out.data() : nullptr,
                          which == BPPGPU_WANT_DP ? out.data() : nullptr, which == BPPGPU_WANT_D2P ? out.data() : nullptr),
          "getPij_t");
    return out;
  }
  mutable RowMatrix<double> pijt_, dpijt_, d2pijt_;
};

// AbstractReversibleSubstitutionModel::updateMatrices (:694-703): generator = exchangeability * frequencies
class AbstractReversibleSubstitutionModel : public AbstractSubstitutionModel {
 public:
  AbstractReversibleSubstitutionModel(const Alphabet* a, size_t n) : AbstractSubstitutionModel(a, n), exch_(n, n) { reversible_ = true; }

 protected:
  void updateReversible() {
    for (size_t i = 0; i < size_; ++i)
      for (size_t j = 0; j < size_; ++j) generator_(i, j) = i == j ? 0.0 : exch_(i, j) * freq_[j];
    setDiagonal();
    double sc = 0.0;
    for (size_t i = 0; i < size_; ++i) sc -= freq_[i] * generator_(i, i);
    scaleGenerator(1.0 / sc);
    const Vdouble keep = freq_;
    updateMatrices(false);
    freq_ = keep;
  }
  RowMatrix<double> exch_;
};

// Model/Nucleotide/GTR.cpp:84-124: exchangeabilities AC=d AG=1 AT=b CG=e CT=a GT=c; theta, theta1, theta2 frequencies
class GTR : public AbstractReversibleSubstitutionModel {
 public:
  GTR(const Alphabet* alpha, double a = 1., double b = 1., double c = 1., double d = 1., double e = 1., double piA = 0.25,
      double piC = 0.25, double piG = 0.25, double piT = 0.25)
      : AbstractReversibleSubstitutionModel(alpha, 4), a_(a), b_(b), c_(c), d_(d), e_(e) {
    freq_[0] = piA; freq_[1] = piC; freq_[2] = piG; freq_[3] = piT;
    update();
  }
  GTR* clone() const { return new GTR(*this); }
  std::string getName() const { return "GTR"; }
  std::vector<std::string> getParameterNames() const { return {"GTR.a", "GTR.b", "GTR.c", "GTR.d", "GTR.e"}; }
  void setParameterValue(const std::string& name, double v) {
    const std::string n = name.substr(name.find('.') == std::string::npos ? 0 : name.find('.') + 1);
    if (n == "a") a_ = v; else if (n == "b") b_ = v; else if (n == "c") c_ = v; else if (n == "d") d_ = v; else if (n == "e") e_ = v;
    else throw ParameterNotFoundException(name);
    update();
  }

 private:
  void update() {
    exch_.resize(4, 4);
    exch_(0, 1) = exch_(1, 0) = d_; exch_(0, 2) = exch_(2, 0) = 1.0; exch_(0, 3) = exch_(3, 0) = b_;
    exch_(1, 2) = exch_(2, 1) = e_; exch_(1, 3) = exch_(3, 1) = a_; exch_(2, 3) = exch_(3, 2) = c_;
    updateReversible();
  }
  double a_, b_, c_, d_, e_;
};

// Model/Nucleotide/HKY85.cpp:80-191 (generic eigen path instead of the closed form: identical P)
class HKY85 : public AbstractReversibleSubstitutionModel {
 public:
  HKY85(const Alphabet* alpha, double kappa = 1., double piA = 0.25, double piC = 0.25, double piG = 0.25, double piT = 0.25,
        const std::string& name = "HKY85")
      : AbstractReversibleSubstitutionModel(alpha, 4), kappa_(kappa), name_(name) {
    freq_[0] = piA; freq_[1] = piC; freq_[2] = piG; freq_[3] = piT;
    update();
  }
  HKY85* clone() const { return new HKY85(*this); }
  std::string getName() const { return name_; }
  std::vector<std::string> getParameterNames() const { return {name_ + ".kappa"}; }
  double getParameterValue(const std::string& name) const {
    if (name == "kappa" || name == name_ + ".kappa") return kappa_;
    throw ParameterNotFoundException(name);
  }
  void setParameterValue(const std::string& name, double v) {
    if (name != "kappa" && name != name_ + ".kappa") throw ParameterNotFoundException(name);
    kappa_ = v;
    update();
  }
  double getKappa() const { return kappa_; }

 protected:
  void update() {
    exch_.resize(4, 4);
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) exch_(i, j) = i == j ? 0.0 : 1.0;
    exch_(0, 2) = exch_(2, 0) = kappa_;
    exch_(1, 3) = exch_(3, 1) = kappa_;
    updateReversible();
  }
  double kappa_;
  std::string name_;
};
// Model/Nucleotide/T92.cpp:81-187: HKY85 with pi = ((1-theta)/2, theta/2, theta/2, (1-theta)/2)
class T92 : public HKY85 {
 public:
  T92(const Alphabet* alpha, double kappa = 1., double theta = 0.5)
      : HKY85(alpha, kappa, (1 - theta) / 2, theta / 2, theta / 2, (1 - theta) / 2, "T92"), theta_(theta) {}
  T92* clone() const { return new T92(*this); }
  std::vector<std::string> getParameterNames() const { return {"T92.kappa", "T92.theta"}; }
  double getParameterValue(const std::string& name) const {
    if (name == "theta" || name == "T92.theta") return theta_;
    return HKY85::getParameterValue(name);
  }
  void setParameterValue(const std::string& name, double v) {
    if (name == "theta" || name == "T92.theta") {   // T92::updateMatrices (T92.cpp:81-94): piA = piT = (1 - theta) / 2, piC = piG = theta / 2
      theta_ = v;
      freq_[0] = freq_[3] = (1 - v) / 2;
      freq_[1] = freq_[2] = v / 2;
      update();
    } else HKY85::setParameterValue(name, v);
  }

 private:
  double theta_;
};
class K80 : public HKY85 {
 public:
  K80(const Alphabet* alpha, double kappa = 1.) : HKY85(alpha, kappa, .25, .25, .25, .25, "K80") {}
  K80* clone() const { return new K80(*this); }
};
class JCnuc : public HKY85 {
 public:
  explicit JCnuc(const Alphabet* alpha) : HKY85(alpha, 1.0, .25, .25, .25, .25, "JC69") {}
  JCnuc* clone() const { return new JCnuc(*this); }
};

// Model/Protein/LG08.cpp:53-62 + the published Le & Gascuel 2008 constants
#include "lg08_data.inc"
class LG08 : public AbstractReversibleSubstitutionModel {
 public:
  explicit LG08(const Alphabet* alpha) : AbstractReversibleSubstitutionModel(alpha, 20) {
    int k = 0;
    for (int i = 1; i < 20; ++i)
      for (int j = 0; j < i; ++j) exch_(i, j) = exch_(j, i) = LG08_LOWER[k++];
    for (int i = 0; i < 20; ++i) freq_[i] = LG08_FREQ[i];
    updateReversible();
  }
  LG08* clone() const { return new LG08(*this); }
  std::string getName() const { return "LG08"; }
  std::vector<std::string> getParameterNames() const { return {}; }
  void setParameterValue(const std::string& name, double) { throw ParameterNotFoundException(name); }
};

// Model/Codon/YN98.cpp:51-77: K80 rate / 3 on single-nucleotide changes, x omega if non-synonymous, x target codon
// frequency, stop codons zeroed (AbstractCodonSubstitutionModel.cpp:174-191), then normalised
class YN98 : public AbstractSubstitutionModel {
 public:
  YN98(const Alphabet* alpha, double kappa = 1., double omega = 1., const Vdouble* codonFreq = nullptr)
      : AbstractSubstitutionModel(alpha, 64), kappa_(kappa), omega_(omega) {
    if (codonFreq) freq_ = *codonFreq;
    else {
      double n = 0;
      for (int i = 0; i < 64; ++i) { freq_[i] = CodonAlphabet::isStop(i) ? 0.0 : 1.0; n += freq_[i]; }
      for (int i = 0; i < 64; ++i) freq_[i] /= n;
    }
    reversible_ = false;  // the reference runs the general EigenValue path for word models
    update();
  }
  YN98* clone() const { return new YN98(*this); }
  std::string getName() const { return "YN98"; }
  double getParameterValue(const std::string& name) const {
    if (name == "kappa" || name == "YN98.kappa") return kappa_;
    if (name == "omega" || name == "YN98.omega") return omega_;
    throw ParameterNotFoundException(name);
  }
  std::vector<std::string> getParameterNames() const { return {"YN98.kappa", "YN98.omega"}; }
  void setParameterValue(const std::string& name, double v) {
    if (name == "kappa" || name == "YN98.kappa") kappa_ = v;
    else if (name == "omega" || name == "YN98.omega") omega_ = v;
    else throw ParameterNotFoundException(name);
    update();
  }

 private:
  void update() {
    for (int i = 0; i < 64; ++i)
      for (int j = 0; j < 64; ++j) {
        generator_(i, j) = 0.0;
        if (i == j) continue;
        const int di[3] = {i / 16, (i / 4) % 4, i % 4}, dj[3] = {j / 16, (j / 4) % 4, j % 4};
        int ndiff = 0, p = -1;
        for (int k = 0; k < 3; ++k)
          if (di[k] != dj[k]) { ++ndiff; p = k; }
        if (ndiff != 1 || CodonAlphabet::isStop(i) || CodonAlphabet::isStop(j)) continue;
        const bool ts = (di[p] == 0 && dj[p] == 2) || (di[p] == 2 && dj[p] == 0) || (di[p] == 1 && dj[p] == 3) || (di[p] == 3 && dj[p] == 1);
        double r = (ts ? kappa_ : 1.0) / (kappa_ + 2.0) / 3.0;
        r *= (CodonAlphabet::aminoAcid(i) == CodonAlphabet::aminoAcid(j) ? 1.0 : omega_) * freq_[j];
        generator_(i, j) = r;
      }
    setDiagonal();
    const Vdouble keep = freq_;
    // reversible w.r.t. the codon frequencies: use the symmetric solver on the sense codons
    reversible_ = true;
    updateMatrices(false);
    reversible_ = false;
    freq_ = keep;
  }
  double kappa_, omega_;
};

// Model/ChromosomeSubstitutionModel.cpp:431-802 (gain / loss / duplication / demi-duplication / base number)
class ChromosomeSubstitutionModel : public AbstractSubstitutionModel {
 public:
  static constexpr double IgnoreParam = -999.0;   // ChromosomeSubstitutionModel.h:15-23
  static constexpr double DemiEqualDupl = -2.0;
  enum rateChangeFunc { LINEAR = 0, EXP = 1 };
  ChromosomeSubstitutionModel(const ChromosomeAlphabet* alpha, double gain, double loss, double dupl, double demi,
                              double gainR = IgnoreParam, double lossR = IgnoreParam, double duplR = IgnoreParam,
                              int baseNum = (int)IgnoreParam, double baseNumR = IgnoreParam, unsigned maxChrRange = 0,
                              rateChangeFunc rc = LINEAR)
      : AbstractSubstitutionModel(alpha, alpha->getSize()), chr_(alpha), gain_(gain), loss_(loss), dupl_(dupl), demi_(demi),
        gainR_(gainR), lossR_(lossR), duplR_(duplR), baseNum_(baseNum), baseNumR_(baseNumR), maxChrRange_(maxChrRange), rc_(rc) {
    isScalable_ = false;  // :58
    extraFlags_ = BPPGPU_MODEL_CLAMP01 | BPPGPU_MODEL_CHR_DERIV | BPPGPU_MODEL_CHR_TAYLOR;
    update();
  }
  ChromosomeSubstitutionModel* clone() const { return new ChromosomeSubstitutionModel(*this); }
  std::string getName() const { return "Chromosome"; }
  std::vector<std::string> getParameterNames() const { return {"Chromosome.gain", "Chromosome.loss", "Chromosome.dupl", "Chromosome.demi"}; }
  double getParameterValue(const std::string& name) const {
    const std::string n = name.substr(name.find('.') == std::string::npos ? 0 : name.find('.') + 1);
    const std::map<std::string, double> m = {{"gain", gain_}, {"loss", loss_}, {"dupl", dupl_}, {"demi", demi_}, {"gainR", gainR_},
                                             {"lossR", lossR_}, {"duplR", duplR_}, {"baseNumR", baseNumR_}};
    if (!m.count(n)) throw ParameterNotFoundException(name);
    return m.at(n);
  }
  void setParameterValue(const std::string& name, double v) {
    const std::string n = name.substr(name.find('.') == std::string::npos ? 0 : name.find('.') + 1);
    if (n == "gain") gain_ = v; else if (n == "loss") loss_ = v; else if (n == "dupl") dupl_ = v; else if (n == "demi") demi_ = v;
    else if (n == "gainR") gainR_ = v; else if (n == "lossR") lossR_ = v; else if (n == "duplR") duplR_ = v; else if (n == "baseNumR") baseNumR_ = v;
    else throw ParameterNotFoundException(name);
    update();
  }

 private:
  double rate_(int i, double c, double lin) const {  // getRate (:504-526)
    if (c == IgnoreParam && lin == IgnoreParam) return 0.0;
    const double total = c == IgnoreParam ? lin : c;
    if (lin == IgnoreParam) return total;
    return rc_ == LINEAR ? total + lin * (i - 1) : total * std::exp(lin * (i - 1));
  }
  void update() {
    const int mn = (int)chr_->getMin(), mx = (int)chr_->getMax(), n = (int)size_;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) generator_(i, j) = 0.0;
    const double demi = demi_ == DemiEqualDupl ? dupl_ : demi_;
    for (int i = mn; i <= mx; ++i) {
      const int r = i - mn;
      if (i + 1 <= mx) generator_(r, r + 1) += rate_(i, gain_, gainR_);
      if (i - 1 >= mn) generator_(r, r - 1) += rate_(i, loss_, lossR_);
      if (2 * i <= mx) generator_(r, 2 * i - mn) += rate_(i, dupl_, duplR_);
      else if (i != mx) generator_(r, mx - mn) += rate_(i, dupl_, duplR_);
      if (demi != IgnoreParam && i != mx) {  // :533-560
        if (i % 2 == 0 && (int)(i * 1.5) <= mx) generator_(r, (int)(i * 1.5) - mn) += demi;
        else if (i % 2 != 0 && (int)std::ceil(i * 1.5) <= mx) {
          if (i == 1) generator_(r, (int)std::ceil(i * 1.5) - mn) += demi;
          else {
            generator_(r, (int)std::ceil(i * 1.5) - mn) += demi / 2;
            generator_(r, (int)std::floor(i * 1.5) - mn) += demi / 2;
          }
        } else generator_(r, mx - mn) += demi;
      }
      if (i < mx && baseNum_ != (int)IgnoreParam)  // :562-577
        for (int j = i + 1; j <= mx; ++j) {
          if (j == mx) { if ((unsigned)(j - i) <= maxChrRange_) generator_(r, j - mn) += baseNumR_; }
          else if ((j - i) % baseNum_ == 0 && (unsigned)(j - i) <= maxChrRange_) generator_(r, j - mn) += baseNumR_;
        }
    }
    setDiagonal();
    updateMatrices(false);
    for (size_t i = 0; i < size_; ++i) freq_[i] = 1.0 / (double)size_;
  }
  const ChromosomeAlphabet* chr_;
  double gain_, loss_, dupl_, demi_, gainR_, lossR_, duplR_;
  int baseNum_;
  double baseNumR_;
  unsigned maxChrRange_;
  rateChangeFunc rc_;
};

// ---- mixtures of one model over a parameter (Model/MixedSubstitutionModel.h, MixtureOfASubstitutionModel) ---------------------------
class MixedSubstitutionModel {
 public:
  virtual ~MixedSubstitutionModel() {}
  virtual std::string getName() const = 0;
  virtual size_t getNumberOfModels() const = 0;
  virtual SubstitutionModel* getNModel(size_t i) const = 0;
  virtual double getNProbability(size_t i) const = 0;
  virtual std::vector<std::string> getParameterNames() const = 0;
  virtual double getParameterValue(const std::string& name) const = 0;
  virtual void setParameterValue(const std::string& name, double v) = 0;
  Vdouble getProbabilities() const {
    Vdouble p(getNumberOfModels());
    for (size_t i = 0; i < p.size(); ++i) p[i] = getNProbability(i);
    return p;
  }
};

// The omega mixtures of the YNGP wrappers: three YN98 that differ in omega, class probabilities from the simplex parameters
// theta1, theta2 (SimpleDiscreteDistribution: p0 = theta1, p1 = (1 - theta1) theta2, p2 = the rest), and the homogenisation of the
// synonymous rate (YNGP_M2::updateMatrices, Model/Codon/YNGP_M2.cpp:134-146): sub-model k gets the relative rate
// 1 / Q_k(synfrom, synto) for the first synonymous pair with non-zero rates, normalised to mean 1 under the probabilities
// (MixtureOfASubstitutionModel::setVRates).
class OmegaMixture_ : public MixedSubstitutionModel {
 public:
  size_t getNumberOfModels() const { return 3; }
  SubstitutionModel* getNModel(size_t i) const { return sub_.at(i).get(); }
  double getNProbability(size_t i) const { return probs_.at(i); }

 protected:
  OmegaMixture_(const Alphabet* alpha, const Vdouble* codonFreq) : alpha_(alpha) { if (codonFreq) codonFreq_ = *codonFreq; }
  void rebuild(double kappa, const double omega[3], double theta1, double theta2) {
    probs_ = {theta1, (1 - theta1) * theta2, (1 - theta1) * (1 - theta2)};
    sub_.clear();
    for (int k = 0; k < 3; ++k) sub_.emplace_back(new YN98(alpha_, kappa, omega[k], codonFreq_.empty() ? nullptr : &codonFreq_));
    size_t from = 0, to = 0;
    bool found = false;
    for (size_t f = 1; f < 64 && !found; ++f)
      for (size_t t = 0; t < f && !found; ++t)
        if (CodonAlphabet::aminoAcid((int)f) == CodonAlphabet::aminoAcid((int)t) && !CodonAlphabet::isStop((int)f) &&
            sub_[0]->getGenerator()(f, t) != 0 && sub_[1]->getGenerator()(f, t) != 0) {
          from = f;
          to = t;
          found = true;
        }
    if (!found) throw Exception("Impossible to find synonymous codons");
    double r[3], mean = 0;
    for (int k = 0; k < 3; ++k) { r[k] = 1.0 / sub_[k]->getGenerator()(from, to); mean += probs_[k] * r[k]; }
    for (int k = 0; k < 3; ++k) sub_[k]->setRate(r[k] / mean);
  }
  const Alphabet* alpha_;
  Vdouble codonFreq_;
  std::vector<std::unique_ptr<YN98> > sub_;
  Vdouble probs_;
};
// Model/Codon/YNGP_M2.cpp:52-146: omega in {omega0 < 1, 1, omega2 > 1}
class YNGP_M2 : public OmegaMixture_ {
 public:
  YNGP_M2(const Alphabet* alpha, double kappa = 1., double omega0 = 0.5, double omega2 = 2., double theta1 = 0.333333, double theta2 = 0.5,
          const Vdouble* codonFreq = nullptr)
      : OmegaMixture_(alpha, codonFreq), kappa_(kappa), omega0_(omega0), omega2_(omega2), theta1_(theta1), theta2_(theta2) { update(); }
  std::string getName() const { return "YNGP_M2"; }
  std::vector<std::string> getParameterNames() const { return {"YNGP_M2.kappa", "YNGP_M2.omega0", "YNGP_M2.omega2", "YNGP_M2.theta1", "YNGP_M2.theta2"}; }
  double getParameterValue(const std::string& name) const { return *slot(name); }
  void setParameterValue(const std::string& name, double v) { *const_cast<double*>(slot(name)) = v; update(); }

 private:
  const double* slot(const std::string& name) const {
    const std::string n = name.substr(name.find('.') == std::string::npos ? 0 : name.find('.') + 1);
    const std::map<std::string, const double*> m = {{"kappa", &kappa_}, {"omega0", &omega0_}, {"omega2", &omega2_}, {"theta1", &theta1_}, {"theta2", &theta2_}};
    if (!m.count(n)) throw ParameterNotFoundException(name);
    return m.at(n);
  }
  void update() { const double w[3] = {omega0_, 1.0, omega2_}; rebuild(kappa_, w, theta1_, theta2_); }
  double kappa_, omega0_, omega2_, theta1_, theta2_;
};
// fork, Model/Codon/RELAX.cpp:52-218: omegas ((p omega1)^k, omega1^k, omega2^k), floored at 0.001 / capped at 999 (:176-205)
class RELAX : public OmegaMixture_ {
 public:
  RELAX(const Alphabet* alpha, double kappa = 1., double p = 0.5, double omega1 = 1., double omega2 = 2., double k = 1.,
        double theta1 = 0.333333, double theta2 = 0.5, const Vdouble* codonFreq = nullptr)
      : OmegaMixture_(alpha, codonFreq), kappa_(kappa), p_(p), omega1_(omega1), omega2_(omega2), k_(k), theta1_(theta1), theta2_(theta2) { update(); }
  std::string getName() const { return "RELAX"; }
  std::vector<std::string> getParameterNames() const {
    return {"RELAX.kappa", "RELAX.p", "RELAX.omega1", "RELAX.omega2", "RELAX.k", "RELAX.theta1", "RELAX.theta2"};
  }
  double getParameterValue(const std::string& name) const { return *slot(name); }
  void setParameterValue(const std::string& name, double v) { *const_cast<double*>(slot(name)) = v; update(); }

 private:
  const double* slot(const std::string& name) const {
    const std::string n = name.substr(name.find('.') == std::string::npos ? 0 : name.find('.') + 1);
    const std::map<std::string, const double*> m = {{"kappa", &kappa_}, {"p", &p_}, {"omega1", &omega1_}, {"omega2", &omega2_},
                                                     {"k", &k_}, {"theta1", &theta1_}, {"theta2", &theta2_}};
    if (!m.count(n)) throw ParameterNotFoundException(name);
    return m.at(n);
  }
  void update() {
    const double w[3] = {std::max(std::pow(p_ * omega1_, k_), 0.001), std::max(std::pow(omega1_, k_), 0.001), std::min(std::pow(omega2_, k_), 999.0)};
    rebuild(kappa_, w, theta1_, theta2_);
  }
  double kappa_, p_, omega1_, omega2_, k_, theta1_, theta2_;
};

// Model/MixedSubstitutionModelSet.h: mixed models on groups of branches; the site paths ("hyper-nodes") supported here are the
// ones test/test_relax.cpp:100-102 sets up -- sub-model k of every model travels together, with the first model's probability.
class MixedSubstitutionModelSet {
 public:
  explicit MixedSubstitutionModelSet(const Alphabet* alpha) : alphabet_(alpha) {}
  void addModel(MixedSubstitutionModel* model, const std::vector<int>& nodesId) {   // owns the model
    if (!models_.empty() && model->getNumberOfModels() != models_[0]->getNumberOfModels()) {
      delete model;
      throw Exception("MixedSubstitutionModelSet: every model needs the same number of sub-models for linked site paths");
    }
    for (int id : nodesId) nodeToModel_[id] = models_.size();
    models_.emplace_back(model);
  }
  size_t getNumberOfModels() const { return models_.size(); }
  MixedSubstitutionModel* getModel(size_t i) const { return models_.at(i).get(); }
  size_t getNumberOfPaths() const { return models_.at(0)->getNumberOfModels(); }
  double getPathProbability(size_t k) const { return models_.at(0)->getNProbability(k); }
  size_t getModelIndexForNode(int nodeId) const {
    std::map<int, size_t>::const_iterator it = nodeToModel_.find(nodeId);
    if (it == nodeToModel_.end()) throw Exception("MixedSubstitutionModelSet: no model associated to node with id " + std::to_string(nodeId));
    return it->second;
  }
  const Alphabet* getAlphabet() const { return alphabet_; }

 private:
  const Alphabet* alphabet_;
  std::vector<std::unique_ptr<MixedSubstitutionModel> > models_;
  std::map<int, size_t> nodeToModel_;
};

// ---- root frequency sets and non-homogeneous model sets ---------------------------------------------------------------------------
// Model/FrequencySet/NucleotideFrequencySet.h: GCFrequencySet (one parameter theta = G+C content), FixedFrequencySet
class FrequencySet {
 public:
  virtual ~FrequencySet() {}
  virtual FrequencySet* clone() const = 0;
  virtual const Vdouble& getFrequencies() const = 0;
  virtual std::vector<std::string> getParameterNames() const = 0;
  virtual double getParameterValue(const std::string& name) const = 0;
  virtual void setParameterValue(const std::string& name, double v) = 0;
};
class GCFrequencySet : public FrequencySet {
 public:
  explicit GCFrequencySet(const Alphabet* = nullptr, double theta = 0.5) : freq_(4) { set(theta); }
  GCFrequencySet* clone() const { return new GCFrequencySet(*this); }
  const Vdouble& getFrequencies() const { return freq_; }
  std::vector<std::string> getParameterNames() const { return {"GC.theta"}; }
  double getParameterValue(const std::string& name) const { if (name != "GC.theta" && name != "theta") throw ParameterNotFoundException(name); return theta_; }
  void setParameterValue(const std::string& name, double v) { if (name != "GC.theta" && name != "theta") throw ParameterNotFoundException(name); set(v); }

 private:
  void set(double theta) { theta_ = theta; freq_[0] = freq_[3] = (1 - theta) / 2; freq_[1] = freq_[2] = theta / 2; }
  double theta_;
  Vdouble freq_;
};
class FixedFrequencySet : public FrequencySet {
 public:
  explicit FixedFrequencySet(const Vdouble& f) : freq_(f) {}
  FixedFrequencySet* clone() const { return new FixedFrequencySet(*this); }
  const Vdouble& getFrequencies() const { return freq_; }
  std::vector<std::string> getParameterNames() const { return {}; }
  double getParameterValue(const std::string& name) const { throw ParameterNotFoundException(name); }
  void setParameterValue(const std::string& name, double) { throw ParameterNotFoundException(name); }

 private:
  Vdouble freq_;
};

// Model/SubstitutionModelSet.h: models attached to the branches above given node ids, root frequencies, parameters named
// "<model parameter>_<model index + 1>" with aliases (SubstitutionModelSet::aliasParameters).  The set owns its models and
// its root frequency set, like the reference.
class SubstitutionModelSet {
 public:
  explicit SubstitutionModelSet(const Alphabet* alpha) : alphabet_(alpha) {}
  SubstitutionModelSet(const SubstitutionModelSet& o) : alphabet_(o.alphabet_), nodeToModel_(o.nodeToModel_), aliases_(o.aliases_) {
    for (const auto& m : o.models_) models_.emplace_back(m->clone());
    if (o.rootFreqs_) rootFreqs_.reset(o.rootFreqs_->clone());
  }
  SubstitutionModelSet& operator=(const SubstitutionModelSet&) = delete;
  SubstitutionModelSet* clone() const { return new SubstitutionModelSet(*this); }
  const Alphabet* getAlphabet() const { return alphabet_; }
  void setRootFrequencies(FrequencySet* f) { rootFreqs_.reset(f); }
  const FrequencySet* getRootFrequencySet() const { return rootFreqs_.get(); }
  bool isStationary() const { return !rootFreqs_; }
  // SubstitutionModelSet::getRootFrequencies: the root set, or (stationary sets) the first model's equilibrium frequencies
  Vdouble getRootFrequencies() const { return rootFreqs_ ? rootFreqs_->getFrequencies() : models_.at(0)->getFrequencies(); }
  void addModel(SubstitutionModel* model, const std::vector<int>& nodesId) {
    std::unique_ptr<SubstitutionModel> own(model);
    if (!models_.empty() && model->getNumberOfStates() != models_[0]->getNumberOfStates())
      throw Exception("SubstitutionModelSet::addModel. A Substitution Model cannot be added to a Model Set if it does not have the same number of states.");
    for (int id : nodesId) {
      if (nodeToModel_.count(id)) throw Exception("SubstitutionModelSet::addModel. Node " + std::to_string(id) + " already has a model.");
      nodeToModel_[id] = models_.size();
    }
    models_.push_back(std::move(own));
  }
  size_t getNumberOfModels() const { return models_.size(); }
  size_t getNumberOfStates() const { return models_.at(0)->getNumberOfStates(); }
  SubstitutionModel* getModel(size_t i) const { return models_.at(i).get(); }
  size_t getModelIndexForNode(int nodeId) const {
    std::map<int, size_t>::const_iterator it = nodeToModel_.find(nodeId);
    if (it == nodeToModel_.end()) throw Exception("SubstitutionModelSet::getModelIndexForNode(). No model associated to node with id " + std::to_string(nodeId));
    return it->second;
  }
  SubstitutionModel* getModelForNode(int nodeId) const { return getModel(getModelIndexForNode(nodeId)); }
  std::vector<int> getNodesWithModel(size_t i) const {
    std::vector<int> v;
    for (const auto& kv : nodeToModel_) if (kv.second == i) v.push_back(kv.first);
    return v;
  }
  // every node of the tree but the root has a model, and only those (SubstitutionModelSet::isFullySetUpFor)
  bool isFullySetUpFor(const Tree& tree) const {
    const std::vector<Node*> nodes = tree.getNodes();
    for (size_t i = 0; i + 1 < nodes.size(); ++i) if (!nodeToModel_.count(nodes[i]->getId())) return false;
    return !models_.empty();
  }
  // `to` follows `from` from now on (both full names, e.g. "T92.kappa_1", "T92.kappa_2")
  void aliasParameters(const std::string& from, const std::string& to) {
    aliases_[from].push_back(to);
    setParameterValue(to, getParameterValue(from));
  }
  // independent parameters: root frequencies first, then the model parameters that are not aliased to another one
  std::vector<std::string> getParameterNames() const {
    std::vector<std::string> names;
    std::map<std::string, bool> follower;
    for (const auto& kv : aliases_) for (const std::string& t : kv.second) follower[t] = true;
    if (rootFreqs_) for (const std::string& n : rootFreqs_->getParameterNames()) if (!follower.count(n)) names.push_back(n);
    for (size_t k = 0; k < models_.size(); ++k)
      for (const std::string& n : models_[k]->getParameterNames()) {
        const std::string full = n + "_" + std::to_string(k + 1);
        if (!follower.count(full)) names.push_back(full);
      }
    return names;
  }
  double getParameterValue(const std::string& name) const {
    size_t k;
    std::string base;
    if (splitName(name, base, k)) return models_[k]->getParameterValue(base);
    if (rootFreqs_) return rootFreqs_->getParameterValue(name);
    throw ParameterNotFoundException("ParameterNotFoundException: " + name);
  }
  void setParameterValue(const std::string& name, double v) {
    size_t k;
    std::string base;
    if (splitName(name, base, k)) models_[k]->setParameterValue(base, v);
    else if (rootFreqs_) rootFreqs_->setParameterValue(name, v);
    else throw ParameterNotFoundException("ParameterNotFoundException: " + name);
    std::map<std::string, std::vector<std::string> >::const_iterator it = aliases_.find(name);
    if (it != aliases_.end()) for (const std::string& t : it->second) setParameterValue(t, v);
  }

 private:
  bool splitName(const std::string& name, std::string& base, size_t& k) const {
    const size_t u = name.rfind('_');
    if (u == std::string::npos || u + 1 >= name.size()) return false;
    char* end = nullptr;
    const long idx = std::strtol(name.c_str() + u + 1, &end, 10);
    if (*end != 0 || idx < 1 || (size_t)idx > models_.size()) return false;
    base = name.substr(0, u);
    k = (size_t)idx - 1;
    return true;
  }
  const Alphabet* alphabet_;
  std::vector<std::unique_ptr<SubstitutionModel> > models_;
  std::map<int, size_t> nodeToModel_;
  std::unique_ptr<FrequencySet> rootFreqs_;
  std::map<std::string, std::vector<std::string> > aliases_;
};

namespace SubstitutionModelSetTools {
// Model/SubstitutionModelSetTools.cpp:78-183: one copy of `model` per branch (model k + 1 on the k-th node of
// tree.getNodesId() with the root removed), the listed global parameters aliased to the first copy's; takes ownership of
// `model` (deleted, like the reference) and of `rootFreqs`.
inline SubstitutionModelSet* createNonHomogeneousModelSet(SubstitutionModel* model, FrequencySet* rootFreqs, const Tree* tree,
                                                          const std::vector<std::string>& globalParameterNames) {
  std::unique_ptr<SubstitutionModel> tmpl(model);
  const std::vector<std::string> modelParams = model->getParameterNames();
  for (const std::string& g : globalParameterNames)
    if (std::find(modelParams.begin(), modelParams.end(), g) == modelParams.end())
      throw Exception("SubstitutionModelSetTools::createNonHomogeneousModelSet. Parameter '" + g + "' is not valid.");
  SubstitutionModelSet* set = new SubstitutionModelSet(model->getAlphabet());
  if (rootFreqs) set->setRootFrequencies(rootFreqs);
  std::vector<int> ids = tree->getNodesId();
  const int rootId = tree->getRootNode()->getId();
  ids.erase(std::find(ids.begin(), ids.end(), rootId));
  for (int id : ids) set->addModel(model->clone(), std::vector<int>(1, id));
  for (const std::string& g : globalParameterNames)
    for (size_t i = 1; i < ids.size(); ++i) set->aliasParameters(g + "_1", g + "_" + std::to_string(i + 1));
  return set;
}
}  // namespace SubstitutionModelSetTools

// ---- site patterns (SitePatterns.cpp:52-106 through the C ABI) -----------------------------------------------------------------
class SitePatterns {
 public:
  // sequences are taken in the order of `names` (PatternTools::getSequenceSubset re-orders to the tree's leaves)
  SitePatterns(const VectorSiteContainer& sites, const std::vector<std::string>& names) {
    const size_t n = sites.getNumberOfSites(), nt = names.size();
    std::vector<const BasicSequence*> seqs;
    for (const std::string& nm : names) seqs.push_back(&sites.getSequence(nm));
    size_t w = 1;
    for (const BasicSequence* s : seqs)
      for (size_t i = 0; i < n; ++i) w = std::max(w, (*s)[i].size());
    std::vector<uint8_t> cols(n * nt * w, 0);
    for (size_t i = 0; i < n; ++i)
      for (size_t t = 0; t < nt; ++t) std::memcpy(&cols[(i * nt + t) * w], (*seqs[t])[i].data(), (*seqs[t])[i].size());
    patternSite_.resize(n);
    weights_.resize(n);
    indices_.resize(n);
    int64_t np = 0;
    check(bppgpu_site_patterns(cols.data(), (int64_t)n, (int32_t)(nt * w), patternSite_.data(), weights_.data(), indices_.data(), &np),
          "SitePatterns");
    patternSite_.resize((size_t)np);
    weights_.resize((size_t)np);
  }
  const std::vector<unsigned int>& getWeights() const { return weights_; }
  const std::vector<int64_t>& getIndices() const { return indices_; }
  const std::vector<int64_t>& getPatternSites() const { return patternSite_; }

 private:
  std::vector<int64_t> patternSite_, indices_;
  std::vector<unsigned int> weights_;
};

// ---- tree likelihood ---------------------------------------------------------------------------------------------------------------
struct Parameter {
  std::string name;
  double value;
};
typedef std::vector<Parameter> ParameterList;

// Common engine-backed implementation of {R,DR}HomogeneousTreeLikelihood and DRNonHomogeneousTreeLikelihood
class AbstractHomogeneousTreeLikelihood {
 public:
  virtual ~AbstractHomogeneousTreeLikelihood() { if (engine_) bppgpu_destroy(engine_); }
  AbstractHomogeneousTreeLikelihood(const AbstractHomogeneousTreeLikelihood&) = delete;
  AbstractHomogeneousTreeLikelihood& operator=(const AbstractHomogeneousTreeLikelihood&) = delete;

  // AbstractHomogeneousTreeLikelihood::initialize (:235-244)
  void initialize() {
    if (initialized_) throw Exception("Object already initialized.");
    if (!hasData_) throw Exception("Impossible to initialize, no data provided.");
    initialized_ = true;
    fireParameterChanged();
  }
  // value of the function = -lnL (RHomogeneousTreeLikelihood.cpp:287-291)
  double getValue() const {
    if (!initialized_) throw Exception("RHomogeneousTreeLikelihood::getValue(). Instance is not initialized.");
    return minusLogLik_;
  }
  double getLogLikelihood() const { return -getValue(); }
  double getLogLikelihoodForASite(size_t site) const { requireInit(); return siteLnl_[(size_t)siteIndex_[site]]; }
  double getLikelihoodForASite(size_t site) const { return std::exp(getLogLikelihoodForASite(site)); }
  Vdouble getLogLikelihoodForEachSite() const {
    requireInit();
    Vdouble v(siteIndex_.size());
    for (size_t i = 0; i < v.size(); ++i) v[i] = siteLnl_[(size_t)siteIndex_[i]];
    return v;
  }
  size_t getNumberOfSites() const { return siteIndex_.size(); }
  size_t getNumberOfDistinctSites() const { return (size_t)nPatterns_; }
  size_t getSiteIndex(size_t site) const { return (size_t)siteIndex_[site]; }
  size_t getNumberOfStates() const { return model_->getNumberOfStates(); }
  size_t getNumberOfClasses() const { return rDist_->getNumberOfCategories(); }
  const Vdouble& getRootFrequencies() const { return rootFreqs_; }
  const Tree& getTree() const { return *tree_; }

  // "BrLen<i>": i-th node of the post-order list with the root dropped (init_ :155-157)
  ParameterList getBranchLengthsParameters() const {
    ParameterList pl;
    for (size_t i = 0; i < brLen_.size(); ++i) {
      if (reparametrizeRoot_ && ((int)i == root1_ || (int)i == root2_)) continue;
      pl.push_back({"BrLen" + std::to_string(i), brLen_[i]});
    }
    if (reparametrizeRoot_) {  // AbstractNonHomogeneousTreeLikelihood::initBranchLengthsParameters (:386-389)
      pl.push_back({"BrLenRoot", brLen_[(size_t)root1_] + brLen_[(size_t)root2_]});
      pl.push_back({"RootPosition", brLen_[(size_t)root1_] / (brLen_[(size_t)root1_] + brLen_[(size_t)root2_])});
    }
    return pl;
  }
  ParameterList getSubstitutionModelParameters() const {
    ParameterList pl;
    if (modelSet_) for (const std::string& n : modelSet_->getParameterNames()) pl.push_back({n, modelSet_->getParameterValue(n)});
    else for (const std::string& n : model_->getParameterNames()) pl.push_back({n, 0.0});
    return pl;
  }
  const SubstitutionModelSet* getSubstitutionModelSet() const { return modelSet_; }
  double getParameterValue(const std::string& name) const {
    if (reparametrizeRoot_ && name == "BrLenRoot") return brLen_[(size_t)root1_] + brLen_[(size_t)root2_];
    if (reparametrizeRoot_ && name == "RootPosition") return brLen_[(size_t)root1_] / (brLen_[(size_t)root1_] + brLen_[(size_t)root2_]);
    int b = brlenIndex(name);
    if (b < 0) throw ParameterNotFoundException("ParameterNotFoundException: " + name);
    return brLen_[(size_t)b];
  }
  // setParameters -> fireParameterChanged (RHomogeneousTreeLikelihood.cpp:255-283): P(t) of the changed branches (or
  // all of them when a model / rate-distribution parameter moved) are rebuilt and the whole tree is re-pruned
  void setParameterValue(const std::string& name, double value) { setParametersValues({{name, value}}); }
  void setParametersValues(const ParameterList& pl) {
    bool modelChanged = false;
    for (const Parameter& p : pl) {
      if (reparametrizeRoot_ && (p.name == "BrLenRoot" || p.name == "RootPosition")) {
        // applyParameters (AbstractNonHomogeneousTreeLikelihood.cpp:319-330): l1 = len * pos, l2 = len * (1 - pos)
        double len = brLen_[(size_t)root1_] + brLen_[(size_t)root2_], pos = brLen_[(size_t)root1_] / len;
        if (p.name == "BrLenRoot") len = p.value; else pos = p.value;
        brLen_[(size_t)root1_] = len * pos;
        brLen_[(size_t)root2_] = len * (1.0 - pos);
        continue;
      }
      if (applyBranchParameter(p.name, p.value)) continue;
      const int b = brlenIndex(p.name);
      if (b >= 0) brLen_[(size_t)b] = std::min(std::max(p.value, minimumBrLen_), maximumBrLen_);
      else {
        try { if (modelSet_) modelSet_->setParameterValue(p.name, p.value); else model_->setParameterValue(p.name, p.value); }
        catch (ParameterNotFoundException&) { rDist_->setParameterValue(p.name, p.value); }
        modelChanged = true;
      }
    }
    if (modelChanged) uploadModel();
    if (initialized_) fireParameterChanged();
  }
  void setParameters(const ParameterList& pl) { setParametersValues(pl); }

  // derivatives w.r.t. branch lengths of -lnL (RHomogeneousTreeLikelihood.cpp:346-361, DRHomogeneousTreeLikelihood.cpp:340-368)
  double getFirstOrderDerivative(const std::string& variable) const { return derivative(variable, 1); }
  double getSecondOrderDerivative(const std::string& variable) const { return derivative(variable, 2); }
  void enableDerivatives(bool yn) { computeDerivatives_ = yn; }

  // pxy_[node][class][x][y] (getTransitionProbabilitiesPerRateClass)
  VVVdouble getTransitionProbabilitiesPerRateClass(int nodeId, size_t /*siteIndex*/ = 0) const {
    requireInit();
    const size_t S = getNumberOfStates(), C = getNumberOfClasses();
    std::vector<double> buf(C * S * S);
    check(bppgpu_get_transition_probabilities(engine_, 0, nodeId, BPPGPU_WANT_P, buf.data()), "getTransitionProbabilities");
    VVVdouble p(C, VVdouble(S, Vdouble(S)));
    for (size_t c = 0; c < C; ++c)
      for (size_t x = 0; x < S; ++x)
        for (size_t y = 0; y < S; ++y) p[c][x][y] = buf[(c * S + x) * S + y];
    return p;
  }
  // ---- DiscreteRatesAcrossSitesTreeLikelihood accessors (DiscreteRatesAcrossSitesTreeLikelihood.h:70-203) ---------------------------
  // All of them read the root arrays of the last evaluation (DRHomogeneousTreeLikelihood.cpp:203-227: rootSiteLikelihoods_[i][c] =
  // sum_x pi_x rootLikelihoods_[i][c][x]); they stay on the device until one of these is called, then come back once, scaled:
  // true value = array * 2^-exponent, and everything below is combined in log space so that sites under 1e-308 stay finite.
  double getLogLikelihoodForASiteForARateClass(size_t site, size_t rateClass) const {
    ensureRootArrays();
    return rootLogS_[(size_t)siteIndex_[site] * getNumberOfClasses() + rateClass];
  }
  double getLikelihoodForASiteForARateClass(size_t site, size_t rateClass) const { return std::exp(getLogLikelihoodForASiteForARateClass(site, rateClass)); }
  double getLogLikelihoodForASiteForARateClassForAState(size_t site, size_t rateClass, int state) const {
    ensureRootArrays();
    const size_t C = getNumberOfClasses(), S = getNumberOfStates(), row = (size_t)siteIndex_[site] * C + rateClass;
    return std::log(rootL_[row * S + (size_t)state]) - rootExp_[row] * 0.693147180559945309417232121458;
  }
  double getLikelihoodForASiteForARateClassForAState(size_t site, size_t rateClass, int state) const {
    ensureRootArrays();
    const size_t C = getNumberOfClasses(), S = getNumberOfStates(), row = (size_t)siteIndex_[site] * C + rateClass;
    return std::ldexp(rootL_[row * S + (size_t)state], -rootExp_[row]);
  }
  // AbstractDiscreteRatesAcrossSitesTreeLikelihood.cpp:110-133: sum_c p_c L[site][c][state]
  double getLikelihoodForASiteForAState(size_t site, int state) const {
    double l = 0;
    for (size_t c = 0; c < getNumberOfClasses(); ++c) l += getLikelihoodForASiteForARateClassForAState(site, c, state) * rDist_->getProbability(c);
    return l;
  }
  double getLogLikelihoodForASiteForAState(size_t site, int state) const { return std::log(getLikelihoodForASiteForAState(site, state)); }
  VVdouble getLikelihoodForEachSiteForEachRateClass() const { return eachSiteEachClass(false); }       // :92-106
  VVdouble getLogLikelihoodForEachSiteForEachRateClass() const { return eachSiteEachClass(true); }     // :137-151
  VVVdouble getLikelihoodForEachSiteForEachRateClassForEachState() const { return eachSiteEachClassEachState(false); }     // :155-174
  VVVdouble getLogLikelihoodForEachSiteForEachRateClassForEachState() const { return eachSiteEachClassEachState(true); }   // :178-197
  // :201-215  pb[i][c] = L[i][c] p_c / L[i]
  VVdouble getPosteriorProbabilitiesOfEachRate() const {
    ensureRootArrays();
    const size_t C = getNumberOfClasses();
    VVdouble pb(siteIndex_.size(), Vdouble(C));
    for (size_t i = 0; i < pb.size(); ++i) {
      const size_t k = (size_t)siteIndex_[i];
      for (size_t c = 0; c < C; ++c) pb[i][c] = std::exp(rootLogS_[k * C + c] - siteLnl_[k]) * rDist_->getProbability(c);
    }
    return pb;
  }
  // :219-234  sum_c (L[i][c] / L[i]) p_c r_c
  Vdouble getPosteriorRateOfEachSite() const {
    const VVdouble pb = getPosteriorProbabilitiesOfEachRate();
    Vdouble rates(pb.size(), 0.0);
    for (size_t i = 0; i < pb.size(); ++i)
      for (size_t c = 0; c < pb[i].size(); ++c) rates[i] += pb[i][c] * rDist_->getCategory(c);
    return rates;
  }
  // :238-248  whichMax of L[i][.] (first maximum; NOT weighted by p_c, like the reference)
  std::vector<size_t> getRateClassWithMaxPostProbOfEachSite() const {
    ensureRootArrays();
    const size_t C = getNumberOfClasses();
    std::vector<size_t> classes(siteIndex_.size(), 0);
    for (size_t i = 0; i < classes.size(); ++i) {
      const double* l = &rootLogS_[(size_t)siteIndex_[i] * C];
      for (size_t c = 1; c < C; ++c) if (l[c] > l[classes[i]]) classes[i] = c;
    }
    return classes;
  }
  Vdouble getRateWithMaxPostProbOfEachSite() const {   // :252-262
    const std::vector<size_t> cl = getRateClassWithMaxPostProbOfEachSite();
    Vdouble rates(cl.size());
    for (size_t i = 0; i < cl.size(); ++i) rates[i] = rDist_->getCategory(cl[i]);
    return rates;
  }

  // DRTreeLikelihood::computeLikelihoodAtNode-style access to the device-resident conditional likelihoods of an internal
  // node (subtree below it): true value = likelihoodArray[i][c][x] * 2^-scale[i][c]
  void getLikelihoodArray(int nodeId, VVVdouble& likelihoodArray, std::vector<std::vector<int> >& scale) const {
    requireInit();
    const size_t S = getNumberOfStates(), C = getNumberOfClasses(), N = (size_t)nPatterns_;
    std::vector<double> buf(N * C * S);
    std::vector<int32_t> ex(N * C);
    check(bppgpu_get_clv(engine_, 0, nodeId, 0, buf.data(), ex.data()), "getLikelihoodArray");
    likelihoodArray.assign(N, VVdouble(C, Vdouble(S)));
    scale.assign(N, std::vector<int>(C));
    for (size_t i = 0; i < N; ++i)
      for (size_t c = 0; c < C; ++c) {
        scale[i][c] = ex[i * C + c];
        for (size_t x = 0; x < S; ++x) likelihoodArray[i][c][x] = buf[(i * C + c) * S + x];
      }
  }
  // DRTreeLikelihood::computeLikelihoodAtNode(nodeId, VVVdouble&) (Likelihood/DRTreeLikelihood.h:92-102): the conditional
  // likelihood of ALL the data given the state at the node, computed on the device from the resident lower / upper
  // arrays; true value = likelihoodArray[i][c][x] * 2^-scale[i][c] (scale may be null: values are then de-scaled)
  void computeLikelihoodAtNode(int nodeId, VVVdouble& likelihoodArray, std::vector<std::vector<int> >* scale = nullptr) const {
    requireInit();
    ensureDerivativePass();
    const size_t S = getNumberOfStates(), C = getNumberOfClasses(), N = (size_t)nPatterns_;
    std::vector<double> buf(N * C * S);
    std::vector<int32_t> ex(N * C);
    check(bppgpu_get_node_posteriors(engine_, 0, nodeId, buf.data(), ex.data(), nullptr), "computeLikelihoodAtNode");
    likelihoodArray.assign(N, VVdouble(C, Vdouble(S)));
    if (scale) scale->assign(N, std::vector<int>(C));
    for (size_t i = 0; i < N; ++i)
      for (size_t c = 0; c < C; ++c) {
        if (scale) (*scale)[i][c] = ex[i * C + c];
        for (size_t x = 0; x < S; ++x)
          likelihoodArray[i][c][x] = scale ? buf[(i * C + c) * S + x] : std::ldexp(buf[(i * C + c) * S + x], -ex[i * C + c]);
      }
  }
  // DRTreeLikelihoodTools::getPosteriorProbabilitiesForEachStateForEachRate(drl, nodeId) (DRTreeLikelihoodTools.cpp:46-119)
  VVVdouble getPosteriorProbabilitiesForEachStateForEachRate(int nodeId) const {
    requireInit();
    ensureDerivativePass();
    const size_t S = getNumberOfStates(), C = getNumberOfClasses(), N = (size_t)nPatterns_;
    std::vector<double> buf(N * C * S);
    check(bppgpu_get_node_posteriors(engine_, 0, nodeId, nullptr, nullptr, buf.data()), "getPosteriorProbabilities");
    VVVdouble p(N, VVdouble(C, Vdouble(S)));
    for (size_t i = 0; i < N; ++i)
      for (size_t c = 0; c < C; ++c)
        for (size_t x = 0; x < S; ++x) p[i][c][x] = buf[(i * C + c) * S + x];
    return p;
  }
  // MarginalAncestralStateReconstruction::getAncestralStatesForNode: argmax_x sum_c posterior, one state per distinct site
  std::vector<size_t> getAncestralStatesForNode(int nodeId) const {
    const VVVdouble p = getPosteriorProbabilitiesForEachStateForEachRate(nodeId);
    std::vector<size_t> best(p.size(), 0);
    for (size_t i = 0; i < p.size(); ++i) {
      double bv = -1;
      for (size_t x = 0; x < p[i][0].size(); ++x) {
        double v = 0;
        for (size_t c = 0; c < p[i].size(); ++c) v += p[i][c][x];
        if (v > bv) { bv = v; best[i] = x; }
      }
    }
    return best;
  }
  // MarginalNonRevAncestralStateReconstruction's per-node tables (fork, MarginalNonRev...cpp:10-136) from the device-resident
  // arrays: post[i][x] = P(node = x | site i), joint[i][x][y] = P(node = x, father = y | site i) (null / ignored at the root)
  void getMarginalPosteriors(int nodeId, VVdouble& post, VVVdouble* joint) const {
    requireInit();
    const size_t S = getNumberOfStates(), N = (size_t)nPatterns_;
    const bool isRoot = nodeId == (int)nodes_.size() - 1;
    if (!isRoot) ensureDerivativePass();
    std::vector<double> pb(N * S), jb(joint && !isRoot ? N * S * S : 0);
    check(bppgpu_get_marginal_posteriors(engine_, 0, nodeId, pb.data(), jb.empty() ? nullptr : jb.data()), "MarginalNonRevAncestralStateReconstruction");
    post.assign(N, Vdouble(S));
    for (size_t i = 0; i < N; ++i)
      for (size_t x = 0; x < S; ++x) post[i][x] = pb[i * S + x];
    if (joint) {
      joint->assign(N, VVdouble(S, Vdouble(S, 0.0)));
      if (!isRoot)
        for (size_t i = 0; i < N; ++i)
          for (size_t x = 0; x < S; ++x)
            for (size_t y = 0; y < S; ++y) (*joint)[i][x][y] = jb[(i * S + x) * S + y];
    }
  }
  // MLAncestralStateReconstruction's result (fork): states[node][distinct site] of the joint ML assignment, computed on the
  // device with the transition probabilities and root frequencies of the last evaluation; bestLogLik[i] = its log joint likelihood
  void getJointMLAncestralStates(std::vector<std::vector<size_t> >& states, Vdouble* bestLogLik = nullptr) const {
    requireInit();
    const size_t N = (size_t)nPatterns_, nn = nodes_.size();
    std::vector<int32_t> buf(nn * N);
    Vdouble best(N);
    check(bppgpu_ml_ancestral_states(engine_, 0, buf.data(), best.data()), "MLAncestralStateReconstruction");
    states.assign(nn, std::vector<size_t>(N));
    for (size_t n = 0; n < nn; ++n)
      for (size_t i = 0; i < N; ++i) states[n][i] = (size_t)buf[n * N + i];
    if (bestLogLik) *bestLogLik = best;
  }
  std::vector<int> getNodesId() const {
    std::vector<int> ids;
    for (const Node* n : nodes_) ids.push_back(n->getId());
    return ids;
  }
  long getNumberOfLikelihoodCalculations() const { return numOfLikelihoodCalculations_; }  // fork: DRNonHomogeneousTreeLikelihood.h:75

 protected:
  AbstractHomogeneousTreeLikelihood(const Tree& tree, SubstitutionModel* model, DiscreteDistribution* rDist, bool checkRooted,
                                    unsigned engineFlags, int device)
      : tree_(new Tree(tree)), model_(model), rDist_(rDist), engine_(nullptr), engineFlags_(engineFlags), device_(device),
        initialized_(false), hasData_(false), computeDerivatives_(true), minusLogLik_(0), nPatterns_(0),
        minimumBrLen_(1e-6), maximumBrLen_(1e4), derivsValid_(false), numOfLikelihoodCalculations_(0) {
    // init_ (AbstractHomogeneousTreeLikelihood.cpp:140-166)
    if (checkRooted && tree_->isRooted()) tree_->unroot();
    tree_->resetNodesId();
    nodes_ = tree_->getNodes();
    for (size_t i = 0; i + 1 < nodes_.size(); ++i) {  // initBranchLengthsParameters (:305-337)
      double d = nodes_[i]->hasDistanceToFather() ? nodes_[i]->getDistanceToFather() : minimumBrLen_;
      d = std::min(std::max(d, minimumBrLen_), maximumBrLen_);
      nodes_[i]->setDistanceToFather(d);
      brLen_.push_back(d);
    }
  }

  // setData (RHomogeneousTreeLikelihood.cpp:131-146): sequences re-ordered to the leaves, global pattern compression, tip codes
  void setData(const VectorSiteContainer& sites) {
    if (sites.getNumberOfSequences() == 0 || sites.getNumberOfSites() == 0)
      throw Exception("DRASRTreeLikelihoodData::initLikelihoods. Can't use empty dataset (0 sequences or 0 sites).");
    const std::vector<std::string> leafNames = tree_->getLeavesNames();
    SitePatterns patterns(sites, leafNames);
    nPatterns_ = (int64_t)patterns.getWeights().size();
    siteIndex_ = patterns.getIndices();
    // distinct characters -> code table rows (getInitValue)
    std::map<std::string, int> codeOf;
    std::vector<std::string> chars;
    std::vector<std::vector<uint16_t> > codes(leafNames.size(), std::vector<uint16_t>((size_t)nPatterns_));
    for (size_t t = 0; t < leafNames.size(); ++t) {
      const BasicSequence& seq = sites.getSequence(leafNames[t]);
      for (int64_t k = 0; k < nPatterns_; ++k) {
        const std::string& ch = seq[(size_t)patterns.getPatternSites()[(size_t)k]];
        std::map<std::string, int>::iterator it = codeOf.find(ch);
        if (it == codeOf.end()) { it = codeOf.insert(std::make_pair(ch, (int)chars.size())).first; chars.push_back(ch); }
        codes[t][(size_t)k] = (uint16_t)it->second;
      }
    }
    const size_t S = model_->getNumberOfStates(), C = rDist_->getNumberOfCategories();
    std::vector<double> table(chars.size() * S);
    for (size_t k = 0; k < chars.size(); ++k)
      for (size_t s = 0; s < S; ++s) table[k * S + s] = model_->getInitValue(s, chars[k]);
    // flattened topology: node id = post-order position
    const int nn = (int)nodes_.size();
    std::vector<int32_t> off(nn + 1, 0), children;
    for (int i = 0; i < nn; ++i) {
      for (size_t k = 0; k < nodes_[i]->getNumberOfSons(); ++k) children.push_back(nodes_[i]->getSon(k)->getId());
      off[i + 1] = (int32_t)children.size();
    }
    bppgpu_config cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.n_states = (int32_t)S; cfg.n_cats = (int32_t)C; cfg.n_patterns = nPatterns_; cfg.n_nodes = nn; cfg.root = nn - 1;
    cfg.child_offsets = off.data(); cfg.children = children.data(); cfg.n_points = nPoints_;
    cfg.n_models = modelSet_ ? (int32_t)modelSet_->getNumberOfModels() : (nModelSlots_ > 0 ? nModelSlots_ : nPoints_);
    cfg.n_codes = (int32_t)chars.size(); cfg.code_bytes = chars.size() > 256 ? 2 : 1; cfg.code_table = table.data();
    cfg.device = device_; cfg.flags = engineFlags_ | BPPGPU_FLAG_KEEP_CLVS;
    if (engine_) { bppgpu_destroy(engine_); engine_ = nullptr; }
    check(bppgpu_create(&cfg, &engine_), "TreeLikelihood::setData");
    const std::vector<Node*> leaves = tree_->getLeaves();
    for (size_t t = 0; t < leaves.size(); ++t) {
      if (cfg.code_bytes == 1) {
        std::vector<uint8_t> c8(codes[t].begin(), codes[t].end());
        check(bppgpu_set_tip_codes(engine_, leaves[t]->getId(), c8.data()), "setData");
      } else {
        check(bppgpu_set_tip_codes(engine_, leaves[t]->getId(), codes[t].data()), "setData");
      }
    }
    check(bppgpu_set_pattern_weights(engine_, patterns.getWeights().data()), "setData");
    hasData_ = true;
    uploadModel();
  }

  // hook for classes whose branch lengths are functions of other parameters (clock heights): true = name consumed
  virtual bool applyBranchParameter(const std::string&, double) { return false; }
  virtual Vdouble rootFrequencies() const { return modelSet_ ? modelSet_->getRootFrequencies() : model_->getFrequencies(); }

  virtual void uploadModel() {
    if (!engine_) return;
    bppgpu_model_desc d;
    if (modelSet_) {
      // AbstractNonHomogeneousTreeLikelihood::computeTransitionProbabilitiesForNode (.cpp:410-468): the branch above a node
      // uses modelSet_->getModelForNode(node id); one device slot per model of the set
      for (size_t k = 0; k < modelSet_->getNumberOfModels(); ++k) {
        modelSet_->getModel(k)->fillModelDesc(d);
        check(bppgpu_set_model(engine_, (int32_t)k, &d), "setModel");
      }
      std::vector<int32_t> slot(nodes_.size(), 0);
      for (size_t i = 0; i + 1 < nodes_.size(); ++i) slot[i] = (int32_t)modelSet_->getModelIndexForNode(nodes_[i]->getId());
      check(bppgpu_set_branch_models(engine_, 0, slot.data()), "setBranchModels");
    } else {
      model_->fillModelDesc(d);
      check(bppgpu_set_model(engine_, 0, &d), "setModel");
    }
    Vdouble r(rDist_->getNumberOfCategories()), p(r.size());
    for (size_t c = 0; c < r.size(); ++c) { r[c] = rDist_->getCategory(c); p[c] = rDist_->getProbability(c); }
    check(bppgpu_set_rates(engine_, r.data(), p.data()), "setRates");
    rootFreqs_ = rootFrequencies();
    check(bppgpu_set_root_freqs(engine_, 0, rootFreqs_.data()), "setRootFreqs");
  }

  // computeAllTransitionProbabilities + computeTreeLikelihood (+ the DR derivative passes) in one device evaluation
  virtual void fireParameterChanged() {
    Vdouble t(nodes_.size(), 0.0);
    for (size_t i = 0; i < brLen_.size(); ++i) t[i] = brLen_[i];
    check(bppgpu_set_branch_lengths(engine_, 0, t.data()), "applyParameters");
    double lnl = 0;
    check(bppgpu_eval(engine_, BPPGPU_EVAL_LNL, &lnl, nullptr, nullptr), "computeTreeLikelihood");
    ++numOfLikelihoodCalculations_;
    minusLogLik_ = -lnl;
    siteLnl_.resize((size_t)nPatterns_);
    check(bppgpu_get_site_lnl(engine_, 0, siteLnl_.data()), "getLogLikelihoodForEachSite");
    if (engineFlags_ & BPPGPU_FLAG_WEIGHTED_ROOT) check(bppgpu_get_root_freqs(engine_, 0, rootFreqs_.data()), "getRootFrequencies");
    derivsValid_ = false;
    rootArraysValid_ = false;
  }

  // the prefix (upper) arrays exist after an evaluation with derivatives
  void ensureDerivativePass() const {
    if (derivsValid_) return;
    d1_.assign(nodes_.size(), 0.0);
    d2_.assign(nodes_.size(), 0.0);
    double lnl = 0;
    check(bppgpu_eval(engine_, BPPGPU_EVAL_LNL | BPPGPU_EVAL_D1 | BPPGPU_EVAL_D2, &lnl, d1_.data(), d2_.data()), "computeTreeDLikelihoods");
    derivsValid_ = true;
  }
  double derivative(const std::string& variable, int order) const {
    requireInit();
    if (reparametrizeRoot_ && (variable == "BrLenRoot" || variable == "RootPosition")) {
      // DRNonHomogeneousTreeLikelihood.cpp:445-478 (first order), :576-867 (second order: needs the cross term of the two
      // root branches, rebuilt on the device)
      ensureDerivativePass();
      double o[4];
      check(bppgpu_get_root_reparam_derivatives(engine_, 0, o), "getSecondOrderDerivative");
      const int k = (order == 1 ? 0 : 2) + (variable == "BrLenRoot" ? 0 : 1);
      return -o[k];
    }
    const int b = brlenIndex(variable);
    if (b < 0) {
      for (const std::string& n : modelSet_ ? modelSet_->getParameterNames() : model_->getParameterNames())
        if (n == variable) throw Exception("Derivatives respective to substitution model parameters are not implemented.");
      throw ParameterNotFoundException("ParameterNotFoundException: " + variable);
    }
    if (!derivsValid_) {
      d1_.assign(nodes_.size(), 0.0);
      d2_.assign(nodes_.size(), 0.0);
      double lnl = 0;
      check(bppgpu_eval(engine_, BPPGPU_EVAL_LNL | BPPGPU_EVAL_D1 | BPPGPU_EVAL_D2, &lnl, d1_.data(), d2_.data()), "computeTreeDLikelihoods");
      derivsValid_ = true;
    }
    return order == 1 ? -d1_[(size_t)b] : -d2_[(size_t)b];
  }
  int brlenIndex(const std::string& name) const {
    if (name.compare(0, 5, "BrLen") != 0) return -1;
    char* end = nullptr;
    const long i = std::strtol(name.c_str() + 5, &end, 10);
    if (*end != 0 || i < 0 || (size_t)i >= brLen_.size()) return -1;
    return (int)i;
  }
  void requireInit() const { if (!initialized_) throw Exception("Instance is not initialized."); }
  // root arrays of the last evaluation: rootL_[i][c][x] (scaled), rootExp_[i][c], rootLogS_[i][c] = log sum_x pi_x L - exp ln 2
  void ensureRootArrays() const {
    requireInit();
    if (rootArraysValid_) return;
    const size_t S = getNumberOfStates(), C = getNumberOfClasses(), N = (size_t)nPatterns_;
    rootL_.resize(N * C * S);
    std::vector<int32_t> ex(N * C);
    check(bppgpu_get_clv(engine_, 0, (int32_t)nodes_.size() - 1, 0, rootL_.data(), ex.data()), "getRootLikelihoodArray");
    rootExp_.assign(ex.begin(), ex.end());
    rootLogS_.resize(N * C);
    for (size_t r = 0; r < N * C; ++r) {
      double s = 0;
      for (size_t x = 0; x < S; ++x) s += rootFreqs_[x] * rootL_[r * S + x];
      rootLogS_[r] = std::log(s) - rootExp_[r] * 0.693147180559945309417232121458;
    }
    rootArraysValid_ = true;
  }
  VVdouble eachSiteEachClass(bool logs) const {
    ensureRootArrays();
    const size_t C = getNumberOfClasses();
    VVdouble l(siteIndex_.size(), Vdouble(C));
    for (size_t i = 0; i < l.size(); ++i)
      for (size_t c = 0; c < C; ++c) {
        const double v = rootLogS_[(size_t)siteIndex_[i] * C + c];
        l[i][c] = logs ? v : std::exp(v);
      }
    return l;
  }
  VVVdouble eachSiteEachClassEachState(bool logs) const {
    const size_t C = getNumberOfClasses(), S = getNumberOfStates();
    VVVdouble l(siteIndex_.size(), VVdouble(C, Vdouble(S)));
    for (size_t i = 0; i < l.size(); ++i)
      for (size_t c = 0; c < C; ++c)
        for (size_t x = 0; x < S; ++x)
          l[i][c][x] = logs ? getLogLikelihoodForASiteForARateClassForAState(i, c, (int)x) : getLikelihoodForASiteForARateClassForAState(i, c, (int)x);
    return l;
  }

  std::unique_ptr<Tree> tree_;
  SubstitutionModel* model_;      // not owned (like the reference)
  SubstitutionModelSet* modelSet_ = nullptr;  // not owned; non-homogeneous classes: one model per branch group
  DiscreteDistribution* rDist_;   // not owned
  bppgpu_engine* engine_;
  unsigned engineFlags_;
  int device_;
  bool initialized_, hasData_, computeDerivatives_;
  double minusLogLik_;
  int64_t nPatterns_;
  std::vector<Node*> nodes_;
  Vdouble brLen_;
  double minimumBrLen_, maximumBrLen_;
  std::vector<int64_t> siteIndex_;
  Vdouble siteLnl_, rootFreqs_;
  mutable Vdouble d1_, d2_;
  mutable bool derivsValid_;
  mutable bool rootArraysValid_ = false;
  mutable Vdouble rootL_, rootLogS_;
  mutable std::vector<int> rootExp_;
  long numOfLikelihoodCalculations_;
  int nPoints_ = 1;  // parameter points evaluated per device call (LikelihoodPointBatch)
  int nModelSlots_ = 0;  // device model slots when they are not one per point (RNonHomogeneousMixedTreeLikelihood)
  bool reparametrizeRoot_ = false;  // BrLenRoot / RootPosition replace the two root branches (NH classes, rooted trees)
  int root1_ = -1, root2_ = -1;     // ids (= BrLen indices) of the root's first two sons
};

// Likelihood/RHomogeneousTreeLikelihood.h:108-138.  `usePatterns` (recursive per-subtree compression) changes only the
// memory layout of the reference, not its results; the device path always uses the global compression.
class RHomogeneousTreeLikelihood : public AbstractHomogeneousTreeLikelihood {
 public:
  RHomogeneousTreeLikelihood(const Tree& tree, const VectorSiteContainer& data, SubstitutionModel* model, DiscreteDistribution* rDist,
                             bool checkRooted = true, bool verbose = true, bool usePatterns = true, int device = 0)
      : AbstractHomogeneousTreeLikelihood(tree, model, rDist, checkRooted, BPPGPU_FLAG_R_SEMANTICS, device) {
    (void)verbose; (void)usePatterns;
    setData(data);
  }
};
// Likelihood/RHomogeneousClockTreeLikelihood.{h,cpp}: the same likelihood with the branch lengths of a rooted, bifurcating tree
// driven by node heights -- "TotalHeight" (height of the root: the longest path to a leaf, TreeTemplateTools::getHeights,
// TreeTemplateTools.cpp:173-186) and "HeightP<id>" (height / father's height) for every internal non-root node
// (initBranchLengthsParameters :121-157, computeBranchLengthsFromHeights :161-179, minimum branch length 0 :87).  No branch
// derivatives (getDerivableParameters is empty, :183-187).
class RHomogeneousClockTreeLikelihood : public RHomogeneousTreeLikelihood {
 public:
  RHomogeneousClockTreeLikelihood(const Tree& tree, const VectorSiteContainer& data, SubstitutionModel* model, DiscreteDistribution* rDist,
                                  bool checkRooted = true, bool verbose = true, int device = 0)
      : RHomogeneousTreeLikelihood(tree, data, model, rDist, false, verbose, true, device) {
    (void)checkRooted;
    if (!tree_->isRooted()) throw Exception("RHomogeneousClockTreeLikelihood::init_(). Tree is unrooted!");
    for (const Node* n : nodes_)
      if (n->getNumberOfSons() > 2) throw Exception("HomogeneousClockTreeLikelihood::init_(). Tree is multifurcating.");
    minimumBrLen_ = 0.0;
    std::vector<double> h(nodes_.size(), 0.0);
    for (size_t i = 0; i < nodes_.size(); ++i)   // post-order: sons first
      for (size_t k = 0; k < nodes_[i]->getNumberOfSons(); ++k) {
        const int s = nodes_[i]->getSon(k)->getId();
        h[i] = std::max(h[i], h[(size_t)s] + brLen_[(size_t)s]);
      }
    totalHeight_ = h.back();
    for (size_t i = 0; i + 1 < nodes_.size(); ++i)
      if (!nodes_[i]->isLeaf()) heightP_[(int)i] = h[i] / h[(size_t)nodes_[i]->getFather()->getId()];
    computeBranchLengthsFromHeights(nodes_.back(), totalHeight_);
  }
  ParameterList getBranchLengthsParameters() const {
    ParameterList pl;
    pl.push_back({"TotalHeight", totalHeight_});
    for (const auto& kv : heightP_) pl.push_back({"HeightP" + std::to_string(kv.first), kv.second});
    return pl;
  }
  double getParameterValue(const std::string& name) const {
    if (name == "TotalHeight") return totalHeight_;
    if (name.compare(0, 7, "HeightP") == 0) {
      std::map<int, double>::const_iterator it = heightP_.find(std::atoi(name.c_str() + 7));
      if (it != heightP_.end()) return it->second;
    }
    throw ParameterNotFoundException("ParameterNotFoundException: " + name);
  }
  ParameterList getDerivableParameters() const { requireInit(); return ParameterList(); }
  double getFirstOrderDerivative(const std::string& variable) const {
    throw Exception("RHomogeneousClockTreeLikelihood: no derivative with respect to " + variable + " (all parameters are non-derivable).");
  }
  double getSecondOrderDerivative(const std::string& variable) const { return getFirstOrderDerivative(variable); }

 protected:
  bool applyBranchParameter(const std::string& name, double value) override {
    if (name == "TotalHeight") totalHeight_ = value;
    else if (name.compare(0, 7, "HeightP") == 0 && heightP_.count(std::atoi(name.c_str() + 7))) heightP_[std::atoi(name.c_str() + 7)] = value;
    else if (name.compare(0, 5, "BrLen") == 0) throw ParameterNotFoundException("ParameterNotFoundException: " + name);
    else return false;
    computeBranchLengthsFromHeights(nodes_.back(), totalHeight_);
    return true;
  }

 private:
  void computeBranchLengthsFromHeights(const Node* node, double height) {
    for (size_t i = 0; i < node->getNumberOfSons(); ++i) {
      const Node* son = node->getSon(i);
      if (son->isLeaf()) brLen_[(size_t)son->getId()] = std::max(minimumBrLen_, height);
      else {
        const double sonHeight = heightP_.at(son->getId()) * height;
        brLen_[(size_t)son->getId()] = std::max(minimumBrLen_, height - sonHeight);
        computeBranchLengthsFromHeights(son, sonHeight);
      }
    }
  }
  double totalHeight_ = 0;
  std::map<int, double> heightP_;
};
// Likelihood/DRHomogeneousTreeLikelihood.h
class DRHomogeneousTreeLikelihood : public AbstractHomogeneousTreeLikelihood {
 public:
  DRHomogeneousTreeLikelihood(const Tree& tree, const VectorSiteContainer& data, SubstitutionModel* model, DiscreteDistribution* rDist,
                              bool checkRooted = true, bool verbose = true, int device = 0)
      : AbstractHomogeneousTreeLikelihood(tree, model, rDist, checkRooted, 0, device) {
    (void)verbose;
    setData(data);
  }
};
// Likelihood/DRNonHomogeneousTreeLikelihood.h:93-146, fork constructor (weightedRootFreq, calculateDerivatives); one model on
// every branch (what ChromosomeNumberOptimizer builds), the tree is kept rooted
class DRNonHomogeneousTreeLikelihood : public AbstractHomogeneousTreeLikelihood {
 public:
  DRNonHomogeneousTreeLikelihood(const Tree& tree, const VectorSiteContainer& data, bool weightedRootFreq, bool calculateDerivatives,
                                 SubstitutionModel* model, DiscreteDistribution* rDist, const Vdouble* rootFreqs = nullptr,
                                 bool verbose = true, int device = 0, bool reparametrizeRoot = false)
      : AbstractHomogeneousTreeLikelihood(tree, model, rDist, false,
                                          BPPGPU_FLAG_NH_DERIV | (weightedRootFreq ? BPPGPU_FLAG_WEIGHTED_ROOT : 0u), device) {
    (void)verbose;
    if (reparametrizeRoot) {  // AbstractNonHomogeneousTreeLikelihood::init_ (:162-189): root1_ / root2_ = the root's two sons
      const Node* root = nodes_.back();
      if (root->getNumberOfSons() != 2) throw Exception("reparametrizeRoot needs a rooted tree (a root with two sons)");
      root1_ = root->getSon(0)->getId();
      root2_ = root->getSon(1)->getId();
      reparametrizeRoot_ = true;
    }
    computeDerivatives_ = calculateDerivatives;
    if (rootFreqs) fixedRootFreqs_ = *rootFreqs;
    setData(data);
  }

  // Likelihood/DRNonHomogeneousTreeLikelihood.h:93-111: (tree, data, SubstitutionModelSet*, rDist, verbose, reparametrizeRoot)
  DRNonHomogeneousTreeLikelihood(const Tree& tree, const VectorSiteContainer& data, SubstitutionModelSet* modelSet,
                                 DiscreteDistribution* rDist, bool verbose = true, bool reparametrizeRoot = false, int device = 0,
                                 unsigned extraFlags = 0)
      : AbstractHomogeneousTreeLikelihood(tree, modelSet->getModel(0), rDist, false, BPPGPU_FLAG_NH_DERIV | extraFlags, device) {
    (void)verbose;
    // AbstractNonHomogeneousTreeLikelihood::setSubstitutionModelSet (.cpp:193-216)
    if (!modelSet->isFullySetUpFor(*tree_)) throw Exception("AbstractNonHomogeneousTreeLikelihood::init_(). Model set is not fully specified.");
    modelSet_ = modelSet;
    if (reparametrizeRoot) {
      const Node* root = nodes_.back();
      if (root->getNumberOfSons() != 2) throw Exception("reparametrizeRoot needs a rooted tree (a root with two sons)");
      root1_ = root->getSon(0)->getId();
      root2_ = root->getSon(1)->getId();
      reparametrizeRoot_ = true;
    }
    setData(data);
  }

 protected:
  Vdouble rootFrequencies() const {
    if (modelSet_) return modelSet_->getRootFrequencies();
    return fixedRootFreqs_.empty() ? model_->getFrequencies() : fixedRootFreqs_;
  }

 private:
  Vdouble fixedRootFreqs_;
};
// Likelihood/RNonHomogeneousTreeLikelihood.h: (tree, data, modelSet, rDist, verbose, usePatterns, reparametrizeRoot); the R
// classes' root reduction (non-positive terms dropped), same device path
class RNonHomogeneousTreeLikelihood : public DRNonHomogeneousTreeLikelihood {
 public:
  RNonHomogeneousTreeLikelihood(const Tree& tree, const VectorSiteContainer& data, SubstitutionModelSet* modelSet,
                                DiscreteDistribution* rDist, bool verbose = true, bool usePatterns = true, bool reparametrizeRoot = false,
                                int device = 0)
      : DRNonHomogeneousTreeLikelihood(tree, data, modelSet, rDist, verbose, reparametrizeRoot, device, BPPGPU_FLAG_R_SEMANTICS) {
    (void)usePatterns;
  }
};

// ---- batched front-end ------------------------------------------------------------------------------------------
// ChromosomeNumberOptimizer keeps a vector of DRNonHomogeneousTreeLikelihood objects, one per starting point, and evaluates
// and line-searches them one after the other (Likelihood/ChromosomeNumberOptimizer.cpp:58, :141-153, :472-517).  This class
// is that vector as ONE device object: same tree and data, one substitution model (and optionally one set of branch lengths)
// per point, every point's -lnL from a single bppgpu_eval (batched P(t) + one launch per node covering all points).
// Values are identical to what a DRNonHomogeneousTreeLikelihood built on models[k] returns.
class LikelihoodPointBatch : public AbstractHomogeneousTreeLikelihood {
 public:
  LikelihoodPointBatch(const Tree& tree, const VectorSiteContainer& data, bool weightedRootFreq,
                       const std::vector<SubstitutionModel*>& models, DiscreteDistribution* rDist, int device = 0,
                       bool checkRooted = false, unsigned engineFlags = BPPGPU_FLAG_NH_DERIV)
      : AbstractHomogeneousTreeLikelihood(tree, models.at(0), rDist, checkRooted,
                                          engineFlags | (weightedRootFreq ? BPPGPU_FLAG_WEIGHTED_ROOT : 0u), device),
        models_(models) {
    nPoints_ = (int)models.size();
    computeDerivatives_ = false;
    pointBrLen_.assign(models.size(), brLen_);
    values_.assign(models.size(), 0.0);
    setData(data);
  }
  size_t getNumberOfPoints() const { return models_.size(); }
  // a model parameter of point k moved (the caller changed models[k]): its eigensystem is re-uploaded before the next evaluation
  void modelChanged(size_t k) { dirty_.at(k) = 1; if (initialized_) fireParameterChanged(); }
  // several points changed: mark them all, then evaluate() once (one device call for every point)
  void markModelChanged(size_t k) { dirty_.at(k) = 1; }
  void evaluate() { requireInit(); fireParameterChanged(); }
  void setBranchLengths(size_t k, const Vdouble& brlen) {
    if (brlen.size() != brLen_.size()) throw Exception("LikelihoodPointBatch::setBranchLengths: wrong number of branch lengths");
    for (size_t i = 0; i < brlen.size(); ++i) pointBrLen_.at(k)[i] = std::min(std::max(brlen[i], minimumBrLen_), maximumBrLen_);
    if (initialized_) fireParameterChanged();
  }
  // -lnL of every point (getValue() of the k-th likelihood of the reference's vector)
  const Vdouble& getValues() const { requireInit(); return values_; }
  double getValue(size_t k) const { requireInit(); return values_.at(k); }
  // index of the best point (the reference sorts its vector with compareLikValues, ChromosomeNumberOptimizer.cpp:156)
  size_t getBestPoint() const {
    requireInit();
    size_t b = 0;
    for (size_t k = 1; k < values_.size(); ++k) if (values_[k] < values_[b]) b = k;
    return b;
  }

 protected:
  void uploadModel() override {
    if (!engine_) return;
    if (dirty_.size() != models_.size()) dirty_.assign(models_.size(), 1);
    Vdouble r(rDist_->getNumberOfCategories()), p(r.size());
    for (size_t c = 0; c < r.size(); ++c) { r[c] = rDist_->getCategory(c); p[c] = rDist_->getProbability(c); }
    check(bppgpu_set_rates(engine_, r.data(), p.data()), "setRates");
    for (size_t k = 0; k < models_.size(); ++k) {
      if (!dirty_[k]) continue;
      bppgpu_model_desc d;
      models_[k]->fillModelDesc(d);
      check(bppgpu_set_model(engine_, (int32_t)k, &d), "setModel");
      const Vdouble f = models_[k]->getFrequencies();
      check(bppgpu_set_root_freqs(engine_, (int32_t)k, f.data()), "setRootFreqs");
      dirty_[k] = 0;
    }
  }
  void fireParameterChanged() override {
    uploadModel();
    for (size_t k = 0; k < models_.size(); ++k) {
      Vdouble t(nodes_.size(), 0.0);
      for (size_t i = 0; i < brLen_.size(); ++i) t[i] = pointBrLen_[k][i];
      check(bppgpu_set_branch_lengths(engine_, (int32_t)k, t.data()), "applyParameters");
    }
    Vdouble lnl(models_.size(), 0.0);
    check(bppgpu_eval(engine_, BPPGPU_EVAL_LNL, lnl.data(), nullptr, nullptr), "computeTreeLikelihood");
    numOfLikelihoodCalculations_ += (long)models_.size();
    for (size_t k = 0; k < models_.size(); ++k) values_[k] = -lnl[k];
    minusLogLik_ = values_[0];
    derivsValid_ = false;
    rootArraysValid_ = false;
  }

 private:
  std::vector<SubstitutionModel*> models_;  // not owned
  std::vector<Vdouble> pointBrLen_;
  std::vector<char> dirty_;
  Vdouble values_;
};

// ---- mixture of substitution models ------------------------------------------------------------------------------
// RHomogeneousMixedTreeLikelihood keeps one RHomogeneousTreeLikelihood per sub-model of a MixedSubstitutionModel and combines
// them per site and rate class with the sub-model probabilities, L_site = sum_k probas_k L_k,site
// (Likelihood/RHomogeneousMixedTreeLikelihood.cpp:191-212; YNGP M-series, RELAX).  Here the sub-likelihoods are the points of one
// device object and the combination is a log-sum-exp over their per-site log-likelihoods.
class RHomogeneousMixedTreeLikelihood : public LikelihoodPointBatch {
 public:
  RHomogeneousMixedTreeLikelihood(const Tree& tree, const VectorSiteContainer& data, const std::vector<SubstitutionModel*>& subModels,
                                  const Vdouble& probas, DiscreteDistribution* rDist, int device = 0)
      : LikelihoodPointBatch(tree, data, false, subModels, rDist, device, /*checkRooted=*/true, /*engineFlags=*/0u), probas_(probas) {
    if (probas_.size() != subModels.size()) throw Exception("RHomogeneousMixedTreeLikelihood: one probability per sub-model");
  }
  // Likelihood/RHomogeneousMixedTreeLikelihood.h: (tree, data, model, rDist, checkRooted, verbose, usePatterns) with a mixed model
  RHomogeneousMixedTreeLikelihood(const Tree& tree, const VectorSiteContainer& data, MixedSubstitutionModel* model, DiscreteDistribution* rDist,
                                  bool checkRooted = true, bool verbose = true, bool usePatterns = true, int device = 0)
      : LikelihoodPointBatch(tree, data, false, subModelsOf(model), rDist, device, checkRooted, /*engineFlags=*/0u),
        probas_(model->getProbabilities()) { (void)verbose; (void)usePatterns; }
  static std::vector<SubstitutionModel*> subModelsOf(const MixedSubstitutionModel* m) {
    std::vector<SubstitutionModel*> v;
    for (size_t k = 0; k < m->getNumberOfModels(); ++k) v.push_back(m->getNModel(k));
    return v;
  }
  void setProbabilities(const Vdouble& p) { probas_ = p; if (initialized_) combine(); }
  double getValue() const { requireInit(); return mixedMinusLogLik_; }
  double getLogLikelihood() const { return -getValue(); }
  double getLogLikelihoodForASite(size_t site) const { requireInit(); return mixedSiteLnl_[(size_t)siteIndex_[site]]; }

 protected:
  void fireParameterChanged() override {
    LikelihoodPointBatch::fireParameterChanged();
    combine();
  }

 private:
  void combine() {
    const size_t K = getNumberOfPoints(), N = (size_t)nPatterns_;
    std::vector<Vdouble> sl(K, Vdouble(N));
    for (size_t k = 0; k < K; ++k) check(bppgpu_get_site_lnl(engine_, (int32_t)k, sl[k].data()), "getLogLikelihoodForEachSite");
    mixedSiteLnl_.assign(N, 0.0);
    for (size_t i = 0; i < N; ++i) {
      double m = -std::numeric_limits<double>::infinity();
      for (size_t k = 0; k < K; ++k) if (probas_[k] > 0) m = std::max(m, sl[k][i]);
      double s = 0;
      for (size_t k = 0; k < K; ++k) if (probas_[k] > 0) s += probas_[k] * std::exp(sl[k][i] - m);
      mixedSiteLnl_[i] = std::isfinite(m) ? m + std::log(s) : m;
    }
    // getLogLikelihood (RHomogeneousTreeLikelihood.cpp:162-176): every site, sorted, summed
    Vdouble la(siteIndex_.size());
    for (size_t j = 0; j < la.size(); ++j) la[j] = mixedSiteLnl_[(size_t)siteIndex_[j]];
    std::sort(la.begin(), la.end());
    double ll = 0;
    for (size_t j = la.size(); j > 0; --j) ll += la[j - 1];
    mixedMinusLogLik_ = -ll;
  }
  Vdouble probas_, mixedSiteLnl_;
  double mixedMinusLogLik_ = 0;
};

// ---- joint ML ancestral reconstruction (fork: Likelihood/MLAncestralStateReconstruction.{h,cpp}, Pupko et al. 2000) -------------
// The reference constructor takes the likelihood, its model, the root frequencies and its pxy_ map
// (MLAncestralStateReconstruction.h:88-103); here the device already holds that likelihood's tables, so `model` and `Pijt` are
// accepted for source compatibility only and `rootFrequencies` must be the likelihood's own.
class MLAncestralStateReconstruction {
 public:
  MLAncestralStateReconstruction(const AbstractHomogeneousTreeLikelihood* drl, const SubstitutionModel* model, const std::vector<double>& rootFrequencies,
                                 const void* Pijt = nullptr)
      : likelihood_(drl) {
    (void)model; (void)Pijt;
    const Vdouble& rf = drl->getRootFrequencies();
    if (rootFrequencies.size() != rf.size()) throw Exception("MLAncestralStateReconstruction: wrong number of root frequencies");
    for (size_t x = 0; x < rf.size(); ++x)
      if (std::fabs(rf[x] - rootFrequencies[x]) > 1e-12)
        throw Exception("MLAncestralStateReconstruction: root frequencies other than the likelihood's own are not supported");
  }
  void computeJointLikelihood() { likelihood_->getJointMLAncestralStates(states_, &bestLogLik_); computed_ = true; }
  // node id -> best state per distinct site (getAllAncestralStates, .cpp:136-141)
  std::map<int, std::vector<size_t> > getAllAncestralStates() const {
    if (!computed_) throw Exception("MLAncestralStateReconstruction: computeJointLikelihood() was not called");
    std::map<int, std::vector<size_t> > ancestors;
    const std::vector<int> ids = likelihood_->getNodesId();
    for (size_t n = 0; n < ids.size(); ++n) ancestors[ids[n]] = states_[n];
    return ancestors;
  }
  const Vdouble& getBestJointLogLikelihoodPerSite() const { return bestLogLik_; }   // not in the reference: log max joint likelihood

 private:
  const AbstractHomogeneousTreeLikelihood* likelihood_;   // not owned
  std::vector<std::vector<size_t> > states_;
  Vdouble bestLogLik_;
  bool computed_ = false;
};

// ---- marginal ancestral reconstruction (Likelihood/MarginalAncestralStateReconstruction.{h,cpp}) ---------------------------------
// getAncestralStatesForNode (.cpp:47-102): probs[i][x] = sum_c computeLikelihoodAtNode[i][c][x] r_c / l_i at an internal node --
// the device's marginal posterior table -- and the best (or a sampled) state per distinct site; a leaf gets the first maximum
// of its leaf likelihoods with probability one (:53-65).
class MarginalAncestralStateReconstruction {
 public:
  explicit MarginalAncestralStateReconstruction(const AbstractHomogeneousTreeLikelihood* drl)
      : likelihood_(drl), nbSites_(drl->getNumberOfSites()), nbDistinctSites_(drl->getNumberOfDistinctSites()),
        nbStates_(drl->getNumberOfStates()) {}
  std::vector<size_t> getAncestralStatesForNode(int nodeId, VVdouble& probs, bool sample = false) const {
    std::vector<size_t> ancestors(nbDistinctSites_, 0);
    if (likelihood_->getTree().getNode(nodeId)->isLeaf()) {
      const VVVdouble leaf = likelihood_->getPosteriorProbabilitiesForEachStateForEachRate(nodeId);   // leaf likelihoods x p_c / sum
      probs.assign(nbDistinctSites_, Vdouble(nbStates_, 0.0));
      for (size_t i = 0; i < nbDistinctSites_; ++i) {
        size_t j = 0;
        for (size_t x = 1; x < nbStates_; ++x) if (leaf[i][0][x] > leaf[i][0][j]) j = x;
        ancestors[i] = j;
        probs[i][j] = 1.0;
      }
      return ancestors;
    }
    likelihood_->getMarginalPosteriors(nodeId, probs, nullptr);
    for (size_t i = 0; i < nbDistinctSites_; ++i) {
      if (sample) {
        const double r = std::generate_canonical<double, 53>(rng_);
        double cum = 0;
        for (size_t j = 0; j < nbStates_; ++j) {
          cum += probs[i][j];
          if (r <= cum) { ancestors[i] = j; break; }
        }
      } else {
        ancestors[i] = (size_t)(std::max_element(probs[i].begin(), probs[i].end()) - probs[i].begin());
      }
    }
    return ancestors;
  }
  std::vector<size_t> getAncestralStatesForNode(int nodeId) const {
    VVdouble probs;
    return getAncestralStatesForNode(nodeId, probs, false);
  }
  // one state per distinct site for every node of the tree (recursiveMarginalAncestralStates)
  std::map<int, std::vector<size_t> > getAllAncestralStates() const {
    std::map<int, std::vector<size_t> > ancestors;
    for (int id : likelihood_->getNodesId()) ancestors[id] = getAncestralStatesForNode(id);
    return ancestors;
  }
  // the node's states site by site (getAncestralSequenceForNode without the Sequence wrapper): site -> pattern via getSiteIndex
  std::vector<size_t> getAncestralStatesPerSiteForNode(int nodeId, VVdouble* probs = nullptr, bool sample = false) const {
    VVdouble patterned;
    const std::vector<size_t> states = getAncestralStatesForNode(nodeId, patterned, sample);
    std::vector<size_t> all(nbSites_);
    if (probs) probs->resize(nbSites_);
    for (size_t i = 0; i < nbSites_; ++i) {
      all[i] = states[likelihood_->getSiteIndex(i)];
      if (probs) (*probs)[i] = patterned[likelihood_->getSiteIndex(i)];
    }
    return all;
  }

 private:
  const AbstractHomogeneousTreeLikelihood* likelihood_;   // not owned
  size_t nbSites_, nbDistinctSites_, nbStates_;
  mutable std::mt19937_64 rng_{20260101};
};

// ---- marginal ancestral reconstruction for non-reversible models (fork) --------------------------------------------------------
// Likelihood/MarginalNonRevAncestralStateReconstruction.h:66-207.  The reference re-runs the prefix pass once per root state
// (DRNonHomogeneousTreeLikelihood::computeLikelihoodPrefixConditionalOnRoot, .cpp:1026-1162: S full passes, S^2 work per
// (node, site, state) on top); the device computes the same tables in one pass per node from the resident arrays
// (bppgpu_get_marginal_posteriors).  Map keys follow the reference: node id -> distinct-site index -> state vector.
class MarginalNonRevAncestralStateReconstruction {
 public:
  explicit MarginalNonRevAncestralStateReconstruction(AbstractHomogeneousTreeLikelihood* drl)
      : likelihood_(drl), nbSites_(drl->getNumberOfSites()), nbDistinctSites_(drl->getNumberOfDistinctSites()),
        nbClasses_(drl->getNumberOfClasses()), nbStates_(drl->getNumberOfStates()) {}

  void computePosteriorProbabilitiesOfNodesForEachStatePerSite() {
    postProbNode_.reset(new std::map<int, std::map<size_t, std::vector<double> > >);
    jointProbabilities_.reset(new std::map<int, std::map<size_t, VVdouble> >);
    for (int id : likelihood_->getNodesId()) {
      VVdouble post;
      VVVdouble joint;
      likelihood_->getMarginalPosteriors(id, post, &joint);
      for (size_t i = 0; i < nbDistinctSites_; ++i) {
        (*postProbNode_)[id][i] = post[i];
        (*jointProbabilities_)[id][i] = joint[i];   // [nodeState][fatherState]; all zero at the root, like the reference
      }
    }
  }
  std::map<int, std::map<size_t, VVdouble> > getAllJointFatherNodeProbabilities() {
    if (!jointProbabilities_) computePosteriorProbabilitiesOfNodesForEachStatePerSite();
    return *jointProbabilities_;
  }
  std::map<int, std::map<size_t, std::vector<double> > >* getPosteriorProbForAllNodesAndStatesPerSite() {
    if (!postProbNode_) computePosteriorProbabilitiesOfNodesForEachStatePerSite();
    return postProbNode_.get();
  }
  // argmax state per SITE (.cpp:155-167).  The reference indexes the distinct-site tables with the site number, which is
  // only right when every site is its own pattern (ChromEvol: one site); here a site reads its pattern's entry.
  const std::map<int, std::vector<size_t> > getAllAncestralStates() const {
    if (!postProbNode_) throw Exception("MarginalNonRevAncestralStateReconstruction: posterior probabilities not computed");
    std::map<int, std::vector<size_t> > ancestors;
    for (const auto& kv : *postProbNode_) {
      std::vector<size_t>& a = ancestors[kv.first];
      a.reserve(nbSites_);
      for (size_t s = 0; s < nbSites_; ++s) {
        const std::vector<double>& p = kv.second.at(likelihood_->getSiteIndex(s));
        a.push_back((size_t)(std::max_element(p.begin(), p.end()) - p.begin()));   // VectorTools::whichMax: first maximum
      }
    }
    return ancestors;
  }
  // posterior of the root state at distinct site 0 (.cpp:139-153)
  std::vector<double> getRootPosteriorProb() const {
    VVdouble post;
    likelihood_->getMarginalPosteriors(likelihood_->getNodesId().back(), post, nullptr);
    return post.at(0);
  }

 private:
  AbstractHomogeneousTreeLikelihood* likelihood_;   // not owned
  size_t nbSites_, nbDistinctSites_, nbClasses_, nbStates_;
  std::unique_ptr<std::map<int, std::map<size_t, std::vector<double> > > > postProbNode_;
  std::unique_ptr<std::map<int, std::map<size_t, VVdouble> > > jointProbabilities_;
};

// Likelihood/RNonHomogeneousMixedTreeLikelihood.{h,cpp}: mixed models on groups of branches.  The reference expands the set into one
// RNonHomogeneousTreeLikelihood per site path ("hyper-node") and adds their site likelihoods with the path probabilities
// (RNonHomogeneousMixedTreeLikelihood.cpp: getLikelihoodForASite = sum_paths p_path L_path); here every path is one point of a
// single device object: K x M model slots (path k, model m), one branch -> slot map per point, one evaluation for all paths.
class RNonHomogeneousMixedTreeLikelihood : public AbstractHomogeneousTreeLikelihood {
 public:
  RNonHomogeneousMixedTreeLikelihood(const Tree& tree, const VectorSiteContainer& data, MixedSubstitutionModelSet* modelSet,
                                     DiscreteDistribution* rDist, bool verbose = true, bool usePatterns = true, int device = 0)
      : AbstractHomogeneousTreeLikelihood(tree, modelSet->getModel(0)->getNModel(0), rDist, false, BPPGPU_FLAG_R_SEMANTICS, device),
        mixedSet_(modelSet) {
    (void)verbose; (void)usePatterns;
    for (size_t i = 0; i + 1 < nodes_.size(); ++i) mixedSet_->getModelIndexForNode(nodes_[i]->getId());   // throws if a branch has no model
    nPoints_ = (int)mixedSet_->getNumberOfPaths();
    nModelSlots_ = nPoints_ * (int)mixedSet_->getNumberOfModels();
    computeDerivatives_ = false;
    setData(data);
  }
  double getValue() const { requireInit(); return mixedMinusLogLik_; }
  double getLogLikelihood() const { return -getValue(); }
  double getLogLikelihoodForASite(size_t site) const { requireInit(); return mixedSiteLnl_[(size_t)siteIndex_[site]]; }
  void computeTreeLikelihood() { fireParameterChanged(); }

 protected:
  // "<Model>.<param>_<m>": parameter of the m-th mixed model of the set (1-based, like SubstitutionModelSet)
  bool applyBranchParameter(const std::string& name, double value) override {
    const size_t u = name.rfind('_');
    if (u == std::string::npos || name.compare(0, 5, "BrLen") == 0) return false;
    char* end = nullptr;
    const long m = std::strtol(name.c_str() + u + 1, &end, 10);
    if (*end != 0 || m < 1 || (size_t)m > mixedSet_->getNumberOfModels()) return false;
    mixedSet_->getModel((size_t)m - 1)->setParameterValue(name.substr(0, u), value);
    return true;
  }
  void uploadModel() override {
    if (!engine_) return;
    const size_t K = mixedSet_->getNumberOfPaths(), M = mixedSet_->getNumberOfModels();
    Vdouble r(rDist_->getNumberOfCategories()), p(r.size());
    for (size_t c = 0; c < r.size(); ++c) { r[c] = rDist_->getCategory(c); p[c] = rDist_->getProbability(c); }
    check(bppgpu_set_rates(engine_, r.data(), p.data()), "setRates");
    for (size_t k = 0; k < K; ++k) {
      for (size_t m = 0; m < M; ++m) {
        bppgpu_model_desc d;
        mixedSet_->getModel(m)->getNModel(k)->fillModelDesc(d);
        check(bppgpu_set_model(engine_, (int32_t)(k * M + m), &d), "setModel");
      }
      std::vector<int32_t> slot(nodes_.size(), 0);
      for (size_t i = 0; i + 1 < nodes_.size(); ++i) slot[i] = (int32_t)(k * M + mixedSet_->getModelIndexForNode(nodes_[i]->getId()));
      check(bppgpu_set_branch_models(engine_, (int32_t)k, slot.data()), "setBranchModels");
      // stationary set (nonhomogeneous.stationarity = yes): the root uses the equilibrium frequencies of the models
      const Vdouble f = mixedSet_->getModel(0)->getNModel(k)->getFrequencies();
      check(bppgpu_set_root_freqs(engine_, (int32_t)k, f.data()), "setRootFreqs");
    }
    rootFreqs_ = mixedSet_->getModel(0)->getNModel(0)->getFrequencies();
  }
  void fireParameterChanged() override {
    uploadModel();
    const size_t K = mixedSet_->getNumberOfPaths(), N = (size_t)nPatterns_;
    Vdouble t(nodes_.size(), 0.0);
    for (size_t i = 0; i < brLen_.size(); ++i) t[i] = brLen_[i];
    for (size_t k = 0; k < K; ++k) check(bppgpu_set_branch_lengths(engine_, (int32_t)k, t.data()), "applyParameters");
    Vdouble lnl(K, 0.0);
    check(bppgpu_eval(engine_, BPPGPU_EVAL_LNL, lnl.data(), nullptr, nullptr), "computeTreeLikelihood");
    numOfLikelihoodCalculations_ += (long)K;
    std::vector<Vdouble> sl(K, Vdouble(N));
    for (size_t k = 0; k < K; ++k) check(bppgpu_get_site_lnl(engine_, (int32_t)k, sl[k].data()), "getLogLikelihoodForEachSite");
    mixedSiteLnl_.assign(N, 0.0);
    for (size_t i = 0; i < N; ++i) {
      double mx = -std::numeric_limits<double>::infinity();
      for (size_t k = 0; k < K; ++k) if (mixedSet_->getPathProbability(k) > 0) mx = std::max(mx, sl[k][i]);
      double sum = 0;
      for (size_t k = 0; k < K; ++k) if (mixedSet_->getPathProbability(k) > 0) sum += mixedSet_->getPathProbability(k) * std::exp(sl[k][i] - mx);
      mixedSiteLnl_[i] = std::isfinite(mx) ? mx + std::log(sum) : mx;
    }
    Vdouble la(siteIndex_.size());   // getLogLikelihood (RHomogeneousTreeLikelihood.cpp:162-176): every site, sorted, summed
    for (size_t j = 0; j < la.size(); ++j) la[j] = mixedSiteLnl_[(size_t)siteIndex_[j]];
    std::sort(la.begin(), la.end());
    double ll = 0;
    for (size_t j = la.size(); j > 0; --j) ll += la[j - 1];
    mixedMinusLogLik_ = -ll;
    minusLogLik_ = mixedMinusLogLik_;
    siteLnl_ = mixedSiteLnl_;
    derivsValid_ = false;
    rootArraysValid_ = false;
  }

 private:
  MixedSubstitutionModelSet* mixedSet_;   // not owned
  Vdouble mixedSiteLnl_;
  double mixedMinusLogLik_ = 0;
};

// ---- batched line searches for the multi-start optimiser ------------------------------------------------------------
// ChromosomeNumberOptimizer::optimizeModelParametersOneDimension (Likelihood/ChromosomeNumberOptimizer.cpp:440-530) runs, for every
// starting point in turn and every parameter in turn, a bounded Brent search whose every probe is a full likelihood evaluation.
// The probes of DIFFERENT points are independent, so here all points search the same parameter in lockstep: one probe per point
// per step, ALL of them evaluated by one device call.  The search itself is Brent's bounded minimiser (golden section + parabolic
// interpolation on [lower, upper]; Forsythe, Malcolm & Moler's fmin) written as a per-point state machine; bpp-core's
// BrentOneDimension is not in the reference tree, so probe sequences are not claimed to match it -- optima are.
//
// F: size(), setParameter(point, name, value), evaluate() [one batched evaluation], value(point), parameter(point, name).
template <class F>
class BatchedBrent {
 public:
  explicit BatchedBrent(F* f) : f_(f), nbBatchEvaluations_(0) {}
  // minimise over `name` in [lower, upper] for every point with active[k]; a point keeps its current value if the search found
  // nothing better.  Returns the number of batched evaluations.
  unsigned search(const std::string& name, double lower, double upper, double tol, const std::vector<char>& active, unsigned maxSteps = 200) {
    const size_t K = f_->size();
    const double c = 0.5 * (3.0 - std::sqrt(5.0)), eps = std::sqrt(std::numeric_limits<double>::epsilon());
    std::vector<State> st(K);
    unsigned evals = 0;
    for (size_t k = 0; k < K; ++k) {
      State& s = st[k];
      s.done = !active[k];
      s.x0 = f_->parameter(k, name);
      s.f0 = f_->value(k);
      s.a = lower; s.b = upper;
      s.v = s.w = s.x = s.a + c * (s.b - s.a);
      s.e = s.d = 0;
      s.u = s.x;
      if (!s.done) f_->setParameter(k, name, s.u);
    }
    f_->evaluate(); ++evals;
    for (size_t k = 0; k < K; ++k) if (!st[k].done) st[k].fx = st[k].fv = st[k].fw = f_->value(k);
    for (unsigned step = 0; step < maxSteps; ++step) {
      bool any = false;
      for (size_t k = 0; k < K; ++k) {
        State& s = st[k];
        if (s.done) continue;
        const double xm = 0.5 * (s.a + s.b), tol1 = eps * std::fabs(s.x) + tol / 3.0, tol2 = 2.0 * tol1;
        if (std::fabs(s.x - xm) <= tol2 - 0.5 * (s.b - s.a)) { s.done = true; continue; }
        bool golden = true;
        if (std::fabs(s.e) > tol1) {   // parabolic step through x, w, v
          double r = (s.x - s.w) * (s.fx - s.fv), q = (s.x - s.v) * (s.fx - s.fw), pp = (s.x - s.v) * q - (s.x - s.w) * r;
          q = 2.0 * (q - r);
          if (q > 0) pp = -pp;
          q = std::fabs(q);
          r = s.e;
          s.e = s.d;
          if (std::fabs(pp) < std::fabs(0.5 * q * r) && pp > q * (s.a - s.x) && pp < q * (s.b - s.x)) {
            s.d = pp / q;
            const double u = s.x + s.d;
            if (u - s.a < tol2 || s.b - u < tol2) s.d = xm >= s.x ? tol1 : -tol1;
            golden = false;
          }
        }
        if (golden) {
          s.e = s.x >= xm ? s.a - s.x : s.b - s.x;
          s.d = c * s.e;
        }
        s.u = std::fabs(s.d) >= tol1 ? s.x + s.d : s.x + (s.d >= 0 ? tol1 : -tol1);
        f_->setParameter(k, name, s.u);
        any = true;
      }
      if (!any) break;
      f_->evaluate(); ++evals;
      for (size_t k = 0; k < K; ++k) {
        State& s = st[k];
        if (s.done) continue;
        const double fu = f_->value(k);
        if (fu <= s.fx) {
          if (s.u >= s.x) s.a = s.x; else s.b = s.x;
          s.v = s.w; s.fv = s.fw; s.w = s.x; s.fw = s.fx; s.x = s.u; s.fx = fu;
        } else {
          if (s.u < s.x) s.a = s.u; else s.b = s.u;
          if (fu <= s.fw || s.w == s.x) { s.v = s.w; s.fv = s.fw; s.w = s.u; s.fw = fu; }
          else if (fu <= s.fv || s.v == s.x || s.v == s.w) { s.v = s.u; s.fv = fu; }
        }
      }
    }
    // settle every searched point on the better of (start, best found); one more batched evaluation brings values() in line
    for (size_t k = 0; k < K; ++k)
      if (active[k]) f_->setParameter(k, name, st[k].fx < st[k].f0 ? st[k].x : st[k].x0);
    f_->evaluate(); ++evals;
    nbBatchEvaluations_ += evals;
    return evals;
  }
  unsigned getNumberOfBatchEvaluations() const { return nbBatchEvaluations_; }

 private:
  struct State {
    bool done;
    double a, b, x, w, v, fx, fw, fv, d, e, u, x0, f0;
  };
  F* f_;
  unsigned nbBatchEvaluations_;
};

// The multi-start driver of ChromosomeNumberOptimizer::optimize (Likelihood/ChromosomeNumberOptimizer.cpp:115-153) on one device
// object: cycle i keeps the numOfPoints[i] best points (the reference sorts its vector of likelihoods, :150) and runs
// numOfIterations[i] rounds of per-parameter line searches (optimizeModelParametersOneDimension), stopping a round early when no
// point improved by more than the tolerance (:520-522, applied to the best change over the points).  Parameters are searched on
// (lowerBoundOfRateParam + 1e-10, upperBoundOfRateParam] = (0, 100] (ChromosomeSubstitutionModel.h:15-18; :505).
class ChromosomeNumberOptimizer {
 public:
  ChromosomeNumberOptimizer(const Tree& tree, const VectorSiteContainer& data, const std::vector<ChromosomeSubstitutionModel*>& startPoints,
                            DiscreteDistribution* rDist, bool weightedRootFreq = true, int device = 0)
      : models_(startPoints), batch_(tree, data, weightedRootFreq, std::vector<SubstitutionModel*>(startPoints.begin(), startPoints.end()), rDist, device),
        adaptor_(this), brent_(&adaptor_), active_(startPoints.size(), 1), order_(startPoints.size()) {
    for (size_t k = 0; k < order_.size(); ++k) order_[k] = k;
    batch_.initialize();
  }
  void setParameterNames(const std::vector<std::string>& names) { paramNames_ = names; }
  // numOfPoints / numOfIterations as in ChromEvolOptions (e.g. {10, 3, 1} and {0, 2, 5})
  void optimize(const std::vector<unsigned>& numOfPoints, const std::vector<unsigned>& numOfIterations, double tol,
                double lower = 1e-10, double upper = 100.0) {
    if (paramNames_.empty()) paramNames_ = models_.at(0)->getParameterNames();
    for (size_t i = 0; i < numOfIterations.size(); ++i) {
      sortPoints();
      for (size_t r = 0; r < order_.size(); ++r) active_[order_[r]] = r < numOfPoints[i] ? 1 : 0;   // clearVectorOfLikelihoods (:156-161)
      for (unsigned it = 0; it < numOfIterations[i]; ++it) {
        const Vdouble before = batch_.getValues();
        for (const std::string& name : paramNames_) brent_.search(name, lower, upper, tol, active_);
        double change = 0;
        for (size_t k = 0; k < active_.size(); ++k) if (active_[k]) change = std::max(change, std::fabs(before[k] - batch_.getValue(k)));
        if (change < tol) break;
      }
    }
    sortPoints();
  }
  // points from best to worst (getVectorOfLikelihoods()[0] is the reference's final answer)
  const std::vector<size_t>& getPointOrder() const { return order_; }
  double getValue(size_t point) const { return batch_.getValue(point); }
  double getBestValue() const { return batch_.getValue(order_[0]); }
  ChromosomeSubstitutionModel* getModel(size_t point) const { return models_.at(point); }
  ChromosomeSubstitutionModel* getBestModel() const { return models_.at(order_[0]); }
  unsigned getNumberOfBatchEvaluations() const { return brent_.getNumberOfBatchEvaluations(); }
  LikelihoodPointBatch& getLikelihoods() { return batch_; }

 private:
  struct Adaptor {
    explicit Adaptor(ChromosomeNumberOptimizer* o) : o_(o) {}
    size_t size() const { return o_->models_.size(); }
    double parameter(size_t k, const std::string& name) const { return o_->models_[k]->getParameterValue(name); }
    void setParameter(size_t k, const std::string& name, double v) { o_->models_[k]->setParameterValue(name, v); o_->batch_.markModelChanged(k); }
    void evaluate() { o_->batch_.evaluate(); }
    double value(size_t k) const { return o_->batch_.getValue(k); }
    ChromosomeNumberOptimizer* o_;
  };
  void sortPoints() {
    std::stable_sort(order_.begin(), order_.end(), [&](size_t a, size_t b) {
      if (active_[a] != active_[b]) return active_[a] > active_[b];   // dropped points stay behind the kept ones
      return batch_.getValue(a) < batch_.getValue(b);
    });
  }
  std::vector<ChromosomeSubstitutionModel*> models_;   // not owned
  LikelihoodPointBatch batch_;
  Adaptor adaptor_;
  BatchedBrent<Adaptor> brent_;
  std::vector<char> active_;
  std::vector<size_t> order_;
  std::vector<std::string> paramNames_;
};

// ---- Newton-Raphson on the branch lengths --------------------------------------------------------------------------
// Likelihood/PseudoNewtonOptimizer.cpp:100-193, the optimiser OptimizationTools::optimizeNumericalParameters puts on the branch
// lengths (OptimizationTools.cpp:187-188): every parameter moves by d1 / d2 at once (the other way when d2 < 0, not at all when
// d2 = 0 or the ratio is NaN), and the whole step is halved -- the Felsenstein-Churchill correction -- up to maxCorrection_ times
// while the function is worse than before; stop when |f - f_previous| < tolerance (FunctionStopCondition).  This is the variant
// with disableCG() (PseudoNewtonOptimizer.h:125): no conjugate-gradient detour at the fourth correction.
// One step costs the reference 2 B per-branch derivative passes + the probes; here it is ONE device evaluation with
// BPPGPU_EVAL_D1 | D2 (all B first and second derivatives) + one value evaluation per probe.
class PseudoNewtonOptimizer {
 public:
  explicit PseudoNewtonOptimizer(AbstractHomogeneousTreeLikelihood* function)
      : f_(function), tolerance_(1e-6), maxCorrection_(10), nbEvalMax_(1000000), nbEval_(0), currentValue_(0), previousValue_(0) {}
  void setMaximumNumberOfCorrections(unsigned mx) { maxCorrection_ = mx; }
  void setMaximumNumberOfEvaluations(unsigned n) { nbEvalMax_ = n; }
  void setTolerance(double t) { tolerance_ = t; }
  void disableCG() {}
  void init(const ParameterList& params) {
    params_ = params;
    f_->setParametersValues(params_);
    currentValue_ = f_->getValue();
    previousValue_ = currentValue_;
    nbEval_ = 0;
  }
  double step() {
    const size_t n = params_.size();
    std::vector<double> movements(n);
    ParameterList newPoint = params_;
    for (size_t i = 0; i < n; ++i) {
      const double d1 = f_->getFirstOrderDerivative(params_[i].name), d2 = f_->getSecondOrderDerivative(params_[i].name);
      if (d2 == 0) movements[i] = 0;
      else if (d2 < 0) movements[i] = -d1 / d2;   // "Moving in the other direction" (:121-127)
      else movements[i] = d1 / d2;
      if (std::isnan(movements[i])) movements[i] = 0;
      newPoint[i].value = params_[i].value - movements[i];
    }
    double newValue = probe(newPoint, &movements);
    unsigned count = 0;
    while (count < maxCorrection_ && (newValue > currentValue_ + tolerance_ || std::isnan(newValue))) {
      for (size_t i = 0; i < n; ++i) {
        movements[i] /= 2;
        newPoint[i].value = params_[i].value - movements[i];
      }
      newValue = probe(newPoint, nullptr);
      ++count;
    }
    if (newValue > currentValue_ + tolerance_) {   // "Value could not be ameliorated!"
      f_->setParametersValues(params_);
      newValue = currentValue_;
      previousValue_ = currentValue_;
    } else {
      previousValue_ = currentValue_;
      params_ = newPoint;
      currentValue_ = newValue;
    }
    return newValue;
  }
  unsigned optimize() {
    for (;;) {
      step();
      if (std::fabs(currentValue_ - previousValue_) < tolerance_ || nbEval_ >= nbEvalMax_) break;
    }
    return nbEval_;
  }
  double getFunctionValue() const { return currentValue_; }
  unsigned getNumberOfEvaluations() const { return nbEval_; }
  const ParameterList& getParameters() const { return params_; }

 private:
  // f(newPoint) with the function's own constraints applied (AUTO policy): the point actually reached is read back, and the
  // movement corrected to it like :139-141
  double probe(ParameterList& point, std::vector<double>* movements) {
    f_->setParametersValues(point);
    ++nbEval_;
    for (size_t i = 0; i < point.size(); ++i) {
      point[i].value = f_->getParameterValue(point[i].name);
      if (movements) (*movements)[i] = params_[i].value - point[i].value;
    }
    return f_->getValue();
  }
  AbstractHomogeneousTreeLikelihood* f_;
  ParameterList params_;
  double tolerance_;
  unsigned maxCorrection_, nbEvalMax_, nbEval_;
  double currentValue_, previousValue_;
};

}  // namespace bppshim
