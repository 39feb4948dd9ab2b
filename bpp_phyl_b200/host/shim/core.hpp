// bppgpu shim (see ../bppgpu_shim.hpp): exceptions, RowMatrix, NumConstants, host linear algebra and special functions (bpp-core stand-ins)
#pragma once
#include <algorithm>
#include <atomic>
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
#include <thread>
#include <vector>

#include "../../../include/bppgpu.h"

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

// rows / eigenvalues of the S >= 96 models (codon 61, chromosome ~200) are independent pieces of O(S^2) work: spread them over the
// host cores.  Every piece writes its own outputs, so the result does not depend on the thread count (BPPGPU_SHIM_THREADS, default
// min(hardware threads, 16); 1 = serial).
inline int shim_threads() {
  static const int n = [] {
    const char* e = std::getenv("BPPGPU_SHIM_THREADS");
    int v = e ? std::atoi(e) : (int)std::thread::hardware_concurrency();
    return std::max(1, std::min(v, 16));
  }();
  return n;
}
template <class F>
inline void parallel_for(int n, int min_parallel, F&& body) {
  const int T = n >= min_parallel ? std::min(shim_threads(), n) : 1;
  if (T <= 1) { for (int i = 0; i < n; ++i) body(i); return; }
  std::atomic<int> next(0);   // pieces are handed out one at a time: their costs differ (conjugate pairs, sparse rows)
  auto work = [&] { for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) body(i); };
  std::vector<std::thread> th;
  for (int t = 1; t < T; ++t) th.emplace_back(work);
  work();
  for (auto& x : th) x.join();
}

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
    parallel_for(n, 96, [&](int i) {
      for (int k = 0; k < n; ++k) {
        const double aik = A[(size_t)i * n + k];
        if (aik == 0.0) continue;
        for (int j = 0; j < n; ++j) R[(size_t)i * n + j] -= aik * inv[(size_t)k * n + j];
      }
      R[(size_t)i * n + i] += 1.0;
    });
    std::vector<double> X(inv);
    parallel_for(n, 96, [&](int i) {
      for (int k = 0; k < n; ++k) {
        const double xik = inv[(size_t)i * n + k];
        if (xik == 0.0) continue;
        for (int j = 0; j < n; ++j) X[(size_t)i * n + j] += xik * R[(size_t)k * n + j];
      }
    });
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

// complex LU of (A - lambda I) with partial pivoting, factored ONCE per eigenvalue and reused by every inverse-iteration
// solve (O(n^3) per eigenvalue at worst, far less on the banded generators of the chromosome models: zero multipliers are
// skipped and remembered)
struct ShiftedLU {
  typedef std::complex<double> cd;
  int n = 0;
  std::vector<cd> a;          // U above / on the diagonal, the multipliers of L below it
  std::vector<int> piv;       // row swapped with `col` at step col
  std::vector<int> nz_ptr, nz_row;   // rows with a non-zero multiplier, per column
  void factor(const std::vector<double>& A, int n_, cd lambda) {
    n = n_;
    a.resize((size_t)n * n);
    piv.resize(n);
    nz_ptr.assign(1, 0);
    nz_row.clear();
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) a[(size_t)i * n + j] = cd(A[(size_t)i * n + j]) - (i == j ? lambda : cd(0));
    for (int col = 0; col < n; ++col) {
      int p = col;
      double best = std::norm(a[(size_t)col * n + col]);
      for (int r = col + 1; r < n; ++r) {
        const double v = std::norm(a[(size_t)r * n + col]);
        if (v > best) { best = v; p = r; }
      }
      if (best < 1e-300) a[(size_t)p * n + col] = cd(1e-300);
      piv[col] = p;
      if (p != col)   // the multipliers already stored left of `col` stay with their rows: solve() replays swap and elimination in step
        for (int k = col; k < n; ++k) std::swap(a[(size_t)p * n + k], a[(size_t)col * n + k]);
      const cd d = cd(1.0) / a[(size_t)col * n + col];
      for (int r = col + 1; r < n; ++r) {
        const cd f = a[(size_t)r * n + col] * d;
        a[(size_t)r * n + col] = f;
        if (f == cd(0)) continue;
        nz_row.push_back(r);
        for (int k = col + 1; k < n; ++k) a[(size_t)r * n + k] -= f * a[(size_t)col * n + k];
      }
      nz_ptr.push_back((int)nz_row.size());
    }
  }
  void solve(std::vector<cd>& x) const {
    for (int col = 0; col < n; ++col) {
      if (piv[col] != col) std::swap(x[piv[col]], x[col]);
      const cd xc = x[col];
      if (xc == cd(0)) continue;
      for (int t = nz_ptr[col]; t < nz_ptr[col + 1]; ++t) x[nz_row[t]] -= a[(size_t)nz_row[t] * n + col] * xc;
    }
    for (int i = n - 1; i >= 0; --i) {
      cd s = x[i];
      const cd* row = &a[(size_t)i * n];
      for (int k = i + 1; k < n; ++k) s -= row[k] * x[k];
      x[i] = s / row[i];
    }
  }
};

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
  std::vector<char> failed(n, 0);
  const std::vector<double> re0(re), im0(im);   // the QR values: what every piece reads (pieces write re / im of their own pair)
  parallel_for(n, 96, [&](int k) {
    if (im0[k] < 0.0) return;  // second member of a pair: filled with the first
    const cd lam(re0[k], im0[k]);
    // inverse iteration; the eigenvalue itself is refined from the iteration (lambda = shift + <x,x>/<x,y> with
    // (A - shift I) y = x), which recovers the digits the unbalanced QR iteration loses on non-normal generators
    cd lamk = lam;
    std::vector<cd> x(n);
    for (int i = 0; i < n; ++i) x[i] = cd(1.0 + 0.37 * ((i * 7919 + k * 104729) % 101) / 101.0, 0.0);
    // two rounds: the first with the QR eigenvalue as the shift, the second with the refined one; each factors once and
    // iterates on the factors (a shift a few ulps off the eigenvalue keeps (A - shift I) numerically invertible)
    // the factor's n x n complex workspace lives as long as the thread: a fresh 640 KB allocation per eigenvalue is an mmap /
    // page-fault / munmap round trip each, and those serialise the worker threads on the process's address-space lock
    // (measured: 8 threads took as long as one)
    static thread_local ShiftedLU lu;
    for (int round = 0; round < 2; ++round) {
      const double eps = std::max(std::abs(lamk), scale * 1e-3) * 4e-15 * (1.0 + (k % 7));
      const cd shifted = lamk + cd(eps, im0[k] != 0.0 ? eps : 0.0);
      lu.factor(A, n, shifted);
      cd lnew = lamk;
      for (int it = 0; it < (round == 0 ? 3 : 2); ++it) {
        static thread_local std::vector<cd> y;
        y = x;
        lu.solve(y);
        cd xy(0), xx(0);
        for (int i = 0; i < n; ++i) { xy += std::conj(x[i]) * y[i]; xx += std::conj(x[i]) * x[i]; }
        double nrm = 0.0;
        for (auto& v : y) nrm = std::max(nrm, std::abs(v));
        if (!(nrm > 0.0) || !std::isfinite(nrm)) { failed[k] = 1; return; }
        for (int i = 0; i < n; ++i) x[i] = y[i] / nrm;
        if (it >= 1 && std::abs(xy) > 0.0) {
          cd l2 = shifted + xx / xy;
          if (im0[k] == 0.0) l2 = cd(l2.real(), 0.0);
          if (std::abs(l2 - lam) <= 1e-6 * std::max(std::abs(lam), scale)) lnew = l2;  // stay on this eigenvalue
        }
      }
      lamk = lnew;
    }
    re[k] = lamk.real();
    if (im0[k] != 0.0) { im[k] = lamk.imag(); re[k + 1] = lamk.real(); im[k + 1] = -lamk.imag(); }
    if (im0[k] == 0.0) {
      // rotate to a real vector
      cd ph(0);
      double best = 0;
      for (auto& v : x)
        if (std::abs(v) > best) { best = std::abs(v); ph = v; }
      for (int i = 0; i < n; ++i) V[(size_t)i * n + k] = (x[i] / ph).real();
    } else {
      if (k + 1 >= n) { failed[k] = 1; return; }
      for (int i = 0; i < n; ++i) {
        V[(size_t)i * n + k] = x[i].real();
        V[(size_t)i * n + k + 1] = x[i].imag();
      }
    }
  });
  for (char f : failed)
    if (f) return false;
  return true;
}

inline std::vector<double> matmul(const std::vector<double>& A, const std::vector<double>& B, int n) {
  std::vector<double> C((size_t)n * n, 0.0);
  parallel_for(n, 96, [&](int i) {
    for (int k = 0; k < n; ++k) {
      const double a = A[(size_t)i * n + k];
      if (a == 0.0) continue;
      for (int j = 0; j < n; ++j) C[(size_t)i * n + j] += a * B[(size_t)k * n + j];
    }
  });
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

}  // namespace bppshim
