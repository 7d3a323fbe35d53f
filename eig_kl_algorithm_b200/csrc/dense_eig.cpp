// dense_eig.cpp -- eigen-decomposition of the small projected matrix T (ncv x ncv, ncv <= a few
// hundred) on the host, as the north star prescribes ("solves the small tridiagonal eigenproblem on
// the host").  After a thick restart T is an arrowhead block plus a tridiagonal tail, so the general
// symmetric path is used: Householder reduction to tridiagonal form, then implicit-shift QL with
// accumulated transformations.  No LAPACK/Eigen exists in the image; this is self-contained.
#include "internal.h"
#include <algorithm>
#include <cmath>
#include <numeric>
#include <vector>

namespace eigkl {

namespace {

// a: n x n row-major symmetric.  On exit a holds the orthogonal Q with Q^T A Q = tridiag(d, e),
// e[i] couples i-1 and i (e[0] = 0).
void householder_tridiagonalize(int n, std::vector<double> &a, std::vector<double> &d, std::vector<double> &e) {
  auto A = [&](int i, int j) -> double & { return a[(size_t)i * n + j]; };
  d.assign(n, 0.0);
  e.assign(n, 0.0);
  for (int i = n - 1; i >= 1; --i) {
    const int l = i - 1;
    double h = 0.0, scale = 0.0;
    if (l > 0) {
      // a row that is already in tridiagonal form (nothing left of the sub-diagonal) needs no reflector: after a
      // thick restart only the arrow row and the block above it do, so the reduction costs O(keep^3), not O(n^3)
      for (int k = 0; k < l; ++k) scale += std::fabs(A(i, k));
      if (scale == 0.0) {
        e[i] = A(i, l);
      } else {
        scale += std::fabs(A(i, l));
        for (int k = 0; k <= l; ++k) { A(i, k) /= scale; h += A(i, k) * A(i, k); }
        double f = A(i, l);
        double g = f >= 0.0 ? -std::sqrt(h) : std::sqrt(h);
        e[i] = scale * g;
        h -= f * g;
        A(i, l) = f - g;
        f = 0.0;
        for (int j = 0; j <= l; ++j) {
          A(j, i) = A(i, j) / h;
          g = 0.0;
          for (int k = 0; k <= j; ++k) g += A(j, k) * A(i, k);
          for (int k = j + 1; k <= l; ++k) g += A(k, j) * A(i, k);
          e[j] = g / h;
          f += e[j] * A(i, j);
        }
        const double hh = f / (h + h);
        for (int j = 0; j <= l; ++j) {
          f = A(i, j);
          e[j] = g = e[j] - hh * f;
          for (int k = 0; k <= j; ++k) A(j, k) -= (f * e[k] + g * A(i, k));
        }
      }
    } else {
      e[i] = A(i, l);
    }
    d[i] = h;
  }
  d[0] = 0.0;
  e[0] = 0.0;
  for (int i = 0; i < n; ++i) {          // accumulate the transformation
    const int l = i - 1;
    if (d[i] != 0.0) {
      for (int j = 0; j <= l; ++j) {
        double g = 0.0;
        for (int k = 0; k <= l; ++k) g += A(i, k) * A(k, j);
        for (int k = 0; k <= l; ++k) A(k, j) -= g * A(k, i);
      }
    }
    d[i] = A(i, i);
    A(i, i) = 1.0;
    for (int j = 0; j <= l; ++j) A(j, i) = A(i, j) = 0.0;
  }
}

// implicit QL on tridiag(d, e) accumulating into z (n x n row-major, columns = eigenvectors)
bool ql_implicit(int n, std::vector<double> &d, std::vector<double> &e, std::vector<double> &z) {
  auto Z = [&](int i, int j) -> double & { return z[(size_t)i * n + j]; };
  for (int i = 1; i < n; ++i) e[i - 1] = e[i];
  e[n - 1] = 0.0;
  for (int l = 0; l < n; ++l) {
    int iter = 0, m;
    do {
      for (m = l; m < n - 1; ++m) {
        const double dd = std::fabs(d[m]) + std::fabs(d[m + 1]);
        if (std::fabs(e[m]) <= 2.220446049250313e-16 * dd) break;
      }
      if (m != l) {
        if (iter++ == 300) return false;
        double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
        double r = std::hypot(g, 1.0);
        g = d[m] - d[l] + e[l] / (g + (g >= 0.0 ? std::fabs(r) : -std::fabs(r)));
        double s = 1.0, c = 1.0, p = 0.0;
        int i;
        for (i = m - 1; i >= l; --i) {
          double f = s * e[i];
          const double b = c * e[i];
          e[i + 1] = (r = std::hypot(f, g));
          if (r == 0.0) {
            d[i + 1] -= p;
            e[m] = 0.0;
            break;
          }
          s = f / r;
          c = g / r;
          g = d[i + 1] - p;
          r = (d[i] - g) * s + 2.0 * c * b;
          d[i + 1] = g + (p = s * r);
          g = c * r - b;
          for (int k = 0; k < n; ++k) {
            f = Z(k, i + 1);
            Z(k, i + 1) = s * Z(k, i) + c * f;
            Z(k, i) = c * Z(k, i) - s * f;
          }
        }
        if (r == 0.0 && i >= l) continue;
        d[l] -= p;
        e[l] = g;
        e[m] = 0.0;
      }
    } while (m != l);
  }
  return true;
}

}  // namespace

// a (n x n row-major, symmetric) -> eigenvectors in the columns of a, eigenvalues ascending in evals
void sym_eig(int n, double *a_io, double *evals) {
  std::vector<double> a(a_io, a_io + (size_t)n * n), d, e;
  householder_tridiagonalize(n, a, d, e);
  if (!ql_implicit(n, d, e, a)) throw Error(EIGKL_E_NOCONV, "dense eigen-solver: QL did not converge");
  std::vector<int> idx(n);
  std::iota(idx.begin(), idx.end(), 0);
  std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return d[x] < d[y]; });
  for (int j = 0; j < n; ++j) {
    evals[j] = d[idx[j]];
    for (int i = 0; i < n; ++i) a_io[(size_t)i * n + j] = a[(size_t)i * n + idx[j]];
  }
}

}  // namespace eigkl

// ---------------------------------------------------------------------------------------------------
// Symmetric TRIDIAGONAL matrices (the first Lanczos cycle, before any thick restart): the k largest
// eigenpairs in O(n k) -- bisection on the Sturm sequence for the values, inverse iteration with a
// pivoted tridiagonal LU for the vectors.  This keeps the per-check host cost at tens of microseconds,
// so convergence can be tested every few Lanczos steps instead of once per 100-step cycle.
// ---------------------------------------------------------------------------------------------------
namespace eigkl {

namespace {

int sturm_count_below(int n, const double *d, const double *e, double x, double tiny) {
  int cnt = 0;
  double q = d[0] - x;
  if (q < 0) ++cnt;
  for (int i = 1; i < n; ++i) {
    if (std::fabs(q) < tiny) q = (q < 0 ? -tiny : tiny);
    q = d[i] - x - e[i - 1] * e[i - 1] / q;
    if (q < 0) ++cnt;
  }
  return cnt;
}

// solves (T - shift I) y = rhs in place (rhs -> y); LAPACK dgttrf/dgtts2 scheme with partial pivoting
void tridiag_shifted_solve(int n, const double *d, const double *e, double shift, double tiny, std::vector<double> &y) {
  std::vector<double> dl(n > 1 ? n - 1 : 1), dd(n), du(n > 1 ? n - 1 : 1), du2(n > 2 ? n - 2 : 1);
  std::vector<char> piv(n, 0);
  for (int i = 0; i < n; ++i) dd[i] = d[i] - shift;
  for (int i = 0; i + 1 < n; ++i) { dl[i] = e[i]; du[i] = e[i]; }
  for (int i = 0; i + 1 < n; ++i) {
    if (std::fabs(dd[i]) >= std::fabs(dl[i])) {
      if (dd[i] == 0.0) dd[i] = tiny;
      const double f = dl[i] / dd[i];
      dl[i] = f;
      dd[i + 1] -= f * du[i];
      if (i + 2 < n) du2[i] = 0.0;
    } else {
      const double f = dd[i] / dl[i];
      dd[i] = dl[i];
      dl[i] = f;
      const double t = du[i];
      du[i] = dd[i + 1];
      dd[i + 1] = t - f * dd[i + 1];
      if (i + 2 < n) { du2[i] = du[i + 1]; du[i + 1] = -f * du[i + 1]; }
      piv[i] = 1;
    }
  }
  if (dd[n - 1] == 0.0) dd[n - 1] = tiny;
  for (int i = 0; i + 1 < n; ++i) {
    if (!piv[i]) y[i + 1] -= dl[i] * y[i];
    else { const double t = y[i]; y[i] = y[i + 1]; y[i + 1] = t - dl[i] * y[i]; }
  }
  y[n - 1] /= dd[n - 1];
  if (n > 1) y[n - 2] = (y[n - 2] - du[n - 2] * y[n - 1]) / dd[n - 2];
  for (int i = n - 3; i >= 0; --i) y[i] = (y[i] - du[i] * y[i + 1] - du2[i] * y[i + 2]) / dd[i];
}

}  // namespace

void tridiag_top_eig(int n, const double *d, const double *e, int k, double *theta, double *Y);

// The k largest eigenpairs of a symmetric matrix (row-major n x n): Householder reduction (cheap for the
// arrowhead + tridiagonal matrices of a thick-restart cycle, see above), bisection + inverse iteration on the
// tridiagonal form for the k pairs, back-transformation.  theta descending, Y column-major n x k.
// This replaces the full QL decomposition (2.4 ms at n = 100, with the GPU idle) by ~0.2 ms.
void sym_top_eig(int n, const double *a_in, int k, double *theta, double *Y) {
  std::vector<double> a(a_in, a_in + (size_t)n * n), d, e;
  householder_tridiagonalize(n, a, d, e);                 // Q in a; e[i] couples i-1 and i
  k = std::min(k, n);
  std::vector<double> Z((size_t)n * k);
  tridiag_top_eig(n, d.data(), e.data() + 1, k, theta, Z.data());
  for (int t = 0; t < k; ++t)
    for (int i = 0; i < n; ++i) {
      double s = 0.0;
      for (int j = 0; j < n; ++j) s += a[(size_t)i * n + j] * Z[(size_t)t * n + j];
      Y[(size_t)t * n + i] = s;
    }
}

// d[0..n), e[0..n-1): the k largest eigenvalues (descending) and unit eigenvectors (Y column-major n x k)
void tridiag_top_eig(int n, const double *d, const double *e, int k, double *theta, double *Y) {
  double lo = d[0], hi = d[0], nrm = 0.0;
  for (int i = 0; i < n; ++i) {
    const double r = (i > 0 ? std::fabs(e[i - 1]) : 0.0) + (i + 1 < n ? std::fabs(e[i]) : 0.0);
    lo = std::min(lo, d[i] - r);
    hi = std::max(hi, d[i] + r);
    nrm = std::max(nrm, std::fabs(d[i]) + r);
  }
  const double tiny = std::max(nrm, 1e-300) * 1e-300 + 2.3e-308 + nrm * 1e-17 * 1e-3;
  std::vector<double> y(n);
  // bisection for all wanted eigenvalues in lockstep: the Sturm recurrences of the different shifts are
  // independent chains, so the inner loop over the shifts pipelines / vectorises (the divisions of ONE chain
  // are a 100-long dependent sequence; 20 values one after the other cost 1.3 ms at n = 100, in lockstep 0.2)
  const int kw = std::min(k, n);
  std::vector<double> lo_t(kw, lo), hi_t(kw, hi), mid_t(kw), q_t(kw), e2(n > 1 ? n - 1 : 1);
  std::vector<int> cnt_t(kw);
  for (int i = 0; i + 1 < n; ++i) e2[i] = e[i] * e[i];
  for (int it = 0; it < 200; ++it) {
    bool any = false;
    for (int t = 0; t < kw; ++t) {
      mid_t[t] = 0.5 * (lo_t[t] + hi_t[t]);
      if (mid_t[t] > lo_t[t] && mid_t[t] < hi_t[t]) any = true;
    }
    if (!any) break;
    for (int t = 0; t < kw; ++t) { q_t[t] = d[0] - mid_t[t]; cnt_t[t] = q_t[t] < 0 ? 1 : 0; }
    for (int i = 1; i < n; ++i) {
      const double di = d[i], ei2 = e2[i - 1];
      for (int t = 0; t < kw; ++t) {
        double q = q_t[t];
        if (std::fabs(q) < tiny) q = (q < 0 ? -tiny : tiny);
        q = di - mid_t[t] - ei2 / q;
        q_t[t] = q;
        cnt_t[t] += q < 0 ? 1 : 0;
      }
    }
    for (int t = 0; t < kw; ++t) {
      if (!(mid_t[t] > lo_t[t] && mid_t[t] < hi_t[t])) continue;       // this interval is already down to one ulp
      if (cnt_t[t] >= n - t) hi_t[t] = mid_t[t]; else lo_t[t] = mid_t[t];   // eigenvalue index n - t (1-based, ascending)
    }
  }
  for (int t = 0; t < k && t < n; ++t) {
    const double th = 0.5 * (lo_t[t] + hi_t[t]);
    theta[t] = th;
    // inverse iteration; the shift is nudged off the eigenvalue so the LU stays finite
    const double shift = th + nrm * 4.4e-16;
    for (int i = 0; i < n; ++i) y[i] = 1.0 + 0.37 * std::sin(1.0 + 1.7 * i + 0.3 * t);
    for (int it = 0; it < 4; ++it) {
      tridiag_shifted_solve(n, d, e, shift, nrm * 1e-30 + 1e-300, y);
      for (int p = 0; p < t; ++p) {               // keep it orthogonal to the vectors already found
        double dot = 0.0;
        for (int i = 0; i < n; ++i) dot += y[i] * Y[(size_t)p * n + i];
        for (int i = 0; i < n; ++i) y[i] -= dot * Y[(size_t)p * n + i];
      }
      double s = 0.0;
      for (int i = 0; i < n; ++i) s += y[i] * y[i];
      s = 1.0 / std::sqrt(s);
      for (int i = 0; i < n; ++i) y[i] *= s;
    }
    for (int i = 0; i < n; ++i) Y[(size_t)t * n + i] = y[i];
  }
}

}  // namespace eigkl
