// asymtx_qr.hpp - TEST INFRASTRUCTURE (host check only; never compiled into libspartacus_b200.so).
//
// The reference's eigen-solver scheme in the reference's operation order, so that the generic solver bodies,
// compiled for the host without FMA contraction, reproduce the no-FMA oracle BIT FOR BIT
// (tests/test_hostcheck.py::test_bit_identical_to_nofma_oracle): that pins scratch indexing, column
// bucketing and chunking of the product's launch plan on the CPU.  The product's device code uses the
// symmetrised cyclic Jacobi solvers instead (csrc/ssb_radtool.cuh, csrc/ssb_layer_math.cuh).
#pragma once
#include "../../spartacus_surface_b200/csrc/ssb_math.cuh"

namespace ssb {

// ---------------------------------------------------------------------------
// Real eigen-decomposition of a real non-symmetric matrix with real spectrum:
// balance -> Householder Hessenberg -> shifted double-QR with accumulation ->
// back substitution (the ASYMTX scheme the reference uses,
// radtool_eigen_decomposition.F90:164-717), closed forms for order 1 and 2
// (:770-816).  `a` (n x n) is destroyed, `v` receives the eigenvectors,
// `wk` needs 2n+1 doubles.  Returns the number of failures (0/1).
// Indexing below is 1-based through the macros, like the algorithm's
// published description, to keep the shifted loop bounds readable.
// ---------------------------------------------------------------------------
SSB_HD inline int eigen_real(int n, double *a, double *eval, double *v, double *wk) {
  const double Tol = SSB_EPS;
  if (n == 1) {
    eval[0] = a[0];
    v[0] = 1.0;
    return 0;
  }
  if (n == 2) {
    const double a11 = a[0], a21 = a[1], a12 = a[2], a22 = a[3];
    int nerr = 0;
    const double disc = (a11 - a22) * (a11 - a22) + 4.0 * a12 * a21;
    if (disc < 0.0) nerr = 1;
    const double mean = 0.5 * (a11 + a22);
    const double h = 0.5 * sqrt(disc);
    if (a11 >= a22) {
      eval[0] = mean + h;
      eval[1] = mean - h;
    } else {
      eval[0] = mean - h;
      eval[1] = mean + h;
    }
    v[0] = 1.0;
    v[3] = 1.0;
    if (a11 == a22 && (a21 == 0.0 || a12 == 0.0)) {
      const double rn = 1.0 / (Tol * fabs(a11) + fabs(a21) + fabs(a12) + fabs(a22));
      v[1] = a21 * rn;
      v[2] = a12 * rn;
    } else {
      v[1] = a21 / (eval[0] - a22);
      v[2] = a12 / (eval[1] - a11);
    }
    return nerr;
  }
#define AB(i, j) a[((i) - 1) + n * ((j) - 1)]
#define EV(i, j) v[((i) - 1) + n * ((j) - 1)]
#define WK(i) wk[(i)]
#define EL(i) eval[(i) - 1]
  const double C1 = 0.4375, C2 = 0.5, C3 = 0.75, C4 = 0.95, C5 = 16.0, C6 = 256.0;
  int nerr = 0;
  bool failed = false;
  for (int j = 1; j <= n; ++j)
    for (int i = 1; i <= n; ++i) EV(i, j) = (i == j) ? 1.0 : 0.0;
  for (int i = 0; i <= 2 * n; ++i) wk[i] = 0.0;
  for (int i = 1; i <= n; ++i) EL(i) = 0.0;
  int lo = 1, hi = n;

  // -- isolate eigenvalues: rows pushed down, columns pushed left ------------
  for (bool again = true; again;) {
    again = false;
    for (int j = hi; j >= 1; --j) {
      double row = 0.0;
      for (int i = 1; i <= hi; ++i)
        if (i != j) row = row + fabs(AB(j, i));
      if (row == 0.0) {
        WK(hi) = j;
        for (int i = 1; i <= hi; ++i) {
          const double t = AB(i, j);
          AB(i, j) = AB(i, hi);
          AB(i, hi) = t;
        }
        for (int i = lo; i <= n; ++i) {
          const double t = AB(j, i);
          AB(j, i) = AB(hi, i);
          AB(hi, i) = t;
        }
        hi = hi - 1;
        again = true;
        break;
      }
    }
  }
  for (bool again = true; again;) {
    again = false;
    for (int j = lo; j <= hi; ++j) {
      double col = 0.0;
      for (int i = lo; i <= hi; ++i)
        if (i != j) col = col + fabs(AB(i, j));
      if (col == 0.0) {
        WK(lo) = j;
        if (j != lo) {
          for (int i = 1; i <= hi; ++i) {
            const double t = AB(i, j);
            AB(i, j) = AB(i, lo);
            AB(i, lo) = t;
          }
          for (int i = lo; i <= n; ++i) {
            const double t = AB(j, i);
            AB(j, i) = AB(lo, i);
            AB(lo, i) = t;
          }
        }
        lo = lo + 1;
        again = true;
        break;
      }
    }
  }

  // -- balance rows lo..hi ---------------------------------------------------
  for (int i = lo; i <= hi; ++i) WK(i) = 1.0;
  for (bool again = true; again;) {
    again = false;
    for (int i = lo; i <= hi; ++i) {
      double col = 0.0, row = 0.0;
      for (int j = lo; j <= hi; ++j)
        if (j != i) {
          col = col + fabs(AB(j, i));
          row = row + fabs(AB(i, j));
        }
      double f = 1.0;
      double g = row / C5;
      const double h = col + row;
      while (col < g) {
        f = f * C5;
        col = col * C6;
      }
      g = row * C5;
      while (col > g) {
        f = f / C5;
        col = col / C6;
      }
      if ((col + row) / f < C4 * h) {
        WK(i) = WK(i) * f;
        again = true;
        for (int j = lo; j <= n; ++j) AB(i, j) = AB(i, j) / f;
        for (int j = 1; j <= hi; ++j) AB(j, i) = AB(j, i) * f;
      }
    }
  }

  // -- Householder reduction to upper Hessenberg form, accumulated in v ------
  if (hi - 1 >= lo + 1) {
    for (int m = lo + 1; m <= hi - 1; ++m) {
      double h = 0.0;
      WK(m + n) = 0.0;
      double scale = 0.0;
      for (int i = m; i <= hi; ++i) scale = scale + fabs(AB(i, m - 1));
      if (scale != 0.0) {
        for (int i = hi; i >= m; --i) {
          WK(i + n) = AB(i, m - 1) / scale;
          h = h + WK(i + n) * WK(i + n);
        }
        const double g = -fsign(sqrt(h), WK(m + n));
        h = h - WK(m + n) * g;
        WK(m + n) = WK(m + n) - g;
        h = 1.0 / h;
        for (int j = m; j <= n; ++j) {
          double f = 0.0;
          for (int i = hi; i >= m; --i) f = f + WK(i + n) * AB(i, j);
          for (int i = m; i <= hi; ++i) AB(i, j) = AB(i, j) - WK(i + n) * f * h;
        }
        for (int i = 1; i <= hi; ++i) {
          double f = 0.0;
          for (int j = hi; j >= m; --j) f = f + WK(j + n) * AB(i, j);
          for (int j = m; j <= hi; ++j) AB(i, j) = AB(i, j) - WK(j + n) * f * h;
        }
        WK(m + n) = scale * WK(m + n);
        AB(m, m - 1) = scale * g;
      }
    }
    for (int m = hi - 2; m >= lo; --m) {
      const int m1 = m + 1, m2 = m + 2;
      double f = AB(m1, m);
      if (f != 0.0) {
        f = f * WK(m1 + n);
        for (int i = m2; i <= hi; ++i) WK(i + n) = AB(i, m);
        if (m1 < hi) {
          for (int j = 1; j <= n; ++j) {
            double g = 0.0;
            for (int i = m1; i <= hi; ++i) g = g + WK(i + n) * EV(i, j);
            g = g / f;
            for (int i = m1; i <= hi; ++i) EV(i, j) = EV(i, j) + g * WK(i + n);
          }
        }
      }
    }
  }

  // -- norm and isolated eigenvalues -----------------------------------------
  double rnorm = 0.0;
  {
    int jstart = 1;
    for (int i = 1; i <= n; ++i) {
      for (int j = jstart; j <= n; ++j) rnorm = rnorm + fabs(AB(i, j));
      jstart = i;
      if (i < lo || i > hi) EL(i) = AB(i, i);
    }
  }

  // -- shifted double-QR iteration -------------------------------------------
  int en = hi;
  double t = 0.0;
  double p = 0, q = 0, r = 0, s = 0, x = 0, y = 0, z = 0, w = 0;
  while (en >= lo && !failed) {
    int iter = 0;
    const int n1 = en - 1, n2 = en - 2;
    for (;;) {
      // look for a single small sub-diagonal element
      int lb = lo;
      for (int i = lo; i <= en; ++i) {
        lb = en + lo - i;
        if (lb == lo) break;
        s = fabs(AB(lb - 1, lb - 1)) + fabs(AB(lb, lb));
        if (s == 0.0) s = rnorm;
        if (fabs(AB(lb, lb - 1)) < Tol * s) break;
      }
      x = AB(en, en);
      if (lb == en) {  // one root found
        AB(en, en) = x + t;
        EL(en) = AB(en, en);
        en = n1;
        break;
      }
      y = AB(n1, n1);
      w = AB(en, n1) * AB(n1, en);
      if (lb == n1) {  // two roots found (forced real)
        p = (y - x) * C2;
        q = p * p + w;
        z = sqrt(fabs(q));
        AB(en, en) = x + t;
        x = AB(en, en);
        AB(n1, n1) = y + t;
        z = p + fsign(z, p);
        EL(n1) = x + z;
        EL(en) = EL(n1);
        if (z != 0.0) EL(en) = x - w / z;
        x = AB(en, n1);
        r = 1.0 / sqrt(x * x + z * z);
        p = x * r;
        q = z * r;
        for (int j = n1; j <= n; ++j) {
          z = AB(n1, j);
          AB(n1, j) = q * z + p * AB(en, j);
          AB(en, j) = q * AB(en, j) - p * z;
        }
        for (int i = 1; i <= en; ++i) {
          z = AB(i, n1);
          AB(i, n1) = q * z + p * AB(i, en);
          AB(i, en) = q * AB(i, en) - p * z;
        }
        for (int i = lo; i <= hi; ++i) {
          z = EV(i, n1);
          EV(i, n1) = q * z + p * EV(i, en);
          EV(i, en) = q * EV(i, en) - p * z;
        }
        en = n2;
        break;
      }
      if (iter == 30) {  // no convergence
        nerr = nerr + 1;
        failed = true;
        break;
      }
      if (iter == 10 || iter == 20) {  // exceptional shift
        t = t + x;
        for (int i = lo; i <= en; ++i) AB(i, i) = AB(i, i) - x;
        s = fabs(AB(en, n1)) + fabs(AB(n1, n2));
        x = C3 * s;
        y = x;
        w = -C1 * s * s;
      }
      iter = iter + 1;
      // look for two consecutive small sub-diagonal elements
      int m = n2;
      for (int j = lb; j <= n2; ++j) {
        m = n2 + lb - j;
        z = AB(m, m);
        r = x - z;
        s = y - z;
        p = (r * s - w) / AB(m + 1, m) + AB(m, m + 1);
        q = AB(m + 1, m + 1) - z - r - s;
        r = AB(m + 2, m + 1);
        s = 1.0 / (fabs(p) + fabs(q) + fabs(r));
        p = p * s;
        q = q * s;
        r = r * s;
        if (m == lb) break;
        const double u = fabs(AB(m, m - 1)) * (fabs(q) + fabs(r));
        const double vv = fabs(p) * (fabs(AB(m - 1, m - 1)) + fabs(z) + fabs(AB(m + 1, m + 1)));
        if (u <= Tol * vv) break;
      }
      AB(m + 2, m) = 0.0;
      for (int j = m + 3; j <= en; ++j) {
        AB(j, j - 2) = 0.0;
        AB(j, j - 3) = 0.0;
      }
      // double QR step on rows lb..en, columns m..en
      for (int k = m; k <= n1; ++k) {
        const bool notlast = (k != n1);
        if (k == m) {
          s = fsign(sqrt(p * p + q * q + r * r), p);
          if (lb != m) AB(k, k - 1) = -AB(k, k - 1);
        } else {
          p = AB(k, k - 1);
          q = AB(k + 1, k - 1);
          r = 0.0;
          if (notlast) r = AB(k + 2, k - 1);
          x = fabs(p) + fabs(q) + fabs(r);
          if (x == 0.0) continue;
          p = p / x;
          q = q / x;
          r = r / x;
          s = fsign(sqrt(p * p + q * q + r * r), p);
          AB(k, k - 1) = -s * x;
        }
        p = p + s;
        s = 1.0 / s;
        x = p * s;
        y = q * s;
        z = r * s;
        p = 1.0 / p;
        q = q * p;
        r = r * p;
        for (int j = k; j <= n; ++j) {  // row modification
          p = AB(k, j) + q * AB(k + 1, j);
          if (notlast) {
            p = p + r * AB(k + 2, j);
            AB(k + 2, j) = AB(k + 2, j) - p * z;
          }
          AB(k + 1, j) = AB(k + 1, j) - p * y;
          AB(k, j) = AB(k, j) - p * x;
        }
        const int iend = imin(en, k + 3);
        for (int i = 1; i <= iend; ++i) {  // column modification
          p = x * AB(i, k) + y * AB(i, k + 1);
          if (notlast) {
            p = p + z * AB(i, k + 2);
            AB(i, k + 2) = AB(i, k + 2) - p * r;
          }
          AB(i, k + 1) = AB(i, k + 1) - p * q;
          AB(i, k) = AB(i, k) - p;
        }
        for (int i = lo; i <= hi; ++i) {  // accumulate
          p = x * EV(i, k) + y * EV(i, k + 1);
          if (notlast) {
            p = p + z * EV(i, k + 2);
            EV(i, k + 2) = EV(i, k + 2) - p * r;
          }
          EV(i, k + 1) = EV(i, k + 1) - p * q;
          EV(i, k) = EV(i, k) - p;
        }
      }
    }
  }

  // -- back substitution for the vectors of the triangular form ---------------
  if (!failed) {
    if (rnorm != 0.0) {
      for (int e = n; e >= 1; --e) {
        int nn = e;
        AB(e, e) = 1.0;
        for (int i = e - 1; i >= 1; --i) {
          w = AB(i, i) - EL(e);
          if (fabs(w) < fabs(Tol * rnorm)) w = fsign(Tol * rnorm, w);
          r = AB(i, e);
          for (int j = nn; j <= e - 1; ++j) r = r + AB(i, j) * AB(j, e);
          AB(i, e) = -r / w;
          nn = i;
        }
      }
      for (int i = 1; i <= n; ++i)
        if (i < lo || i > hi)
          for (int j = i; j <= n; ++j) EV(i, j) = AB(i, j);
      for (int j = n; j >= lo; --j)
        for (int i = lo; i <= hi; ++i) {
          z = 0.0;
          const int jend = imin(j, hi);
          for (int k = lo; k <= jend; ++k) z = z + EV(i, k) * AB(k, j);
          EV(i, j) = z;
        }
    }
    for (int i = lo; i <= hi; ++i)
      for (int j = 1; j <= n; ++j) EV(i, j) = EV(i, j) * WK(i);
    for (int i = lo - 1; i >= 1; --i) {
      const int j = (int)(WK(i) + 0.5);
      if (i < j)
        for (int k = 1; k <= n; ++k) {
          const double tt = EV(i, k);
          EV(i, k) = EV(j, k);
          EV(j, k) = tt;
        }
    }
    for (int i = hi + 1; i <= n; ++i) {
      const int j = (int)(WK(i) + 0.5);
      if (i != j)
        for (int k = 1; k <= n; ++k) {
          const double tt = EV(i, k);
          EV(i, k) = EV(j, k);
          EV(j, k) = tt;
        }
    }
  }
#undef AB
#undef EV
#undef WK
#undef EL
  return nerr;
}

}  // namespace ssb
