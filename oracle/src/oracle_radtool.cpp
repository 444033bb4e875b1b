// oracle_radtool.cpp - CPU restatement of the reference's radtool/ layer.
// TEST INFRASTRUCTURE ONLY (see oracle_radtool.hpp header).
#include "oracle_radtool.hpp"

namespace orc {

thread_local double g_flops = 0.0;
#if defined(ORACLE_QUAD)
static const int kMaxQrIter = 90;
#else
static const int kMaxQrIter = 30;
#endif

static const real kPi = 3.14159265358979323846; // radiation_constants.F90:24

// calc_legendre_gauss: radtool/radtool_legendre_gauss.F90:119-170.
void calc_legendre_gauss(int nnode, real x1, real x2, Vec &xnode, Vec &weight) {
  const int n = nnode;
  Vec ynode(n), ynode0(n), lgvm_deriv(n);
  std::vector<Vec> lgvm(n + 1, Vec(n)); // lgvm[k][node], k = 0..n
  for (int jn = 1; jn <= n; ++jn) {
    // "(0.27/nnode)" is a default-real (single precision) expression (:142)
    const float c027 = 0.27f / (float)n;
    ynode[jn - 1] = std::cos((2 * (jn - 1) + 1) * kPi / (2 * n)) +
                    (real)c027 * std::sin(kPi * (-1.0 + 2.0 * jn) / (n + 1));
    ynode0[jn - 1] = 2.0;
  }
  const real eps = std::numeric_limits<real>::epsilon();
  for (;;) {
    real maxdiff = 0.0;
    for (int i = 0; i < n; ++i) maxdiff = rmax(maxdiff, std::fabs(ynode[i] - ynode0[i]));
    if (!(maxdiff > eps)) break;
    for (int i = 0; i < n; ++i) {
      lgvm[0][i] = 1.0;
      if (n >= 1) lgvm[1][i] = ynode[i];
    }
    for (int jn = 2; jn <= n; ++jn)
      for (int i = 0; i < n; ++i)
        lgvm[jn][i] =
            ((2 * jn - 1) * ynode[i] * lgvm[jn - 1][i] - (jn - 1) * lgvm[jn - 2][i]) / jn;
    for (int i = 0; i < n; ++i)
      lgvm_deriv[i] =
          (n + 1) * (lgvm[n - 1][i] - ynode[i] * lgvm[n][i]) / (1.0 - ynode[i] * ynode[i]);
    ynode0 = ynode;
    for (int i = 0; i < n; ++i) ynode[i] = ynode0[i] - lgvm[n][i] / lgvm_deriv[i];
  }
  xnode.assign(n, 0.0);
  weight.assign(n, 0.0);
  for (int i = 0; i < n; ++i) {
    // sic: both terms use (1-y) (:165)
    xnode[i] = 0.5 * (x1 * (1.0 - ynode[i]) + x2 * (1.0 - ynode[i]));
    weight[i] = (((n + 1) * (n + 1)) / (real)(n * n)) * (x2 - x1) /
                ((1.0 - ynode[i] * ynode[i]) * lgvm_deriv[i] * lgvm_deriv[i]);
  }
}

// initialize_legendre_gauss: radtool/radtool_legendre_gauss.F90:52-100.
void legendre_gauss_initialize(LegendreGauss &lg, int nstream) {
  lg.nstream = nstream;
  calc_legendre_gauss(nstream, 0.0, 1.0, lg.mu, lg.weight);
  lg.sin_ang.resize(nstream);
  lg.tan_ang.resize(nstream);
  lg.hweight.resize(nstream);
  lg.vweight.resize(nstream);
  real sumh = 0.0, sumv = 0.0;
  for (int i = 0; i < nstream; ++i) {
    lg.sin_ang[i] = std::sqrt(1.0 - lg.mu[i] * lg.mu[i]);
    lg.tan_ang[i] = lg.sin_ang[i] / lg.mu[i];
    lg.hweight[i] = lg.weight[i] * lg.mu[i];
    lg.vweight[i] = lg.weight[i] * lg.sin_ang[i];
  }
  for (int i = 0; i < nstream; ++i) {
    sumh += lg.hweight[i];
    sumv += lg.vweight[i];
  }
  for (int i = 0; i < nstream; ++i) {
    lg.hweight[i] = lg.hweight[i] / sumh;
    lg.vweight[i] = lg.vweight[i] / sumv;
  }
  lg.vadjustment = 1.0;
  real s = 0.0;
  for (int i = 0; i < nstream; ++i) s += lg.weight[i] * lg.sin_ang[i];
  lg.vadjustment2 = (kPi / 4.0) / s;
}

static inline real fsign(real a, real b) { // Fortran SIGN(a,b)
  return std::signbit(b) ? -std::fabs(a) : std::fabs(a);
}

// eigen_decomposition_real: radtool/radtool_eigen_decomposition.F90:51-828
// (ASYMTX of DISORT: balance, Hessenberg, shifted real-QR, back
// substitution).  1-based indexing is kept internally so that loop bounds and
// the post-loop values of loop variables (App. B9) read like the Fortran.
int eigen_decomposition_real(int norder, const Mat &amat, Vec &eigenvalue, Mat &eigenvector) {
  count_flops(norder >= 3 ? 25.0 * norder * norder * norder : (norder == 2 ? 20.0 : 1.0));
  const real Tol = std::numeric_limits<real>::epsilon();
  const real C1 = 0.4375, C2 = 0.5, C3 = 0.75, C4 = 0.95, C5 = 16.0, C6 = 256.0;
  const int n = norder;
  eigenvalue.assign(n, 0.0);
  eigenvector = Mat(n, n);
  int nerror = 0;

  if (n > 2) {
    std::vector<real> abal_((size_t)(n + 1) * (n + 1), 0.0), evec_((size_t)(n + 1) * (n + 1), 0.0);
    std::vector<real> eval_(n + 1, 0.0), wkd(2 * n + 1, 0.0);
#define ABAL(i, j) abal_[(size_t)(i) + (size_t)(n + 1) * (j)]
#define EVEC(i, j) evec_[(size_t)(i) + (size_t)(n + 1) * (j)]
    bool is_error = false;
    for (int ji = 1; ji <= n; ++ji) EVEC(ji, ji) = 1.0;
    for (int i = 1; i <= n; ++i)
      for (int j = 1; j <= n; ++j) ABAL(i, j) = amat(i - 1, j - 1);

    real rnorm = 0.0;
    int ll = 1, kk = n;
    int ji = 0, jj = 0, jn = 0;
    real tmp;

    // Search for rows isolating an eigenvalue and push them down (:184-222)
    bool not_finished = true;
    while (not_finished) {
      not_finished = false;
      const int kkk = kk;
      for (jj = kkk; jj >= 1; --jj) {
        real row = 0.0;
        for (ji = 1; ji <= kk; ++ji)
          if (ji != jj) row = row + std::fabs(ABAL(jj, ji));
        // here ji == kk+1 (loop-exit value), so "ji /= kk" below is always true
        if (row == 0.0) {
          wkd[kk] = jj;
          if (ji != kk) {
            for (ji = 1; ji <= kk; ++ji) {
              tmp = ABAL(ji, jj);
              ABAL(ji, jj) = ABAL(ji, kk);
              ABAL(ji, kk) = tmp;
            }
            for (ji = ll; ji <= n; ++ji) {
              tmp = ABAL(jj, ji);
              ABAL(jj, ji) = ABAL(kk, ji);
              ABAL(kk, ji) = tmp;
            }
          }
          kk = kk - 1;
          not_finished = true;
          break;
        } else {
          not_finished = false;
        }
      }
    }

    // Search for columns isolating an eigenvalue and push them left (:227-262)
    not_finished = true;
    while (not_finished) {
      not_finished = false;
      const int lll = ll;
      for (jj = lll; jj <= kk; ++jj) {
        real column = 0.0;
        for (ji = ll; ji <= kk; ++ji)
          if (ji != jj) column = column + std::fabs(ABAL(ji, jj));
        if (column == 0.0) {
          wkd[ll] = jj;
          if (jj != ll) {
            for (ji = 1; ji <= kk; ++ji) {
              tmp = ABAL(ji, jj);
              ABAL(ji, jj) = ABAL(ji, ll);
              ABAL(ji, ll) = tmp;
            }
            for (ji = ll; ji <= n; ++ji) {
              tmp = ABAL(jj, ji);
              ABAL(jj, ji) = ABAL(ll, ji);
              ABAL(ll, ji) = tmp;
            }
          }
          ll = ll + 1;
          not_finished = true;
          break;
        } else {
          not_finished = false;
        }
      }
    }

    // Balance the submatrix in rows ll through kk (:265-310)
    for (ji = ll; ji <= kk; ++ji) wkd[ji] = 1.0;
    not_finished = true;
    while (not_finished) {
      not_finished = false;
      for (ji = ll; ji <= kk; ++ji) {
        real column = 0.0, row = 0.0;
        for (jj = ll; jj <= kk; ++jj)
          if (jj != ji) {
            column = column + std::fabs(ABAL(jj, ji));
            row = row + std::fabs(ABAL(ji, jj));
          }
        real ff = 1.0;
        real gg = row / C5;
        const real hh = column + row;
        while (column < gg) {
          ff = ff * C5;
          column = column * C6;
        }
        gg = row * C5;
        while (column > gg) {
          ff = ff / C5;
          column = column / C6;
        }
        if ((column + row) / ff < C4 * hh) {
          wkd[ji] = wkd[ji] * ff;
          not_finished = true;
          for (jj = ll; jj <= n; ++jj) ABAL(ji, jj) = ABAL(ji, jj) / ff;
          for (jj = 1; jj <= kk; ++jj) ABAL(jj, ji) = ABAL(jj, ji) * ff;
        }
      }
    }

    // Reduce to Hessenberg form and accumulate (:313-395)
    if (kk - 1 >= ll + 1) {
      for (jn = ll + 1; jn <= kk - 1; ++jn) {
        real hh = 0.0;
        wkd[jn + n] = 0.0;
        real scale = 0.0;
        for (ji = jn; ji <= kk; ++ji) scale = scale + std::fabs(ABAL(ji, jn - 1));
        if (scale != 0.0) {
          for (ji = kk; ji >= jn; --ji) {
            wkd[ji + n] = ABAL(ji, jn - 1) / scale;
            hh = hh + wkd[ji + n] * wkd[ji + n];
          }
          real gg = -fsign(std::sqrt(hh), wkd[jn + n]);
          hh = hh - wkd[jn + n] * gg;
          wkd[jn + n] = wkd[jn + n] - gg;
          hh = 1.0 / hh;
          for (jj = jn; jj <= n; ++jj) {
            real ff = 0.0;
            for (ji = kk; ji >= jn; --ji) ff = ff + wkd[ji + n] * ABAL(ji, jj);
            for (ji = jn; ji <= kk; ++ji) ABAL(ji, jj) = ABAL(ji, jj) - wkd[ji + n] * ff * hh;
          }
          for (ji = 1; ji <= kk; ++ji) {
            real ff = 0.0;
            for (jj = kk; jj >= jn; --jj) ff = ff + wkd[jj + n] * ABAL(ji, jj);
            for (jj = jn; jj <= kk; ++jj) ABAL(ji, jj) = ABAL(ji, jj) - wkd[jj + n] * ff * hh;
          }
          wkd[jn + n] = scale * wkd[jn + n];
          ABAL(jn, jn - 1) = scale * gg;
        }
      }
      for (jn = kk - 2; jn >= ll; --jn) {
        const int n1 = jn + 1, n2 = jn + 2;
        real ff = ABAL(n1, jn);
        if (ff != 0.0) {
          ff = ff * wkd[jn + 1 + n];
          for (ji = n2; ji <= kk; ++ji) wkd[ji + n] = ABAL(ji, jn);
          if (n1 < kk) {
            for (jj = 1; jj <= n; ++jj) {
              real gg = 0.0;
              for (ji = n1; ji <= kk; ++ji) gg = gg + wkd[ji + n] * EVEC(ji, jj);
              gg = gg / ff;
              for (ji = n1; ji <= kk; ++ji) EVEC(ji, jj) = EVEC(ji, jj) + gg * wkd[ji + n];
            }
          }
        }
      }
    }

    // Norm and isolated eigenvalues (:397-408)
    jn = 1;
    for (ji = 1; ji <= n; ++ji) {
      for (jj = jn; jj <= n; ++jj) rnorm = rnorm + std::fabs(ABAL(ji, jj));
      jn = ji;
      if (ji < ll || ji > kk) eval_[ji] = ABAL(ji, ji);
    }
    jn = kk;
    real tt = 0.0;
    real pp = 0, qq = 0, rr = 0, ss = 0, xx = 0, yy = 0, zz = 0, ww = 0, uu = 0, vv = 0;

    // Search for next eigenvalue (:411-635)
    not_finished = true;
    while (not_finished && jn >= ll) {
      not_finished = false;
      int in = 0;
      const int n1 = jn - 1, n2 = jn - 2;
      int lb = ll;
      for (;;) { // do while (not_found)
        for (ji = ll; ji <= jn; ++ji) {
          lb = jn + ll - ji;
          if (lb == ll) break;
          ss = std::fabs(ABAL(lb - 1, lb - 1)) + std::fabs(ABAL(lb, lb));
          if (ss == 0.0) ss = rnorm;
          if (std::fabs(ABAL(lb, lb - 1)) < Tol * ss) break;
        }
        xx = ABAL(jn, jn);
        if (lb == jn) {
          // One eigenvalue found
          ABAL(jn, jn) = xx + tt;
          eval_[jn] = ABAL(jn, jn);
          jn = n1;
          not_finished = true;
          break;
        }
        yy = ABAL(n1, n1);
        ww = ABAL(jn, n1) * ABAL(n1, jn);
        if (lb == n1) {
          // Two eigenvalues found
          pp = (yy - xx) * C2;
          qq = pp * pp + ww;
          zz = std::sqrt(std::fabs(qq));
          ABAL(jn, jn) = xx + tt;
          xx = ABAL(jn, jn);
          ABAL(n1, n1) = yy + tt;
          zz = pp + fsign(zz, pp);
          eval_[n1] = xx + zz;
          eval_[jn] = eval_[n1];
          if (zz != 0.0) eval_[jn] = xx - ww / zz;
          xx = ABAL(jn, n1);
          rr = 1.0 / std::sqrt(xx * xx + zz * zz);
          pp = xx * rr;
          qq = zz * rr;
          for (jj = n1; jj <= n; ++jj) {
            zz = ABAL(n1, jj);
            ABAL(n1, jj) = qq * zz + pp * ABAL(jn, jj);
            ABAL(jn, jj) = qq * ABAL(jn, jj) - pp * zz;
          }
          for (ji = 1; ji <= jn; ++ji) {
            zz = ABAL(ji, n1);
            ABAL(ji, n1) = qq * zz + pp * ABAL(ji, jn);
            ABAL(ji, jn) = qq * ABAL(ji, jn) - pp * zz;
          }
          for (ji = ll; ji <= kk; ++ji) {
            zz = EVEC(ji, n1);
            EVEC(ji, n1) = qq * zz + pp * EVEC(ji, jn);
            EVEC(ji, jn) = qq * EVEC(ji, jn) - pp * zz;
          }
          jn = n2;
          not_finished = true;
          break;
        }
        if (in == kMaxQrIter) {
          // no convergence after 30 iterations (:498-510); the quad build gets more
          // iterations: its tolerance is 1e-34, not 2e-16
          nerror = nerror + 1;
          is_error = true;
          not_finished = false;
          break;
        }
        // Form shift (:513-523)
        if (in == 10 || in == 20 || in == 40 || in == 60) {
          tt = tt + xx;
          for (ji = ll; ji <= jn; ++ji) ABAL(ji, ji) = ABAL(ji, ji) - xx;
          ss = std::fabs(ABAL(jn, n1)) + std::fabs(ABAL(n1, n2));
          xx = C3 * ss;
          yy = xx;
          ww = -C1 * ss * ss;
        }
        in = in + 1;
        // Look for two consecutive small sub-diagonal elements (:528-549)
        for (jj = lb; jj <= n2; ++jj) {
          ji = n2 + lb - jj;
          zz = ABAL(ji, ji);
          rr = xx - zz;
          ss = yy - zz;
          pp = (rr * ss - ww) / ABAL(ji + 1, ji) + ABAL(ji, ji + 1);
          qq = ABAL(ji + 1, ji + 1) - zz - rr - ss;
          rr = ABAL(ji + 2, ji + 1);
          ss = 1.0 / (std::fabs(pp) + std::fabs(qq) + std::fabs(rr));
          pp = pp * ss;
          qq = qq * ss;
          rr = rr * ss;
          if (ji == lb) break;
          uu = std::fabs(ABAL(ji, ji - 1)) * (std::fabs(qq) + std::fabs(rr));
          vv = std::fabs(pp) * (std::fabs(ABAL(ji - 1, ji - 1)) + std::fabs(zz) +
                                std::fabs(ABAL(ji + 1, ji + 1)));
          if (uu <= Tol * vv) break;
        }
        ABAL(ji + 2, ji) = 0.0;
        for (jj = ji + 3; jj <= jn; ++jj) {
          ABAL(jj, jj - 2) = 0.0;
          ABAL(jj, jj - 3) = 0.0;
        }
        // Double QR step involving rows K to N and columns M to N (:559-633).
        // The Fortran reuses ji as an inner loop variable; DO bounds are
        // evaluated once, and "ka == ji" can only hold on the first trip.
        const int ka_first = ji;
        for (int ka = ka_first; ka <= n1; ++ka) {
          const bool not_last = (ka != n1);
          if (ka == ka_first) {
            ss = fsign(std::sqrt(pp * pp + qq * qq + rr * rr), pp);
            if (lb != ka_first) ABAL(ka, ka - 1) = -ABAL(ka, ka - 1);
          } else {
            pp = ABAL(ka, ka - 1);
            qq = ABAL(ka + 1, ka - 1);
            rr = 0.0;
            if (not_last) rr = ABAL(ka + 2, ka - 1);
            xx = std::fabs(pp) + std::fabs(qq) + std::fabs(rr);
            if (xx == 0.0) continue;
            pp = pp / xx;
            qq = qq / xx;
            rr = rr / xx;
            ss = fsign(std::sqrt(pp * pp + qq * qq + rr * rr), pp);
            ABAL(ka, ka - 1) = -ss * xx;
          }
          pp = pp + ss;
          ss = 1.0 / ss;
          xx = pp * ss;
          yy = qq * ss;
          zz = rr * ss;
          pp = 1.0 / pp;
          qq = qq * pp;
          rr = rr * pp;
          // Row modification
          for (jj = ka; jj <= n; ++jj) {
            pp = ABAL(ka, jj) + qq * ABAL(ka + 1, jj);
            if (not_last) {
              pp = pp + rr * ABAL(ka + 2, jj);
              ABAL(ka + 2, jj) = ABAL(ka + 2, jj) - pp * zz;
            }
            ABAL(ka + 1, jj) = ABAL(ka + 1, jj) - pp * yy;
            ABAL(ka, jj) = ABAL(ka, jj) - pp * xx;
          }
          // Column modification
          const int iend = std::min(jn, ka + 3);
          for (ji = 1; ji <= iend; ++ji) {
            pp = xx * ABAL(ji, ka) + yy * ABAL(ji, ka + 1);
            if (not_last) {
              pp = pp + zz * ABAL(ji, ka + 2);
              ABAL(ji, ka + 2) = ABAL(ji, ka + 2) - pp * rr;
            }
            ABAL(ji, ka + 1) = ABAL(ji, ka + 1) - pp * qq;
            ABAL(ji, ka) = ABAL(ji, ka) - pp;
          }
          // Accumulate transformations
          for (ji = ll; ji <= kk; ++ji) {
            pp = xx * EVEC(ji, ka) + yy * EVEC(ji, ka + 1);
            if (not_last) {
              pp = pp + zz * EVEC(ji, ka + 2);
              EVEC(ji, ka + 2) = EVEC(ji, ka + 2) - pp * rr;
            }
            EVEC(ji, ka + 1) = EVEC(ji, ka + 1) - pp * qq;
            EVEC(ji, ka) = EVEC(ji, ka) - pp;
          }
        }
      }
    }

    // Back-substitution (:640-717)
    if (!is_error) {
      if (rnorm != 0.0) {
        for (jn = n; jn >= 1; --jn) {
          int n2 = jn;
          ABAL(jn, jn) = 1.0;
          for (ji = jn - 1; ji >= 1; --ji) {
            ww = ABAL(ji, ji) - eval_[jn];
            if (std::fabs(ww) < std::fabs(Tol * rnorm)) ww = fsign(Tol * rnorm, ww);
            rr = ABAL(ji, jn);
            for (jj = n2; jj <= jn - 1; ++jj) rr = rr + ABAL(ji, jj) * ABAL(jj, jn);
            ABAL(ji, jn) = -rr / ww;
            n2 = ji;
          }
        }
        for (ji = 1; ji <= n; ++ji)
          if (ji < ll || ji > kk)
            for (jj = ji; jj <= n; ++jj) EVEC(ji, jj) = ABAL(ji, jj);
        if ((real)kk != 0.0) { // sic (:672)
          for (jj = n; jj >= ll; --jj)
            for (ji = ll; ji <= kk; ++ji) {
              zz = 0.0;
              const int jend = std::min(jj, kk);
              for (jn = ll; jn <= jend; ++jn) zz = zz + EVEC(ji, jn) * ABAL(jn, jj);
              EVEC(ji, jj) = zz;
            }
        }
      }
      for (ji = ll; ji <= kk; ++ji)
        for (jj = 1; jj <= n; ++jj) EVEC(ji, jj) = EVEC(ji, jj) * wkd[ji];
      for (ji = ll - 1; ji >= 1; --ji) {
        jj = (int)std::lround(wkd[ji]);
        if (ji < jj)
          for (jn = 1; jn <= n; ++jn) {
            tmp = EVEC(ji, jn);
            EVEC(ji, jn) = EVEC(jj, jn);
            EVEC(jj, jn) = tmp;
          }
      }
      for (ji = kk + 1; ji <= n; ++ji) {
        jj = (int)std::lround(wkd[ji]);
        if (ji != jj)
          for (jn = 1; jn <= n; ++jn) {
            tmp = EVEC(ji, jn);
            EVEC(ji, jn) = EVEC(jj, jn);
            EVEC(jj, jn) = tmp;
          }
      }
    }
    for (int i = 1; i <= n; ++i) {
      eigenvalue[i - 1] = eval_[i];
      for (int j = 1; j <= n; ++j) eigenvector(i - 1, j - 1) = EVEC(i, j);
    }
#undef ABAL
#undef EVEC
  } else if (n == 2) {
    // :770-803
    const real a11 = amat(0, 0), a12 = amat(0, 1), a21 = amat(1, 0), a22 = amat(1, 1);
    const real discriminant = (a11 - a22) * (a11 - a22) + 4.0 * a12 * a21;
    if (discriminant < 0.0) nerror = nerror + 1;
    eigenvalue[0] = 0.5 * (a11 + a22);
    eigenvalue[1] = eigenvalue[0];
    const real half_sqrt_disc = 0.5 * std::sqrt(discriminant);
    if (a11 >= a22) {
      eigenvalue[0] = eigenvalue[0] + half_sqrt_disc;
      eigenvalue[1] = eigenvalue[1] - half_sqrt_disc;
    } else {
      eigenvalue[0] = eigenvalue[0] - half_sqrt_disc;
      eigenvalue[1] = eigenvalue[1] + half_sqrt_disc;
    }
    eigenvector(0, 0) = 1.0;
    eigenvector(1, 1) = 1.0;
    if (a11 == a22 && (a21 == 0.0 || a12 == 0.0)) {
      // sic: Tol multiplies only the first term (:795-797)
      const real rnorm =
          1.0 / (Tol * std::fabs(a11) + std::fabs(a21) + std::fabs(a12) + std::fabs(a22));
      eigenvector(1, 0) = a21 * rnorm;
      eigenvector(0, 1) = a12 * rnorm;
    } else {
      eigenvector(1, 0) = a21 / (eigenvalue[0] - a22);
      eigenvector(0, 1) = a12 / (eigenvalue[1] - a11);
    }
  } else if (n == 1) {
    eigenvalue[0] = amat(0, 0);
    eigenvector(0, 0) = 1.0;
  } else {
    nerror = nerror + 1;
  }
  return nerror;
}

// schur_invert_sw: radtool/radtool_schur.F90:32-53.
void schur_invert_sw(const Mat &g0, const Mat &g1, const Mat &g2, const Mat &g3, Mat &g0i,
                     Mat &g1i, Mat &g2i, Mat &g3i) {
  g0i = invert(g0);
  g1i = invert(g1 - matmul(g2, solve_mat(g1, g2)));
  g2i = matmul(g1i, matmul(g2, invert(g1)));
  g3i = matmul(g1i - g2i, matmul(g3, g0i));
}

// direct_diffuse_part: radtool/radtool_calc_matrices_sw_eig.F90:303-386.
static void direct_diffuse_part(int ndiff, int ndir, const Vec &exp_lambda_dz,
                                const Vec &exp_eigenval_dir_dz, const Mat &g0, const Mat &g1,
                                const Mat &g2, const Mat &g3, const Mat &g4, Mat &s_up,
                                Mat &s_dn) {
  const int N = 2 * ndiff + ndir;
  Mat g_d(N, N), g_d2(ndiff, N), rhs(N, ndir);
  paste(g_d, 0, 0, g1);
  Mat g2d = scale_cols(g2, exp_lambda_dz);
  paste(g_d, ndiff, 0, g2d);
  paste(g_d, 0, ndiff, g2d);
  paste(g_d, ndiff, ndiff, g1);
  paste(g_d, 2 * ndiff, 2 * ndiff, g0);
  paste(g_d, 0, 2 * ndiff, scale_cols(g3, exp_eigenval_dir_dz));
  paste(g_d, ndiff, 2 * ndiff, g4);
  for (int jj = 0; jj < ndir; ++jj) rhs(2 * ndiff + jj, jj) = 1.0;
  Mat cprime_dir = solve_rect_mat(g_d, rhs);
  paste(g_d2, 0, 0, scale_cols(g1, exp_lambda_dz));
  paste(g_d2, 0, ndiff, g2);
  paste(g_d2, 0, 2 * ndiff, g3);
  s_up = matmul(g_d2, cprime_dir);
  paste(g_d2, 0, 0, g2);
  paste(g_d2, 0, ndiff, scale_cols(g1, exp_lambda_dz));
  paste(g_d2, 0, 2 * ndiff, scale_cols(g4, exp_eigenval_dir_dz));
  s_dn = matmul(g_d2, cprime_dir);
}

// calc_matrices_sw_eig: radtool/radtool_calc_matrices_sw_eig.F90:30-298.
void calc_matrices_sw_eig(int ndiff, int ndir, real dz, real /*mu0 unused*/,
                          const Mat &gamma0, const Mat &gamma1, const Mat &gamma2,
                          const Mat &gamma3, Mat &reflectance, Mat &transmittance, Mat &s_up,
                          Mat &s_dn, Mat &trans_dir, Mat &int_dir, Mat &int_diff,
                          Mat &int_dir_diff) {
  // Section 1 (:180-196)
  Mat gamma_diff = gamma1 - gamma2;
  Mat gamma_product = matmul(gamma_diff, gamma1 + gamma2);
  Vec eigenval_prod;
  Mat eigenvec_prod;
  eigen_decomposition_real(ndiff, gamma_product, eigenval_prod, eigenvec_prod);
  Vec lambda(ndiff), exp_lambda_dz(ndiff);
  for (int i = 0; i < ndiff; ++i) {
    lambda[i] = std::sqrt(rmax(0.0, eigenval_prod[i]));
    exp_lambda_dz[i] = std::exp(-lambda[i] * dz);
  }
  Mat tmp_mat = neg(solve_mat(gamma_diff, eigenvec_prod));
  tmp_mat = scale_cols(tmp_mat, lambda);
  Mat g1 = eigenvec_prod + tmp_mat;
  Mat g2 = eigenvec_prod - tmp_mat;

  // Section 2 (:205-221)
  Mat g1_d = scale_cols(g1, exp_lambda_dz);
  Mat g2_d = scale_cols(g2, exp_lambda_dz);
  Mat cprime_lower = invert(g1 - matmul(g2_d, solve_mat(g1, g2_d)));
  Mat cprime_upper = neg(solve_mat(g1, matmul(g2_d, cprime_lower)));
  reflectance = matmul(g1_d, cprime_upper) + matmul(g2, cprime_lower);
  transmittance = matmul(g2, cprime_upper) + matmul(g1_d, cprime_lower);

  // Section 3 (:225-229)
  Vec eigenval_dir;
  Mat g0;
  eigen_decomposition_real(ndir, gamma0, eigenval_dir, g0);
  Mat g0_inv = invert(g0);
  Vec exp_eig_dir(ndir);
  for (int i = 0; i < ndir; ++i) exp_eig_dir[i] = std::exp(eigenval_dir[i] * dz);
  trans_dir = matmul(scale_cols(g0, exp_eig_dir), g0_inv);

  // Section 4 (:232-253)
  Mat gamma3_g0 = matmul(gamma3, g0);
  Mat gamma1_d = gamma1;
  Mat g3(ndiff, ndir), g4(ndiff, ndir);
  for (int jd = 0; jd < ndir; ++jd) {
    for (int jo = 0; jo < ndiff; ++jo) gamma1_d(jo, jo) = gamma1(jo, jo) + eigenval_dir[jd];
    Mat gamma2_inv_gamma1_d = matmul(gamma2, invert(gamma1_d));
    tmp_mat = gamma1 - matmul(gamma2_inv_gamma1_d, gamma2);
    for (int jo = 0; jo < ndiff; ++jo) tmp_mat(jo, jo) = tmp_mat(jo, jo) - eigenval_dir[jd];
    for (int jo = 0; jo < ndiff; ++jo)
      gamma2_inv_gamma1_d(jo, jo) = gamma2_inv_gamma1_d(jo, jo) - 1.0;
    Vec col(ndiff);
    for (int i = 0; i < ndiff; ++i) col[i] = gamma3_g0(i, jd);
    Vec g4c = solve_vec(tmp_mat, matvec(gamma2_inv_gamma1_d, col));
    Vec g3c = solve_vec(gamma1_d, col + matvec(gamma2, g4c));
    for (int i = 0; i < ndiff; ++i) {
      g4(i, jd) = g4c[i];
      g3(i, jd) = -g3c[i];
    }
  }
  direct_diffuse_part(ndiff, ndir, exp_lambda_dz, exp_eig_dir, g0, g1, g2, g3, g4, s_up, s_dn);

  // Integrated-flux matrices (:289-296)
  Mat gamma0i, gamma1i, gamma2i, gamma3i;
  schur_invert_sw(gamma0, gamma1, gamma2, gamma3, gamma0i, gamma1i, gamma2i, gamma3i);
  int_dir = neg(gamma0i);
  int_diff = gamma2i - gamma1i;
  int_dir_diff = Mat(ndiff, ndir);
  for (size_t k = 0; k < gamma3i.a.size(); ++k) int_dir_diff.a[k] = 2.0 * gamma3i.a[k];
}

// calc_matrices_lw_eig: radtool/radtool_calc_matrices_lw_eig.F90:32-230.
void calc_matrices_lw_eig(int norder, real dz, const Mat &gamma1, const Mat &gamma2,
                          const Vec &emiss_rate, Mat &reflectance, Mat &transmittance,
                          Vec &source, Mat &int_flux, Vec &int_flux_source) {
  const int n = norder;
  Mat gamma_diff = gamma1 - gamma2;
  Mat gamma_product = matmul(gamma_diff, gamma1 + gamma2);
  Vec eigenval_prod;
  Mat eigenvec_prod;
  eigen_decomposition_real(n, gamma_product, eigenval_prod, eigenvec_prod);
  Vec lambda(n), exp_lambda_dz(n);
  for (int i = 0; i < n; ++i) {
    lambda[i] = std::sqrt(rmax(0.0, eigenval_prod[i]));
    exp_lambda_dz[i] = std::exp(-lambda[i] * dz);
  }
  Mat tmp_mat = neg(solve_mat(gamma_diff, eigenvec_prod));
  tmp_mat = scale_cols(tmp_mat, lambda);
  Mat g1 = eigenvec_prod + tmp_mat;
  Mat g2 = eigenvec_prod - tmp_mat;
  Mat g1_d = scale_cols(g1, exp_lambda_dz);
  Mat g2_d = scale_cols(g2, exp_lambda_dz);
  Mat cprime_lower = invert(g1 - matmul(g2_d, solve_mat(g1, g2_d)));
  Mat cprime_upper = neg(solve_mat(g1, matmul(g2_d, cprime_lower)));
  reflectance = matmul(g1_d, cprime_upper) + matmul(g2, cprime_lower);
  transmittance = matmul(g2, cprime_upper) + matmul(g1_d, cprime_lower);

  // Source terms (:188-211)
  Mat gamma2_inv_gamma1 = matmul(gamma2, invert(gamma1));
  tmp_mat = gamma1 - matmul(gamma2_inv_gamma1, gamma2);
  for (int jo = 0; jo < n; ++jo) gamma2_inv_gamma1(jo, jo) = gamma2_inv_gamma1(jo, jo) - 1.0;
  Vec inv_gamma_b = solve_vec(tmp_mat, matvec(gamma2_inv_gamma1, emiss_rate));
  Vec inv_g1_inv_gamma_b = solve_vec(g1, inv_gamma_b);
  Vec tmp_vec = inv_gamma_b - matvec(g2_d, inv_g1_inv_gamma_b);
  Vec cb_prime = matvec(cprime_lower, tmp_vec);
  for (int i = 0; i < n; ++i) cb_prime[i] = -cb_prime[i];
  source = matvec(g1_d + g2, cb_prime) + inv_gamma_b;

  // Integrated fluxes (:213-227)
  for (int i = 0; i < n; ++i) exp_lambda_dz[i] = (1.0 - exp_lambda_dz[i]) / lambda[i];
  g1 = scale_cols(g1, exp_lambda_dz);
  g2 = scale_cols(g2, exp_lambda_dz);
  tmp_mat = g1 + g2;
  int_flux = matmul(tmp_mat, cprime_lower + cprime_upper);
  Vec t2 = matvec(tmp_mat, cb_prime);
  int_flux_source.assign(n, 0.0);
  for (int i = 0; i < n; ++i) int_flux_source[i] = 2.0 * (t2[i] + inv_gamma_b[i] * dz);
}

} // namespace orc
