// ssb_radtool.cuh - layer transfer matrices for one (column, interval, layer)
// problem per thread.
//
// Replaces calc_matrices_sw_eig + direct_diffuse_part
// (radtool/radtool_calc_matrices_sw_eig.F90:30-298, :303-386),
// calc_matrices_lw_eig (radtool/radtool_calc_matrices_lw_eig.F90:32-230) and
// schur_invert_sw (radtool/radtool_schur.F90:32-53).
//
// NC is the compile-time capacity of the diffuse order n = nreg*ns; the
// direct order d = nreg <= 3.  All matrices are compact column-major with the
// RUNTIME orders n, d (sub-block branches pass smaller orders).  Work buffers
// are reused aggressively to bound the per-thread stack.
#pragma once
#include "ssb_math.cuh"

namespace ssb {

template <int NC>
struct RadtoolWork {
  static constexpr int NN = NC * NC;
  static constexpr int NBIG = 2 * NC + 3;
  double b0[NN], b1[NN], b2[NN], b3[NN], b4[NN], b5[NN], b6[NN];
  double big[NBIG * NBIG];
  double cp[NBIG * 3];
  double lam[NC], elz[NC], wk[2 * NC + 1], col[NC], rhs[NC], g4c[NC];
  double g3[NC * 3], g4[NC * 3], g3g0[NC * 3];
};

// Steps common to SW and LW (sw_eig:180-221, lw_eig:142-180): on return
//   w.b0 = G1, w.b3 = G2, w.b1 = G1*diag(e), w.b2 = G2*diag(e),
//   w.b5 = C'_lower, w.b6 = C'_upper, w.b4 = LU(G1), w.lam, w.elz
// and R, T are filled.  Returns eigen failures.
template <int NC>
SSB_HD inline int diffuse_part(int n, double dz, const double *g1, const double *g2, double *R, double *T,
                               RadtoolWork<NC> &w) {
  const int nn = n * n;
  double *gdiff = w.b2, *P = w.b0, *V = w.b1;
  for (int i = 0; i < nn; ++i) {
    gdiff[i] = g1[i] - g2[i];
    w.b3[i] = g1[i] + g2[i];
  }
  mat_mul(n, n, n, gdiff, w.b3, P);
  double ev[NC];
  int nerr = eigen_real(n, P, ev, V, w.wk);
  for (int i = 0; i < n; ++i) {
    w.lam[i] = sqrt(dmax(0.0, ev[i]));
    w.elz[i] = exp(-w.lam[i] * dz);
  }
  // tmp = -(g1-g2)^-1 V, scaled by lambda per column
  double *tmp = w.b3;
  solve_mat(n, gdiff, V, tmp, w.b4);
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) tmp[i + n * j] = (-tmp[i + n * j]) * w.lam[j];
  double *G1 = w.b0, *G2 = w.b3;
  for (int i = 0; i < nn; ++i) {
    const double vv = V[i], tt = tmp[i];
    G1[i] = vv + tt;
    G2[i] = vv - tt;
  }
  double *G1d = w.b1, *G2d = w.b2;
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) {
      G1d[i + n * j] = G1[i + n * j] * w.elz[j];
      G2d[i + n * j] = G2[i + n * j] * w.elz[j];
    }
  // C'_lower = (G1 - G2d G1^-1 G2d)^-1 ; C'_upper = -G1^-1 G2d C'_lower
  double *X = w.b5, *S = w.b6, *Clo = w.b5, *Cup = w.b6;
  solve_mat(n, G1, G2d, X, w.b4);
  mat_mul(n, n, n, G2d, X, S);
  for (int i = 0; i < nn; ++i) S[i] = G1[i] - S[i];
  invert(n, S, Clo);                 // X dead, S destroyed
  mat_mul(n, n, n, G2d, Clo, w.b4);  // Y
  solve_mat(n, G1, w.b4, Cup, T);    // T used as LU scratch here
  for (int i = 0; i < nn; ++i) Cup[i] = -Cup[i];
  // R = G1d Cup + G2 Clo ; T = G2 Cup + G1d Clo (two products, then the sum)
  mat_mul(n, n, n, G1d, Cup, R);
  mat_mul(n, n, n, G2, Clo, w.b4);
  for (int i = 0; i < nn; ++i) R[i] = R[i] + w.b4[i];
  mat_mul(n, n, n, G2, Cup, T);
  mat_mul(n, n, n, G1d, Clo, w.b4);
  for (int i = 0; i < nn; ++i) T[i] = T[i] + w.b4[i];
  return nerr;
}

// calc_matrices_sw_eig: outputs R,T,Idiff (n x n), Sup,Sdn,Idd (n x d), E,Idir (d x d).
template <int NC>
SSB_HD inline int calc_matrices_sw(int n, int d, double dz, const double *g0, const double *g1,
                                   const double *g2, const double *g3, double *R, double *T, double *Sup,
                                   double *Sdn, double *E, double *Idir, double *Idiff, double *Idd,
                                   RadtoolWork<NC> &w) {
  const int nn = n * n;
  int nerr = diffuse_part<NC>(n, dz, g1, g2, R, T, w);
  // b0 = G1, b3 = G2 stay live until direct_diffuse; b1,b2,b4,b5,b6 are free now.

  // Section 3 (:225-229): E = G0 diag(exp(eps dz)) G0^-1
  double g0c[9], G0[9], G0i[9], eps[3], e0[3], t9[9], wk0[7];
  for (int i = 0; i < d * d; ++i) g0c[i] = g0[i];
  nerr += eigen_real(d, g0c, eps, G0, wk0);
  for (int i = 0; i < d * d; ++i) t9[i] = G0[i];
  invert(d, t9, G0i);
  for (int i = 0; i < d; ++i) e0[i] = exp(eps[i] * dz);
  for (int j = 0; j < d; ++j)
    for (int i = 0; i < d; ++i) t9[i + d * j] = G0[i + d * j] * e0[j];
  mat_mul(d, d, d, t9, G0i, E);

  // Section 4 (:232-253): particular solutions g3, g4 per direct eigen-mode
  mat_mul(n, d, d, g3, G0, w.g3g0);
  double *g1d = w.b1, *inv = w.b2, *Q = w.b4, *tmp = w.b5, *lu = w.b6;
  for (int jd = 0; jd < d; ++jd) {
    for (int i = 0; i < nn; ++i) g1d[i] = g1[i];
    for (int i = 0; i < n; ++i) g1d[i + n * i] = g1[i + n * i] + eps[jd];
    for (int i = 0; i < nn; ++i) lu[i] = g1d[i];
    invert(n, lu, inv);
    mat_mul(n, n, n, g2, inv, Q);
    mat_mul(n, n, n, Q, g2, tmp);
    for (int i = 0; i < nn; ++i) tmp[i] = g1[i] - tmp[i];
    for (int i = 0; i < n; ++i) tmp[i + n * i] = tmp[i + n * i] - eps[jd];
    for (int i = 0; i < n; ++i) Q[i + n * i] = Q[i + n * i] - 1.0;
    const double *colp = w.g3g0 + n * jd;
    mat_vec(n, n, Q, colp, w.rhs);
    solve_vec(n, tmp, w.rhs, w.g4c, inv);  // inv reused as LU scratch
    mat_vec(n, n, g2, w.g4c, w.rhs);
    for (int i = 0; i < n; ++i) w.rhs[i] = colp[i] + w.rhs[i];
    solve_vec(n, g1d, w.rhs, w.col, inv);
    for (int i = 0; i < n; ++i) {
      w.g4[i + n * jd] = w.g4c[i];
      w.g3[i + n * jd] = -w.col[i];
    }
  }

  // direct_diffuse_part (:303-386)
  {
    const int N = 2 * n + d;
    const double *G1 = w.b0, *G2 = w.b3;
    double *gd = w.big;
    for (int i = 0; i < N * N; ++i) gd[i] = 0.0;
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) {
        const double a = G1[i + n * j];
        const double b = G2[i + n * j] * w.elz[j];
        gd[i + N * j] = a;
        gd[(n + i) + N * j] = b;
        gd[i + N * (n + j)] = b;
        gd[(n + i) + N * (n + j)] = a;
      }
    for (int j = 0; j < d; ++j) {
      for (int i = 0; i < d; ++i) gd[(2 * n + i) + N * (2 * n + j)] = G0[i + d * j];
      for (int i = 0; i < n; ++i) {
        gd[i + N * (2 * n + j)] = w.g3[i + n * j] * e0[j];
        gd[(n + i) + N * (2 * n + j)] = w.g4[i + n * j];
      }
    }
    lu_factor(N, gd);
    double *cp = w.cp;  // N x d
    for (int j = 0; j < d; ++j) {
      for (int i = 0; i < N; ++i) cp[i + N * j] = 0.0;
      cp[(2 * n + j) + N * j] = 1.0;
      lu_subst(N, gd, cp + N * j);
    }
    // Sup = [G1d G2 G3] C ; Sdn = [G2 G1d G4 diag(e0)] C, accumulated in block order
    for (int j = 0; j < d; ++j)
      for (int i = 0; i < n; ++i) {
        double su = 0.0, sd = 0.0;
        for (int k = 0; k < n; ++k) {
          su = su + (G1[i + n * k] * w.elz[k]) * cp[k + N * j];
          sd = sd + G2[i + n * k] * cp[k + N * j];
        }
        for (int k = 0; k < n; ++k) {
          su = su + G2[i + n * k] * cp[(n + k) + N * j];
          sd = sd + (G1[i + n * k] * w.elz[k]) * cp[(n + k) + N * j];
        }
        for (int k = 0; k < d; ++k) {
          su = su + w.g3[i + n * k] * cp[(2 * n + k) + N * j];
          sd = sd + (w.g4[i + n * k] * e0[k]) * cp[(2 * n + k) + N * j];
        }
        Sup[i + n * j] = su;
        Sdn[i + n * j] = sd;
      }
  }

  // Schur-complement inverse of the full Gamma matrix (radtool_schur.F90:45-51)
  {
    double *X = w.b0, *S = w.b1, *g1i = w.b2, *g1inv = w.b3, *g2i = w.b4, *lu2 = w.b5, *tmp2 = w.b6;
    double g0i[9], t0[9];
    for (int i = 0; i < d * d; ++i) t0[i] = g0[i];
    invert(d, t0, g0i);
    solve_mat(n, g1, g2, X, lu2);
    mat_mul(n, n, n, g2, X, S);
    for (int i = 0; i < nn; ++i) S[i] = g1[i] - S[i];
    invert(n, S, g1i);
    for (int i = 0; i < nn; ++i) lu2[i] = g1[i];
    invert(n, lu2, g1inv);
    mat_mul(n, n, n, g2, g1inv, tmp2);
    mat_mul(n, n, n, g1i, tmp2, g2i);
    // g3i = (g1i - g2i) (g3 g0i)
    double *g3g0i = w.g3g0;
    mat_mul(n, d, d, g3, g0i, g3g0i);
    for (int i = 0; i < nn; ++i) tmp2[i] = g1i[i] - g2i[i];
    mat_mul(n, n, d, tmp2, g3g0i, Idd);
    for (int i = 0; i < n * d; ++i) Idd[i] = 2.0 * Idd[i];
    for (int i = 0; i < d * d; ++i) Idir[i] = -g0i[i];
    for (int i = 0; i < nn; ++i) Idiff[i] = g2i[i] - g1i[i];
  }
  return nerr;
}

// calc_matrices_lw_eig: outputs R,T,IF (n x n), src, isrc (n).
template <int NC>
SSB_HD inline int calc_matrices_lw(int n, double dz, const double *g1, const double *g2, const double *b,
                                   double *R, double *T, double *src, double *IF, double *isrc,
                                   RadtoolWork<NC> &w) {
  const int nn = n * n;
  int nerr = diffuse_part<NC>(n, dz, g1, g2, R, T, w);
  // live: b0 = G1, b3 = G2, b1 = G1d, b2 = G2d, b5 = Clo, b6 = Cup ; free: b4, big
  const double *G1 = w.b0, *G2 = w.b3, *G1d = w.b1, *G2d = w.b2, *Clo = w.b5, *Cup = w.b6;
  double *Q = w.b4, *inv = w.big, *tmp = w.big + nn, *lu = w.big + 2 * nn;
  // y = (g1 - Q g2)^-1 (Q - I) b with Q = g2 g1^-1 (:188-197)
  for (int i = 0; i < nn; ++i) lu[i] = g1[i];
  invert(n, lu, inv);
  mat_mul(n, n, n, g2, inv, Q);
  mat_mul(n, n, n, Q, g2, tmp);
  for (int i = 0; i < nn; ++i) tmp[i] = g1[i] - tmp[i];
  for (int i = 0; i < n; ++i) Q[i + n * i] = Q[i + n * i] - 1.0;
  mat_vec(n, n, Q, b, w.rhs);
  double *y = w.g4c;
  solve_vec(n, tmp, w.rhs, y, lu);
  // c_b = -Clo (y - G2d G1^-1 y) (:206-208)
  solve_vec(n, G1, y, w.col, lu);
  mat_vec(n, n, G2d, w.col, w.rhs);
  for (int i = 0; i < n; ++i) w.rhs[i] = y[i] - w.rhs[i];
  double *cb = w.col;
  {
    double t[NC];
    mat_vec(n, n, Clo, w.rhs, t);
    for (int i = 0; i < n; ++i) cb[i] = -t[i];
  }
  // source = (G1d + G2) c_b + y (:211)
  for (int i = 0; i < n; ++i) {
    double s = 0.0;
    for (int j = 0; j < n; ++j) s = s + (G1d[i + n * j] + G2[i + n * j]) * cb[j];
    src[i] = s + y[i];
  }
  // integrated fluxes (:213-227)
  double zf[NC];
  for (int i = 0; i < n; ++i) zf[i] = (1.0 - w.elz[i]) / w.lam[i];
  double *GZ = Q;
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) GZ[i + n * j] = G1[i + n * j] * zf[j] + G2[i + n * j] * zf[j];
  for (int i = 0; i < nn; ++i) tmp[i] = Clo[i] + Cup[i];
  mat_mul(n, n, n, GZ, tmp, IF);
  for (int i = 0; i < n; ++i) {
    double s = 0.0;
    for (int j = 0; j < n; ++j) s = s + GZ[i + n * j] * cb[j];
    isrc[i] = 2.0 * (s + y[i] * dz);
  }
  return nerr;
}

}  // namespace ssb
