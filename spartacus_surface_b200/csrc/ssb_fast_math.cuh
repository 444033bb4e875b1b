// ssb_fast_math.cuh - register-resident evaluation of the layer transfer
// matrices for compile-time orders (NR regions x NS streams).
//
// Computes the same outputs as calc_matrices_sw_eig / calc_matrices_lw_eig
// (radtool/radtool_calc_matrices_sw_eig.F90:30-386, _lw_eig.F90:32-230,
// radtool/radtool_schur.F90:32-53) but through an equivalent formulation that
// suits one-thread-per-problem execution on the GPU:
//
//  * Gamma1, Gamma2 are "N-symmetric": Gamma * diag(1/N) is symmetric with
//    N_i = 1 / (weight_js * mu_js * frac_r) (detailed balance of the exchange
//    and scattering terms).  With -N(G1-G2) = L L^T (Cholesky) the eigenproblem
//    of P = (G1-G2)(G1+G2) becomes the SYMMETRIC problem Y = L^T K L,
//    K = -(G1+G2)/N, solved by cyclic Jacobi: no data-dependent control flow,
//    all indices static, eigenvectors V = N^-1 L U with V^-1 = U^T L^-1 N for
//    free, and (G1-G2)^-1 V = -L^-T U without another factorisation.
//  * The 2n x 2n two-point boundary problem is block-symmetric, so it splits
//    into sum and difference problems of order n:
//      R + T = B+ A+^-1,  R - T = -B- A-^-1,
//      A± = V(1±e) + M(1∓e),  B± = ±[V(1±e) - M(1∓e)]   (e = exp(-lambda dz)).
//    The reference's (2n+d)^2 direct-diffuse system reduces to products with
//    R±T, its per-eigenmode inversions to (eps^2 - P)^-1 through the
//    eigenvectors, and the Schur inverse to (G1+G2)^-1.
// Results agree with the reference route to rounding sensitivity (tests).
#pragma once
#include "ssb_small.cuh"

namespace ssb {

template <int N>
struct FastDiffuse {
  double V[N * N];    // eigenvectors of P
  double M[N * N];    // -(G1-G2)^-1 V diag(lambda)
  double LUp[N * N];  // LU of A+ (reciprocal pivots on the diagonal)
  double Xp[N * N];   // R + T
  double Xm[N * N];   // -(R - T)
  double lam[N], e[N];
};

constexpr int kJacobiSweeps(int n) { return n <= 2 ? 3 : (n <= 4 ? 7 : (n <= 6 ? 8 : 10)); }

// Common part: eigen-system of P and R±T.  `nsc[i]` = N_i.  On return also
// L (Cholesky factor of -N D, lower), U (Jacobi vectors) and Linv_diag for the
// caller's use of V^-1.
template <int N>
SSB_HDI void fast_diffuse(double dz, const double *g1, const double *g2, const double *nsc, FastDiffuse<N> &o,
                          double *L, double *Ldinv, double *U, double *lam2) {
  double Y[N * N];
  // L <- -N (G1 - G2)   (lower triangle), K <- -(G1 + G2) / N  (symmetrised)
  SSB_UNROLL
  for (int j = 0; j < N; ++j) {
    SSB_UNROLL
    for (int i = 0; i < N; ++i) {
      L[i + N * j] = -nsc[i] * (g1[i + N * j] - g2[i + N * j]);
      Y[i + N * j] = -(g1[i + N * j] + g2[i + N * j]) / nsc[j];
    }
  }
  sm_cholesky<N>(L, Ldinv);
  // Y <- L^T K L using the lower triangles only
  {
    double T1[N * N];  // K L  (column j needs rows k >= j of L)
    SSB_UNROLL
    for (int j = 0; j < N; ++j) {
      SSB_UNROLL
      for (int i = 0; i < N; ++i) {
        double s = 0.0;
        SSB_UNROLL
        for (int k = j; k < N; ++k) {
          // K is symmetric up to rounding: read the lower triangle
          const double kik = (i >= k) ? Y[i + N * k] : Y[k + N * i];
          s = fma(kik, L[k + N * j], s);
        }
        T1[i + N * j] = s;
      }
    }
    SSB_UNROLL
    for (int j = 0; j < N; ++j) {
      SSB_UNROLL
      for (int i = j; i < N; ++i) {
        double s = 0.0;
        SSB_UNROLL
        for (int k = i; k < N; ++k) s = fma(L[k + N * i], T1[k + N * j], s);
        Y[i + N * j] = s;
        Y[j + N * i] = s;
      }
    }
  }
  sm_jacobi<N>(Y, U, lam2, kJacobiSweeps(N));
  SSB_UNROLL
  for (int k = 0; k < N; ++k) {
    o.lam[k] = sqrt(dmax(0.0, lam2[k]));
    o.e[k] = exp(-o.lam[k] * dz);
  }
  // V = N^-1 L U ; M = L^-T U diag(lambda)
  SSB_UNROLL
  for (int k = 0; k < N; ++k) {
    SSB_UNROLL
    for (int i = 0; i < N; ++i) {
      double s = 0.0;
      SSB_UNROLL
      for (int j = 0; j <= i; ++j) s = fma(L[i + N * j], U[j + N * k], s);
      o.V[i + N * k] = s / nsc[i];
    }
    // back substitution L^T m = u
    SSB_UNROLL
    for (int i = N - 1; i >= 0; --i) {
      double s = U[i + N * k];
      SSB_UNROLL
      for (int j = i + 1; j < N; ++j) s = fma(-L[j + N * i], o.M[j + N * k], s);
      o.M[i + N * k] = s * Ldinv[i];
    }
    SSB_UNROLL
    for (int i = 0; i < N; ++i) o.M[i + N * k] *= o.lam[k];
  }
  // sum problem
  double Am[N * N];
  SSB_UNROLL
  for (int k = 0; k < N; ++k) {
    const double ep = 1.0 + o.e[k], em = 1.0 - o.e[k];
    SSB_UNROLL
    for (int i = 0; i < N; ++i) {
      const double v = o.V[i + N * k], m = o.M[i + N * k];
      o.LUp[i + N * k] = fma(v, ep, m * em);
      o.Xp[i + N * k] = fma(v, ep, -(m * em));
      Am[i + N * k] = fma(m, ep, v * em);
      o.Xm[i + N * k] = fma(m, ep, -(v * em));
    }
  }
  sm_lu<N>(o.LUp);
  sm_lu_solve_right<N, N>(o.LUp, o.Xp);
  sm_lu<N>(Am);
  sm_lu_solve_right<N, N>(Am, o.Xm);
}

// Shortwave layer: outputs as in calc_matrices_sw_eig (orders n = NR*NS, d = NR).
// `frac[r]` are the region fractions of the solved regions.
template <int NR, int NS>
SSB_HDI void fast_layer_sw_math(double dz, const double *g0, const double *g1, const double *g2, const double *g3,
                                const double *nsc, const double *frac, double *R, double *T, double *Sup,
                                double *Sdn, double *E, double *Idir, double *Idiff, double *Idd) {
  constexpr int N = NR * NS, D = NR;
  FastDiffuse<N> o;
  double L[N * N], Ldinv[N], U[N * N], lam2[N];
  fast_diffuse<N>(dz, g1, g2, nsc, o, L, Ldinv, U, lam2);
  SSB_UNROLL
  for (int i = 0; i < N * N; ++i) {
    R[i] = 0.5 * (o.Xp[i] - o.Xm[i]);
    T[i] = 0.5 * (o.Xp[i] + o.Xm[i]);
  }
  // direct beam: g0 = B0 diag(1/frac) with B0 symmetric -> symmetric Jacobi
  double G0[D * D], G0i[D * D], eps[D], e0[D];
  {
    double Y0[D * D], U0[D * D], sq[D];
    SSB_UNROLL
    for (int r = 0; r < D; ++r) sq[r] = sqrt(frac[r]);
    SSB_UNROLL
    for (int j = 0; j < D; ++j) {
      SSB_UNROLL
      for (int i = 0; i < D; ++i) {
        // N0^(1/2) g0 N0^(-1/2), symmetrised from the lower triangle
        const double lo = (i >= j) ? g0[i + D * j] * sq[j] / sq[i] : g0[j + D * i] * sq[i] / sq[j];
        Y0[i + D * j] = lo;
      }
    }
    sm_jacobi<D>(Y0, U0, eps, kJacobiSweeps(D) + 2);
    SSB_UNROLL
    for (int k = 0; k < D; ++k) {
      e0[k] = exp(eps[k] * dz);
      SSB_UNROLL
      for (int i = 0; i < D; ++i) {
        G0[i + D * k] = sq[i] * U0[i + D * k];
        G0i[k + D * i] = U0[i + D * k] / sq[i];
      }
    }
  }
  double g0inv[D * D];
  SSB_UNROLL
  for (int j = 0; j < D; ++j) {
    SSB_UNROLL
    for (int i = 0; i < D; ++i) {
      double se = 0.0, si = 0.0;
      SSB_UNROLL
      for (int k = 0; k < D; ++k) {
        const double gk = G0[i + D * k] * G0i[k + D * j];
        se = fma(gk, e0[k], se);
        si = fma(gk, 1.0 / eps[k], si);
      }
      E[i + D * j] = se;
      g0inv[i + D * j] = si;
      Idir[i + D * j] = -si;
    }
  }
  // particular solutions per direct eigen-mode:
  //   a = g3p + g4p = 2 V (eps^2 - Lambda)^-1 V^-1 (G1-G2) c ,  b = g3p - g4p = -((G1+G2) a + 2c)/eps
  double c[N * D], G3p[N * D], G4p[N * D];
  sm_mul<N, D, D>(g3, G0, c);
  SSB_UNROLL
  for (int jd = 0; jd < D; ++jd) {
    double w[N], t[N], a[N];
    SSB_UNROLL
    for (int i = 0; i < N; ++i) {
      double s = 0.0;
      SSB_UNROLL
      for (int k = 0; k < N; ++k) s = fma(g1[i + N * k] - g2[i + N * k], c[k + N * jd], s);
      w[i] = s * nsc[i];
    }
    // t = U^T L^-1 w
    SSB_UNROLL
    for (int i = 0; i < N; ++i) {
      double s = w[i];
      SSB_UNROLL
      for (int k = 0; k < i; ++k) s = fma(-L[i + N * k], w[k], s);
      w[i] = s * Ldinv[i];
    }
    SSB_UNROLL
    for (int k = 0; k < N; ++k) {
      double s = 0.0;
      SSB_UNROLL
      for (int i = 0; i < N; ++i) s = fma(U[i + N * k], w[i], s);
      t[k] = 2.0 * s / (eps[jd] * eps[jd] - lam2[k]);
    }
    sm_mulvec<N, N>(o.V, t, a);
    SSB_UNROLL
    for (int i = 0; i < N; ++i) {
      double s = 2.0 * c[i + N * jd];
      SSB_UNROLL
      for (int k = 0; k < N; ++k) s = fma(g1[i + N * k] + g2[i + N * k], a[k], s);
      const double b = -s / eps[jd];
      G3p[i + N * jd] = 0.5 * (a[i] + b);
      G4p[i + N * jd] = 0.5 * (a[i] - b);
    }
  }
  // S_up ± S_dn
  {
    double rp[N * D], rm[N * D], qp[N * D], qm[N * D];
    SSB_UNROLL
    for (int j = 0; j < D; ++j) {
      SSB_UNROLL
      for (int i = 0; i < N; ++i) {
        double sp = 0.0, sm = 0.0, tp = 0.0, tm = 0.0;
        SSB_UNROLL
        for (int k = 0; k < D; ++k) {
          const double gi = G0i[k + D * j];
          const double g3e = G3p[i + N * k] * e0[k], g4e = G4p[i + N * k] * e0[k];
          sp = fma(-(g3e + G4p[i + N * k]), gi, sp);  // r1 + r2
          sm = fma(-(g3e - G4p[i + N * k]), gi, sm);  // r1 - r2
          tp = fma(G3p[i + N * k] + g4e, gi, tp);
          tm = fma(G3p[i + N * k] - g4e, gi, tm);
        }
        rp[i + N * j] = sp;
        rm[i + N * j] = sm;
        qp[i + N * j] = tp;
        qm[i + N * j] = tm;
      }
    }
    double up[N * D], um[N * D];
    sm_mul<N, N, D>(o.Xp, rp, up);
    sm_mul<N, N, D>(o.Xm, rm, um);
    SSB_UNROLL
    for (int i = 0; i < N * D; ++i) {
      const double sum = up[i] + qp[i], dif = um[i] + qm[i];
      Sup[i] = 0.5 * (sum + dif);
      Sdn[i] = 0.5 * (sum - dif);
    }
  }
  // integrated-flux matrices: Idiff = -(G1+G2)^-1, Idd = 2 (G1+G2)^-1 G3 G0^-1
  {
    double Sm[N * N];
    SSB_UNROLL
    for (int i = 0; i < N * N; ++i) Sm[i] = g1[i] + g2[i];
    sm_lu<N>(Sm);
    SSB_UNROLL
    for (int j = 0; j < N; ++j) {
      SSB_UNROLL
      for (int i = 0; i < N; ++i) Idiff[i + N * j] = (i == j) ? -1.0 : 0.0;
    }
    sm_lu_solve_left<N, N>(Sm, Idiff);
    double g3g0i[N * D];
    sm_mul<N, D, D>(g3, g0inv, g3g0i);
    sm_mul<N, N, D>(Idiff, g3g0i, Idd);
    SSB_UNROLL
    for (int i = 0; i < N * D; ++i) Idd[i] = -2.0 * Idd[i];
  }
}

// Longwave layer: outputs as in calc_matrices_lw_eig.
template <int NR, int NS>
SSB_HDI void fast_layer_lw_math(double dz, const double *g1, const double *g2, const double *b, const double *nsc,
                                double *R, double *T, double *src, double *IF, double *isrc) {
  constexpr int N = NR * NS;
  FastDiffuse<N> o;
  double L[N * N], Ldinv[N], U[N * N], lam2[N];
  fast_diffuse<N>(dz, g1, g2, nsc, o, L, Ldinv, U, lam2);
  SSB_UNROLL
  for (int i = 0; i < N * N; ++i) {
    R[i] = 0.5 * (o.Xp[i] - o.Xm[i]);
    T[i] = 0.5 * (o.Xp[i] + o.Xm[i]);
  }
  // y = -(G1+G2)^-1 b ; source = y - (R+T) y
  double y[N];
  {
    double Sm[N * N];
    SSB_UNROLL
    for (int i = 0; i < N * N; ++i) Sm[i] = g1[i] + g2[i];
    sm_lu<N>(Sm);
    SSB_UNROLL
    for (int i = 0; i < N; ++i) y[i] = -b[i];
    sm_lu_solve_left<N, 1>(Sm, y);
  }
  double xy[N];
  sm_mulvec<N, N>(o.Xp, y, xy);
  SSB_UNROLL
  for (int i = 0; i < N; ++i) src[i] = y[i] - xy[i];
  // int_flux = 2 V Z A+^-1 ; int_flux_source = 2 y dz - 2 int_flux y
  SSB_UNROLL
  for (int k = 0; k < N; ++k) {
    const double z = 2.0 * (1.0 - o.e[k]) / o.lam[k];
    SSB_UNROLL
    for (int i = 0; i < N; ++i) IF[i + N * k] = o.V[i + N * k] * z;
  }
  sm_lu_solve_right<N, N>(o.LUp, IF);
  double fy[N];
  sm_mulvec<N, N>(IF, y, fy);
  SSB_UNROLL
  for (int i = 0; i < N; ++i) isrc[i] = 2.0 * (y[i] * dz - fy[i]);
}

}  // namespace ssb
