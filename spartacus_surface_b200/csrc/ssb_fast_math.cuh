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
  double V[N * N];   // eigenvectors of P = D S
  double L[N * N];   // Cholesky factor of -N D (lower triangle)
  double U[N * N];   // Jacobi eigenvectors of L^T K L
  double Ldinv[N], lam[N], lam2[N], e[N];
};

SSB_HD constexpr int kJacobiSweeps(int n) { return n <= 2 ? 4 : 12; }

// Eigen-system of P = D S with D = G1-G2, S = G1+G2 (both N-symmetric), then
// R and T from the sum and difference problems.  `ninv[i]` = 1/N_i = w mu frac,
// `nsc[i]` = N_i.  When IF is non-null it receives 2 V Z A+^-1 (longwave).
template <int N>
SSB_HDI void fast_diffuse(double dz, const double *Dm, const double *Sm, const double *nsc, const double *ninv,
                          FastDiffuse<N> &o, double *R, double *T, double *IF) {
  {
    double Y[N * N];
    // L <- -N D (lower triangle used), Y <- K = -S / N (symmetric up to rounding)
    SSB_UNROLL
    for (int j = 0; j < N; ++j) {
      SSB_UNROLL
      for (int i = 0; i < N; ++i) {
        o.L[i + N * j] = -nsc[i] * Dm[i + N * j];
        Y[i + N * j] = -Sm[i + N * j] * ninv[j];
      }
    }
    sm_cholesky<N>(o.L, o.Ldinv);
    // Y <- L^T K L using the lower triangles only
    {
      double T1[N * N];
      SSB_UNROLL
      for (int j = 0; j < N; ++j) {
        SSB_UNROLL
        for (int i = 0; i < N; ++i) {
          double s = 0.0;
          SSB_UNROLL
          for (int k = j; k < N; ++k) {
            const double kik = (i >= k) ? Y[i + N * k] : Y[k + N * i];
            s = fma(kik, o.L[k + N * j], s);
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
          for (int k = i; k < N; ++k) s = fma(o.L[k + N * i], T1[k + N * j], s);
          Y[i + N * j] = s;
          Y[j + N * i] = s;
        }
      }
    }
    sm_jacobi<N>(Y, o.U, o.lam2, kJacobiSweeps(N));
  }
  SSB_UNROLL
  for (int k = 0; k < N; ++k) {
    o.lam[k] = sqrt(dmax(0.0, o.lam2[k]));
    o.e[k] = exp(-o.lam[k] * dz);
  }
  // V = N^-1 L U ; M = L^-T U diag(lambda)
  double M[N * N];
  SSB_UNROLL
  for (int k = 0; k < N; ++k) {
    SSB_UNROLL
    for (int i = 0; i < N; ++i) {
      double s = 0.0;
      SSB_UNROLL
      for (int j = 0; j <= i; ++j) s = fma(o.L[i + N * j], o.U[j + N * k], s);
      o.V[i + N * k] = s * ninv[i];
    }
    SSB_UNROLL
    for (int i = N - 1; i >= 0; --i) {
      double s = o.U[i + N * k];
      SSB_UNROLL
      for (int j = i + 1; j < N; ++j) s = fma(-o.L[j + N * i], M[j + N * k], s);
      M[i + N * k] = s * o.Ldinv[i];
    }
    SSB_UNROLL
    for (int i = 0; i < N; ++i) M[i + N * k] *= o.lam[k];
  }
  // sum (sigma=+1) and difference (sigma=-1) problems:
  //   A = V(1+sigma e) + M(1-sigma e), B = V(1+sigma e) - M(1-sigma e), X = B A^-1
  //   R = (X+ + X-)/2, T = (X+ - X-)/2
  SSB_UNROLL
  for (int i = 0; i < N * N; ++i) {
    R[i] = 0.0;
    T[i] = 0.0;
  }
#if defined(__CUDACC__)
#pragma unroll 1
#endif
  for (int sg = 0; sg < 2; ++sg) {
    const double sigma = sg ? -1.0 : 1.0;
    double A[N * N], B[N * N];
    SSB_UNROLL
    for (int k = 0; k < N; ++k) {
      const double ep = fma(sigma, o.e[k], 1.0), em = fma(-sigma, o.e[k], 1.0);
      SSB_UNROLL
      for (int i = 0; i < N; ++i) {
        const double v = o.V[i + N * k] * ep, m = M[i + N * k] * em;
        A[i + N * k] = v + m;
        B[i + N * k] = v - m;
      }
    }
    sm_lu<N>(A);
    sm_lu_solve_right<N, N>(A, B);
    const double hs = 0.5 * sigma;
    SSB_UNROLL
    for (int i = 0; i < N * N; ++i) {
      R[i] = fma(0.5, B[i], R[i]);
      T[i] = fma(hs, B[i], T[i]);
    }
    if (IF != nullptr && sg == 0) {
      SSB_UNROLL
      for (int k = 0; k < N; ++k) {
        const double z = 2.0 * (1.0 - o.e[k]) / o.lam[k];
        SSB_UNROLL
        for (int i = 0; i < N; ++i) IF[i + N * k] = o.V[i + N * k] * z;
      }
      sm_lu_solve_right<N, N>(A, IF);
    }
  }
}

// Shortwave layer: outputs as in calc_matrices_sw_eig (orders n = NR*NS, d = NR).
// Dm = G1-G2, Sm = G1+G2; `frac[r]` are the region fractions of the solved regions.
template <int NR, int NS>
SSB_HDI void fast_layer_sw_math(double dz, const double *g0, const double *Dm, const double *Sm, const double *g3,
                                const double *nsc, const double *ninv, const double *frac, double *R, double *T,
                                double *Sup, double *Sdn, double *E, double *Idir, double *Idiff, double *Idd) {
  constexpr int N = NR * NS, D = NR;
  FastDiffuse<N> o;
  fast_diffuse<N>(dz, Dm, Sm, nsc, ninv, o, R, T, nullptr);
  // direct beam: g0 = B0 diag(1/frac) with B0 symmetric -> symmetric Jacobi
  double G0[D * D], G0i[D * D], eps[D], e0[D];
  {
    double Y0[D * D], U0[D * D], sq[D], rsq[D];
    SSB_UNROLL
    for (int r = 0; r < D; ++r) {
      sq[r] = sqrt(frac[r]);
      rsq[r] = 1.0 / sq[r];
    }
    SSB_UNROLL
    for (int j = 0; j < D; ++j) {
      SSB_UNROLL
      for (int i = 0; i < D; ++i)
        Y0[i + D * j] = (i >= j) ? g0[i + D * j] * sq[j] * rsq[i] : g0[j + D * i] * sq[i] * rsq[j];
    }
    sm_jacobi<D>(Y0, U0, eps, kJacobiSweeps(D));
    SSB_UNROLL
    for (int k = 0; k < D; ++k) {
      e0[k] = exp(eps[k] * dz);
      SSB_UNROLL
      for (int i = 0; i < D; ++i) {
        G0[i + D * k] = sq[i] * U0[i + D * k];
        G0i[k + D * i] = U0[i + D * k] * rsq[i];
      }
    }
  }
  double g0inv[D * D], reps[D];
  SSB_UNROLL
  for (int k = 0; k < D; ++k) reps[k] = 1.0 / eps[k];
  SSB_UNROLL
  for (int j = 0; j < D; ++j) {
    SSB_UNROLL
    for (int i = 0; i < D; ++i) {
      double se = 0.0, si = 0.0;
      SSB_UNROLL
      for (int k = 0; k < D; ++k) {
        const double gk = G0[i + D * k] * G0i[k + D * j];
        se = fma(gk, e0[k], se);
        si = fma(gk, reps[k], si);
      }
      E[i + D * j] = se;
      g0inv[i + D * j] = si;
      Idir[i + D * j] = -si;
    }
  }
  // particular solutions per direct eigen-mode:
  //   a = g3p + g4p = 2 V (eps^2 - Lambda)^-1 V^-1 D c ,  b = g3p - g4p = -(S a + 2c)/eps
  // and the source terms  S_up ± S_dn = ±(R±T)(r1±r2) + (G3p ± G4p e0) G0^-1
  double rp[N * D], rm[N * D], qp[N * D], qm[N * D];
  SSB_UNROLL
  for (int i = 0; i < N * D; ++i) rp[i] = rm[i] = qp[i] = qm[i] = 0.0;
  SSB_UNROLL
  for (int jd = 0; jd < D; ++jd) {
    double c[N], w[N], t[N], a[N];
    SSB_UNROLL
    for (int i = 0; i < N; ++i) {
      double s = 0.0;
      SSB_UNROLL
      for (int k = 0; k < D; ++k) s = fma(g3[i + N * k], G0[k + D * jd], s);
      c[i] = s;
    }
    SSB_UNROLL
    for (int i = 0; i < N; ++i) {
      double s = 0.0;
      SSB_UNROLL
      for (int k = 0; k < N; ++k) s = fma(Dm[i + N * k], c[k], s);
      w[i] = s * nsc[i];
    }
    // t = U^T L^-1 w, scaled by 2 / (eps^2 - lambda^2)
    SSB_UNROLL
    for (int i = 0; i < N; ++i) {
      double s = w[i];
      SSB_UNROLL
      for (int k = 0; k < i; ++k) s = fma(-o.L[i + N * k], w[k], s);
      w[i] = s * o.Ldinv[i];
    }
    SSB_UNROLL
    for (int k = 0; k < N; ++k) {
      double s = 0.0;
      SSB_UNROLL
      for (int i = 0; i < N; ++i) s = fma(o.U[i + N * k], w[i], s);
      t[k] = 2.0 * s / fma(eps[jd], eps[jd], -o.lam2[k]);
    }
    sm_mulvec<N, N>(o.V, t, a);
    const double mreps = -reps[jd];
    SSB_UNROLL
    for (int i = 0; i < N; ++i) {
      double s = 2.0 * c[i];
      SSB_UNROLL
      for (int k = 0; k < N; ++k) s = fma(Sm[i + N * k], a[k], s);
      const double b = s * mreps;
      const double g3p = 0.5 * (a[i] + b), g4p = 0.5 * (a[i] - b);
      const double g3e = g3p * e0[jd], g4e = g4p * e0[jd];
      SSB_UNROLL
      for (int j = 0; j < D; ++j) {
        const double gi = G0i[jd + D * j];
        rp[i + N * j] = fma(-(g3e + g4p), gi, rp[i + N * j]);  // r1 + r2
        rm[i + N * j] = fma(-(g3e - g4p), gi, rm[i + N * j]);  // r1 - r2
        qp[i + N * j] = fma(g3p + g4e, gi, qp[i + N * j]);
        qm[i + N * j] = fma(g3p - g4e, gi, qm[i + N * j]);
      }
    }
  }
  // S_up + S_dn = (R+T) rp + qp ; S_up - S_dn = (T-R) rm + qm
  SSB_UNROLL
  for (int j = 0; j < D; ++j) {
    SSB_UNROLL
    for (int i = 0; i < N; ++i) {
      double sum = qp[i + N * j], dif = qm[i + N * j];
      SSB_UNROLL
      for (int k = 0; k < N; ++k) {
        const double r = R[i + N * k], tt = T[i + N * k];
        sum = fma(r + tt, rp[k + N * j], sum);
        dif = fma(tt - r, rm[k + N * j], dif);
      }
      Sup[i + N * j] = 0.5 * (sum + dif);
      Sdn[i + N * j] = 0.5 * (sum - dif);
    }
  }
  // integrated-flux matrices: Idiff = -S^-1, Idd = 2 S^-1 G3 G0^-1
  {
    double LUs[N * N];
    SSB_UNROLL
    for (int i = 0; i < N * N; ++i) LUs[i] = Sm[i];
    sm_lu<N>(LUs);
    SSB_UNROLL
    for (int j = 0; j < N; ++j) {
      SSB_UNROLL
      for (int i = 0; i < N; ++i) Idiff[i + N * j] = (i == j) ? -1.0 : 0.0;
    }
    sm_lu_solve_left<N, N>(LUs, Idiff);
    double g3g0i[N * D];
    sm_mul<N, D, D>(g3, g0inv, g3g0i);
    sm_mul<N, N, D>(Idiff, g3g0i, Idd);
    SSB_UNROLL
    for (int i = 0; i < N * D; ++i) Idd[i] = -2.0 * Idd[i];
  }
}

// Longwave layer: outputs as in calc_matrices_lw_eig.
template <int NR, int NS>
SSB_HDI void fast_layer_lw_math(double dz, const double *Dm, const double *Sm, const double *b, const double *nsc,
                                const double *ninv, double *R, double *T, double *src, double *IF, double *isrc) {
  constexpr int N = NR * NS;
  FastDiffuse<N> o;
  fast_diffuse<N>(dz, Dm, Sm, nsc, ninv, o, R, T, IF);
  // y = -S^-1 b ; source = y - (R+T) y ; int_flux_source = 2 y dz - 2 int_flux y
  double y[N];
  {
    double LUs[N * N];
    SSB_UNROLL
    for (int i = 0; i < N * N; ++i) LUs[i] = Sm[i];
    sm_lu<N>(LUs);
    SSB_UNROLL
    for (int i = 0; i < N; ++i) y[i] = -b[i];
    sm_lu_solve_left<N, 1>(LUs, y);
  }
  SSB_UNROLL
  for (int i = 0; i < N; ++i) {
    double s = 0.0, f = 0.0;
    SSB_UNROLL
    for (int k = 0; k < N; ++k) {
      s = fma(R[i + N * k] + T[i + N * k], y[k], s);
      f = fma(IF[i + N * k], y[k], f);
    }
    src[i] = y[i] - s;
    isrc[i] = 2.0 * (y[i] * dz - f);
  }
}

}  // namespace ssb
