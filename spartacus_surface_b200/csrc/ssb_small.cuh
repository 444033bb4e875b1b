// ssb_small.cuh - compile-time-sized dense algebra for the register-resident
// ("fast") kernels: every loop has constant bounds and is fully unrolled, every
// index is static, so a matrix lives in registers of the one thread that owns
// the problem.  No dynamic control flow except a fixed number of Jacobi sweeps.
//
// These routines back ssb_layer_math.cuh, which evaluates the same layer
// quantities as radtool_calc_matrices_{sw,lw}_eig.F90 through an algebraically
// equivalent but cheaper and better conditioned route.
#pragma once
#include "ssb_math.cuh"

#if defined(__CUDACC__)
#define SSB_UNROLL _Pragma("unroll")
#define SSB_ROLLED _Pragma("unroll 1")  // keep a loop rolled: the code must fit the instruction cache
#else
#define SSB_UNROLL
#define SSB_ROLLED
#endif

namespace ssb {

// C (R x C) = A (R x K) * B (K x C)
template <int R, int K, int C>
SSB_HDI void sm_mul(const double *A, const double *B, double *Cm) {
  SSB_UNROLL
  for (int j = 0; j < C; ++j) {
    SSB_UNROLL
    for (int i = 0; i < R; ++i) {
      double s = 0.0;
      SSB_UNROLL
      for (int k = 0; k < K; ++k) s = fma(A[i + R * k], B[k + K * j], s);
      Cm[i + R * j] = s;
    }
  }
}

// y (R) = A (R x C) x (C)
template <int R, int C>
SSB_HDI void sm_mulvec(const double *A, const double *x, double *y) {
  SSB_UNROLL
  for (int i = 0; i < R; ++i) {
    double s = 0.0;
    SSB_UNROLL
    for (int j = 0; j < C; ++j) s = fma(A[i + R * j], x[j], s);
    y[i] = s;
  }
}

// In-place LU without pivoting, Doolittle (unit lower factor); the reciprocals
// of the pivots are stored on the diagonal so that solves need no division.
template <int N>
SSB_HDI void sm_lu(double *A) {
  SSB_UNROLL
  for (int k = 0; k < N; ++k) {
    const double inv = 1.0 / A[k + N * k];
    A[k + N * k] = inv;
    SSB_UNROLL
    for (int i = k + 1; i < N; ++i) A[i + N * k] *= inv;
    SSB_UNROLL
    for (int j = k + 1; j < N; ++j) {
      const double akj = A[k + N * j];
      SSB_UNROLL
      for (int i = k + 1; i < N; ++i) A[i + N * j] = fma(-A[i + N * k], akj, A[i + N * j]);
    }
  }
}

// Solve A X = B in place (B: N x C) with the factors of sm_lu.
template <int N, int C>
SSB_HDI void sm_lu_solve_left(const double *LU, double *B) {
  SSB_UNROLL
  for (int j = 0; j < C; ++j) {
    double *x = B + N * j;
    SSB_UNROLL
    for (int i = 1; i < N; ++i) {
      double s = x[i];
      SSB_UNROLL
      for (int k = 0; k < i; ++k) s = fma(-LU[i + N * k], x[k], s);
      x[i] = s;
    }
    SSB_UNROLL
    for (int i = N - 1; i >= 0; --i) {
      double s = x[i];
      SSB_UNROLL
      for (int k = i + 1; k < N; ++k) s = fma(-LU[i + N * k], x[k], s);
      x[i] = s * LU[i + N * i];
    }
  }
}

// Solve X A = B in place (B: R x N): X = B U^-1 L^-1.
template <int R, int N>
SSB_HDI void sm_lu_solve_right(const double *LU, double *B) {
  // X U = B : forward over columns
  SSB_UNROLL
  for (int j = 0; j < N; ++j) {
    SSB_UNROLL
    for (int k = 0; k < j; ++k) {
      const double u = LU[k + N * j];
      SSB_UNROLL
      for (int i = 0; i < R; ++i) B[i + R * j] = fma(-B[i + R * k], u, B[i + R * j]);
    }
    const double inv = LU[j + N * j];
    SSB_UNROLL
    for (int i = 0; i < R; ++i) B[i + R * j] *= inv;
  }
  // Y L = X : backward over columns (unit diagonal)
  SSB_UNROLL
  for (int j = N - 2; j >= 0; --j) {
    SSB_UNROLL
    for (int k = j + 1; k < N; ++k) {
      const double l = LU[k + N * j];
      SSB_UNROLL
      for (int i = 0; i < R; ++i) B[i + R * j] = fma(-B[i + R * k], l, B[i + R * j]);
    }
  }
}

// Cholesky factor of a symmetric positive definite matrix (lower triangle of A
// read, L written in the lower triangle, reciprocal diagonal in `dinv`).
template <int N>
SSB_HDI void sm_cholesky(double *A, double *dinv) {
  SSB_UNROLL
  for (int j = 0; j < N; ++j) {
    double d = A[j + N * j];
    SSB_UNROLL
    for (int k = 0; k < j; ++k) d = fma(-A[j + N * k], A[j + N * k], d);
    const double l = sqrt(d);
    const double inv = 1.0 / l;
    A[j + N * j] = l;
    dinv[j] = inv;
    SSB_UNROLL
    for (int i = j + 1; i < N; ++i) {
      double s = A[i + N * j];
      SSB_UNROLL
      for (int k = 0; k < j; ++k) s = fma(-A[i + N * k], A[j + N * k], s);
      A[i + N * j] = s * inv;
    }
  }
}

// warp-wide "everybody converged" vote (plain value on the host).  Only used to leave a
// loop early; results never depend on it.
SSB_HDI bool all_lanes(bool pred) {
#if defined(__CUDA_ARCH__)
  return __all_sync(__activemask(), pred);
#else
  return pred;
#endif
}

SSB_HDI double rsqrt_pos(double x) {
#if defined(__CUDA_ARCH__)
  return rsqrt(x);
#else
  return 1.0 / sqrt(x);
#endif
}

// Cyclic Jacobi eigen-decomposition of a symmetric matrix (full storage):
// on return A holds the eigenvalues on its diagonal and U the orthonormal
// eigenvectors (columns).  The rotation parameters come from two reciprocal
// square roots (no division):  with alpha = (aqq-app)/2, beta = apq,
// h = sqrt(alpha^2+beta^2):  cos^2 = (1 + |alpha|/h)/2,
// sin = sign(alpha) beta / (2 h cos).  A problem is converged when its
// off-diagonal mass is below eps^2 of the diagonal mass; from then on it only
// applies identity rotations, so its result does not depend on how many more
// sweeps its warp (or block) neighbours need - the sweeps stop when all of them
// have converged or after `max_sweeps`.  The pair loops are unrolled so all
// indices stay static.
template <int N>
SSB_HDI void sm_jacobi(double *A, double *U, double *eval, int max_sweeps) {
  SSB_UNROLL
  for (int j = 0; j < N; ++j) {
    SSB_UNROLL
    for (int i = 0; i < N; ++i) U[i + N * j] = (i == j) ? 1.0 : 0.0;
  }
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    double off = 0.0, diag = 0.0;
    SSB_UNROLL
    for (int p = 0; p < N; ++p) {
      diag = fma(A[p + N * p], A[p + N * p], diag);
      SSB_UNROLL
      for (int q = p + 1; q < N; ++q) off = fma(A[p + N * q], A[p + N * q], off);
    }
    const bool converged = off <= 1.0e-33 * diag;
    if (all_lanes(converged)) break;
    SSB_UNROLL
    for (int p = 0; p < N - 1; ++p) {
      SSB_UNROLL
      for (int q = p + 1; q < N; ++q) {
        const double beta = A[p + N * q];
        const double app = A[p + N * p], aqq = A[q + N * q];
        const double alpha = 0.5 * (aqq - app);
        const double h2 = fma(alpha, alpha, beta * beta);
        // negligible pivot (or an exactly zero 2x2 block): identity rotation
        const bool skip = converged || !(beta * beta > 1.0e-40 * h2);
        const double rh = rsqrt_pos(skip ? 1.0 : h2);
        const double x = fma(0.5 * fabs(alpha), rh, 0.5);
        const double rc = rsqrt_pos(x);
        const double c = skip ? 1.0 : x * rc;
        const double s = skip ? 0.0 : (alpha < 0.0 ? -0.5 : 0.5) * beta * rh * rc;
        const double t = s * rc;  // tan = sin / cos
        A[p + N * p] = fma(-t, beta, app);
        A[q + N * q] = fma(t, beta, aqq);
        A[p + N * q] = 0.0;
        A[q + N * p] = 0.0;
        SSB_UNROLL
        for (int k = 0; k < N; ++k) {
          if (k != p && k != q) {
            const double akp = A[k + N * p], akq = A[k + N * q];
            const double nkp = fma(c, akp, -(s * akq));
            const double nkq = fma(s, akp, c * akq);
            A[k + N * p] = nkp;
            A[p + N * k] = nkp;
            A[k + N * q] = nkq;
            A[q + N * k] = nkq;
          }
        }
        SSB_UNROLL
        for (int k = 0; k < N; ++k) {
          const double ukp = U[k + N * p], ukq = U[k + N * q];
          U[k + N * p] = fma(c, ukp, -(s * ukq));
          U[k + N * q] = fma(s, ukp, c * ukq);
        }
      }
    }
  }
  SSB_UNROLL
  for (int i = 0; i < N; ++i) eval[i] = A[i + N * i];
}

}  // namespace ssb
