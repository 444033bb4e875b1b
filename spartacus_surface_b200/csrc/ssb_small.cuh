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


namespace ssb {

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

}  // namespace ssb
