// ssb_math.cuh - batched tiny dense algebra for one problem per thread.
//
// Replaces the live subset of radtool/radtool_matrix.F90 (mat_x_mat :248,
// solve_vec_2/3 :779/:827, solve_mat_2/3 :801/:870, lu_factorization :982,
// lu_substitution :1024, lu_invert :1057, solve_rect_mat :1119, invert :1203,
// the "expanded" overlap products :505-649) and
// radtool/radtool_eigen_decomposition.F90:51-828 for the generic kernels.
//
// Design: every routine works on ONE problem held in the calling thread's
// registers / local memory (column-major, runtime order <= compile-time
// capacity), because the batch dimension of the B200 path is the CUDA thread
// index, not a leading array dimension.  Arithmetic order follows the
// reference (LU without pivoting, Cramer for order 2, explicit LU for order 3)
// so that results agree with it to rounding level; no pivoting is added.
// The file is plain C++ so that tests can compile the same code for the host.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define SSB_UNROLL _Pragma("unroll")
#define SSB_ROLLED _Pragma("unroll 1")  // keep a loop rolled: the code must fit the instruction cache
#else
#define SSB_UNROLL
#define SSB_ROLLED
#endif

#if defined(__CUDACC__)
#define SSB_HD __host__ __device__
#define SSB_HDI __host__ __device__ __forceinline__
#else
#define SSB_HD
#define SSB_HDI inline
#endif

namespace ssb {

#define SSB_EPS 2.220446049250313e-16
#define SSB_PI 3.14159265358979323846

SSB_HDI double dmax(double a, double b) { return a > b ? a : b; }
SSB_HDI double dmin(double a, double b) { return a < b ? a : b; }
SSB_HDI int imin(int a, int b) { return a < b ? a : b; }
SSB_HDI double fsign(double a, double b) { return signbit(b) ? -fabs(a) : fabs(a); }

// C(r x c) = A(r x k) * B(k x c); all column-major and compact.
SSB_HDI void mat_mul(int r, int k, int c, const double *A, const double *B, double *C) {
  for (int j = 0; j < c; ++j) {
    for (int i = 0; i < r; ++i) C[i + r * j] = 0.0;
    for (int l = 0; l < k; ++l) {
      const double b = B[l + k * j];
      for (int i = 0; i < r; ++i) C[i + r * j] = C[i + r * j] + A[i + r * l] * b;
    }
  }
}

// y(r) = A(r x c) * x(c)
SSB_HDI void mat_vec(int r, int c, const double *A, const double *x, double *y) {
  for (int i = 0; i < r; ++i) {
    double s = 0.0;
    for (int j = 0; j < c; ++j) s = s + A[i + r * j] * x[j];
    y[i] = s;
  }
}

// In-place LU factorisation without pivoting (column sweep, unit lower factor).
SSB_HDI void lu_factor(int n, double *A) {
  for (int j = 0; j < n; ++j) {
    for (int i = 0; i < j; ++i) {
      double s = A[i + n * j];
      for (int k = 0; k < i; ++k) s = s - A[i + n * k] * A[k + n * j];
      A[i + n * j] = s;
    }
    for (int i = j; i < n; ++i) {
      double s = A[i + n * j];
      for (int k = 0; k < j; ++k) s = s - A[i + n * k] * A[k + n * j];
      A[i + n * j] = s;
    }
    if (j != n - 1) {
      const double inv = 1.0 / A[j + n * j];
      for (int i = j + 1; i < n; ++i) A[i + n * j] = A[i + n * j] * inv;
    }
  }
}

// Forward/back substitution of one right-hand side, in place (stride 1).
SSB_HDI void lu_subst(int n, const double *LU, double *x) {
  for (int i = 1; i < n; ++i)
    for (int k = 0; k < i; ++k) x[i] = x[i] - x[k] * LU[i + n * k];
  for (int i = n - 1; i >= 0; --i) {
    for (int k = i + 1; k < n; ++k) x[i] = x[i] - x[k] * LU[i + n * k];
    x[i] = x[i] / LU[i + n * i];
  }
}

// X = inverse from an LU factorisation (identity right-hand sides).
SSB_HDI void lu_inverse(int n, const double *LU, double *X) {
  for (int j = 0; j < n; ++j) {
    double *x = X + n * j;
    for (int i = 0; i < n; ++i) x[i] = 0.0;
    x[j] = 1.0;
    lu_subst(n, LU, x);
  }
}

// General inverse: A is destroyed (holds its LU afterwards).
SSB_HDI void invert(int n, double *A, double *X) {
  lu_factor(n, A);
  lu_inverse(n, A, X);
}

// X(n x c) = A^-1 B for a general right-hand side block; A destroyed.
SSB_HDI void solve_rect(int n, int c, double *A, const double *B, double *X) {
  lu_factor(n, A);
  for (int j = 0; j < c; ++j) {
    for (int i = 0; i < n; ++i) X[i + n * j] = B[i + n * j];
    lu_subst(n, A, X + n * j);
  }
}

// Order-3 elimination shared by the vector and matrix solves.
struct Lu3 {
  double L21, L31, L32, U22, U23, U33;
};
SSB_HDI Lu3 lu3(const double *A) {
  Lu3 f;
  f.L21 = A[1] / A[0];
  f.L31 = A[2] / A[0];
  f.U22 = A[4] - f.L21 * A[3];
  f.U23 = A[7] - f.L21 * A[6];
  f.L32 = (A[5] - f.L31 * A[3]) / f.U22;
  f.U33 = A[8] - f.L31 * A[6] - f.L32 * f.U23;
  return f;
}
SSB_HDI void lu3_apply(const double *A, const Lu3 &f, const double *b, double *x) {
  const double y2 = b[1] - f.L21 * b[0];
  const double y3 = b[2] - f.L31 * b[0] - f.L32 * y2;
  x[2] = y3 / f.U33;
  x[1] = (y2 - f.U23 * x[2]) / f.U22;
  x[0] = (b[0] - A[3] * x[1] - A[6] * x[2]) / A[0];
}

// x = A^-1 b with the reference's order-specific forms; `work` (n*n) receives
// the LU factors for n > 3.  A is left untouched.
SSB_HDI void solve_vec(int n, const double *A, const double *b, double *x, double *work) {
  if (n == 2) {
    const double inv_det = 1.0 / (A[0] * A[3] - A[2] * A[1]);
    const double b0 = b[0], b1 = b[1];
    x[0] = inv_det * (A[3] * b0 - A[2] * b1);
    x[1] = inv_det * (A[0] * b1 - A[1] * b0);
  } else if (n == 3) {
    const Lu3 f = lu3(A);
    double t[3] = {b[0], b[1], b[2]};
    lu3_apply(A, f, t, x);
  } else {
    for (int i = 0; i < n * n; ++i) work[i] = A[i];
    lu_factor(n, work);
    for (int i = 0; i < n; ++i) x[i] = b[i];
    lu_subst(n, work, x);
  }
}

// X = A^-1 B (square B); `work` receives LU factors for n > 3 (or n == 1).
SSB_HDI void solve_mat(int n, const double *A, const double *B, double *X, double *work) {
  if (n == 2) {
    const double inv_det = 1.0 / (A[0] * A[3] - A[2] * A[1]);
    for (int j = 0; j < 2; ++j) {
      const double b0 = B[2 * j], b1 = B[1 + 2 * j];
      X[2 * j] = inv_det * (A[3] * b0 - A[2] * b1);
      X[1 + 2 * j] = inv_det * (A[0] * b1 - A[1] * b0);
    }
  } else if (n == 3) {
    const Lu3 f = lu3(A);
    for (int j = 0; j < 3; ++j) {
      double t[3] = {B[3 * j], B[1 + 3 * j], B[2 + 3 * j]};
      lu3_apply(A, f, t, X + 3 * j);
    }
  } else {
    for (int i = 0; i < n * n; ++i) work[i] = A[i];
    solve_rect(n, n, work, B, X);
  }
}

// (The reference-order QR eigen-solver lives in tests/hostcheck/asymtx_qr.hpp: it is compiled into the HOST
// check only, where it pins the generic bodies bit for bit against the oracle.  Every device path uses the
// symmetrised Jacobi solvers of ssb_radtool.cuh / ssb_layer_math.cuh.)

}  // namespace ssb
