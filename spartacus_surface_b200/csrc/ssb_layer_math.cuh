// ssb_layer_math.cuh - layer transfer matrices for compile-time orders (NR
// solved regions x NS streams), one thread per layer problem, organised so
// that the problem stays on chip.
//
// Computes the same outputs as calc_matrices_sw_eig / calc_matrices_lw_eig
// (radtool/radtool_calc_matrices_sw_eig.F90:30-386, _lw_eig.F90:32-230,
// radtool/radtool_schur.F90:32-53) through an equivalent formulation without
// data-dependent control flow:
//
//  * Gamma1, Gamma2 are "N-symmetric": Gamma * diag(1/N) is symmetric with
//    N_i = 1 / (weight_js * mu_js * frac_r) (detailed balance of the exchange
//    and scattering terms).  With D = G1-G2, S = G1+G2 and -N D = L L^T
//    (Cholesky) the eigenproblem of P = D S becomes the SYMMETRIC problem
//    Y = L^T K L, K = -S/N, solved by cyclic Jacobi; eigenvectors
//    V = N^-1 L U and M = -(D^-1 V) Lambda^1/2 = L^-T U Lambda^1/2 come
//    without another factorisation.
//  * D couples regions but not streams, so in the (stream, region) index
//    L and L^-1 are block sparse: only same-stream entries exist.  All index
//    tests below are on unrolled (compile-time) indices and fold away.
//  * The 2n x 2n two-point problem splits into sum and difference problems:
//    X+ = R+T = B+ A+^-1, X- = R-T = B- A-^-1 with
//    A(s) = V(1+s e) + M(1-s e), B(s) = V(1+s e) - M(1-s e), e = exp(-lambda dz).
//  * The direct-diffuse coupling reduces to products with X+ and X-, the
//    per-mode inversions to (eps^2 - P)^-1 through the eigenvectors, and the
//    Schur inverse to S^-1.
//
// Register discipline (what bounds these kernels on B200): a thread never
// holds more than two order-n matrices.  The Gamma matrices are never stored -
// `LayerCoef` evaluates their entries from ~30 scalars; results are written to
// the layer scratch as soon as they exist (X+ is parked in the R slot between
// the two stages); what must survive several phases (U, L, L^-1, the direct
// modes) lives in a per-thread slice of shared memory (`StateMem`).
#pragma once
#include "ssb_small.cuh"
#include "ssb_solver.cuh"

namespace ssb {

// per-thread slice: element e at p[e * stride] (shared memory with stride =
// blockDim.x on the device: conflict-free; a plain local array on the host)
struct StateMem {
  double *p;
  int stride;
  SSB_HDI double &operator()(int e) const { return p[(size_t)e * stride]; }
};

// The same pointer, opaque to the optimiser (an offset of zero it cannot see through, which
// keeps the address space known): values parked in the scratch or in the stack slice are
// re-read through it, so that the compiler really drops them from registers instead of
// forwarding the stores to the loads.
template <class T>
SSB_HDI T *opaque_ptr(T *p) {
#if defined(__CUDA_ARCH__)
  size_t zero = 0;
  asm volatile("" : "+l"(zero));
  return p + zero;
#else
  return p;
#endif
}
SSB_HDI StateMem opaque(const StateMem &s) { return StateMem{opaque_ptr(s.p), s.stride}; }

SSB_HD constexpr int kJacobiSweeps(int n) { return n <= 2 ? 4 : 12; }
constexpr double kJacobiTol2 = 1.0e-31;  // (3e-16)^2

// slot of L(i,j), i >= j in the same stream, in the packed block-sparse storage
template <int NR, int NS>
SSB_HD constexpr int lslot(int i, int j) {
  return (i % NS) * (NR * (NR + 1) / 2) + (i / NS) * ((i / NS) + 1) / 2 + (j / NS);
}
template <int NR, int NS>
struct LayerStack {  // shared-memory slice layout
  static constexpr int N = NR * NS, NL = NS * NR * (NR + 1) / 2;
  static constexpr int oU = 0, oL = N * N, oLi = oL + NL, oX = oLi + NL;
  // shortwave: direct modes U0 (NR x NR), eps, sqrt(frac), 1/sqrt(frac); longwave: y (N)
  static constexpr int sw_doubles = oX + NR * NR + 3 * NR, lw_doubles = oX + N;
};

// Entries of the Gamma matrices of one layer from scalars (regions r = 0..NR-1 are the
// SOLVED regions; index i = js + r * NS).
template <int NR, int NS>
struct LayerCoef {
  double dx[NR * NR];  // exchange rate into region rt from region rf, [rt + NR * rf], rt != rf
  double loss[NR];     // total exchange rate out of region rf (to every region of the layer)
  double ext[NR], es[NR], fw[NR], frac[NR], rfrac[NR];
  double wall_ext, wall_factor;
  double tan0, sin0, rcos;  // shortwave direct beam: tan, sin and 1/cos of the solar zenith angle
  double rmu[NS], wmu[NS], rwmu[NS];
  const LgTable *lg;
  SSB_HDI void set_streams(const LgTable *t) {
    lg = t;
    SSB_UNROLL
    for (int js = 0; js < NS; ++js) {
      rmu[js] = 1.0 / t->mu[js];
      wmu[js] = t->weight[js] * t->mu[js];
      rwmu[js] = 1.0 / wmu[js];
    }
  }
  // D = Gamma1 - Gamma2: exchange between regions and extinction, per stream
  SSB_HDI double D(int i, int j) const {
    const int js = i % NS, rt = i / NS, rf = j / NS;
    if (js != j % NS) return 0.0;
    if (rt == rf) return -fma(lg->tan_ang[js], fma(fw[rf], wall_ext, loss[rf]), ext[rf] * rmu[js]);
    return lg->tan_ang[js] * dx[rt + NR * rf];
  }
  // 2 Gamma2: scattering (jt, r) <- (js, r), per region
  SSB_HDI double G2x2(int i, int j) const {
    const int jt = i % NS, js = j % NS, r = j / NS;
    if (i / NS != r) return 0.0;
    return fma(lg->weight[jt] * es[r], rmu[js], lg->vweight[jt] * lg->tan_ang[js] * (fw[r] * wall_factor));
  }
  SSB_HDI double S(int i, int j) const { return D(i, j) + G2x2(i, j); }
  SSB_HDI double ninv(int i) const { return wmu[i % NS] * frac[i / NS]; }
  SSB_HDI double nsc(int i) const { return rwmu[i % NS] * rfrac[i / NS]; }
  // Gamma3 has one entry per row: (js, r) <- direct beam in region r
  SSB_HDI double g3(int i) const {
    const int js = i % NS, r = i / NS;
    return 0.5 * fma(lg->weight[js], es[r], lg->vweight[js] * sin0 * (fw[r] * wall_factor));
  }
  // Gamma0 (direct beam, NR x NR)
  SSB_HDI double g0(int i, int j) const {
    if (i == j) return -fma(tan0, fma(fw[i], wall_ext, loss[i]), ext[i] * rcos);
    return tan0 * dx[i + NR * j];
  }
};

// Cyclic Jacobi on a symmetric matrix held as its lower triangle (Y[i + N*j], i >= j);
// eigenvalues return on the diagonal, eigenvectors in the columns of U.  (Jacobi, not
// tridiagonal QL: the spectrum spans ten orders of magnitude between clear air and dense
// vegetation and the small eigenvalues are needed to high RELATIVE accuracy; a round-robin
// pair ordering that exposes three independent rotations at a time was measured: no gain; at order 12
// a form with the pair loops rolled and Y, U addressed dynamically - the 66 unrolled rotations of a sweep
// overflow the instruction cache - was measured as well: 28 -> 45 ms for the 4-stream layer kernels.)
// The rotation
// parameters come from two reciprocal square roots (no division): with alpha = (aqq-app)/2,
// beta = apq, h = sqrt(alpha^2+beta^2): cos^2 = (1 + |alpha|/h)/2, sin = sign(alpha) beta /
// (2 h cos).  A problem is converged when its off-diagonal mass is below eps^2 of the
// diagonal mass; from then on it applies identity rotations only, so its result does not
// depend on how many more sweeps its warp neighbours need (the sweeps stop when all lanes
// have converged or after `max_sweeps`).
// Returns false when the scaled criterion is still violated after `max_sweeps` sweeps (the
// reference counts such problems in nerror / ierror, radtool_eigen_decomposition.F90:92-100,
// 498-510); the caller reports it through the status word.
template <int N>
SSB_HDI bool sm_jacobi_sym(double *Y, double *U, int max_sweeps) {
#define SSB_YS(a, b) Y[((a) >= (b)) ? ((a) + N * (b)) : ((b) + N * (a))]
  SSB_UNROLL
  for (int j = 0; j < N; ++j) {
    SSB_UNROLL
    for (int i = 0; i < N; ++i) U[i + N * j] = (i == j) ? 1.0 : 0.0;
  }
  bool done = false;
  for (int sweep = 0; sweep <= max_sweeps; ++sweep) {
    // scaled criterion |a_pq| <= tol sqrt(a_pp a_qq) for every pair: what gives the small
    // eigenvalues of a graded matrix their RELATIVE accuracy (a test against the total diagonal
    // mass stops while the entries that couple the small eigenvalues are still 1e-4 of them:
    // test/rami5 has layers whose regions differ by 1e6 in area and 1e6 in extinction)
    bool converged = true;
    SSB_UNROLL
    for (int p = 0; p < N; ++p) {
      SSB_UNROLL
      for (int q = p + 1; q < N; ++q) {
        // ... or the pair is numerically diagonal already: y_pq so small against the difference of the
        // diagonal entries that its rotation is the identity (it is skipped below) - on strongly graded
        // layers (y_pp ~ 1e-2, y_qq ~ 1e-13) the scaled test alone can stay violated for ever
        const double b2 = Y[q + N * p] * Y[q + N * p], al = 0.5 * (Y[q + N * q] - Y[p + N * p]);
        converged = converged && (b2 <= kJacobiTol2 * fabs(Y[p + N * p] * Y[q + N * q]) || !(b2 > 1.0e-40 * fma(al, al, b2)));
      }
    }
    done = converged;
    if (sweep == max_sweeps || all_lanes(converged)) break;  // (the last pass only tests)
    SSB_UNROLL
    for (int p = 0; p < N - 1; ++p) {
      SSB_UNROLL
      for (int q = p + 1; q < N; ++q) {
        const double beta = Y[q + N * p];
        const double app = Y[p + N * p], aqq = Y[q + N * q];
        const double alpha = 0.5 * (aqq - app);
        const double h2 = fma(alpha, alpha, beta * beta);
        const bool skip = converged || !(beta * beta > 1.0e-40 * h2);
        // (skipping, with a warp vote, the rotations that are the identity for every problem of the warp
        // was measured: the votes and branches cost the layer kernels 10 % - straight-line code wins)
        const double rh = rsqrt_pos(skip ? 1.0 : h2);
        const double x = fma(0.5 * fabs(alpha), rh, 0.5);
        const double rc = rsqrt_pos(x);
        const double c = skip ? 1.0 : x * rc;
        const double s = skip ? 0.0 : (alpha < 0.0 ? -0.5 : 0.5) * beta * rh * rc;
        const double t = s * rc;
        Y[p + N * p] = fma(-t, beta, app);
        Y[q + N * q] = fma(t, beta, aqq);
        Y[q + N * p] = 0.0;
        SSB_UNROLL
        for (int k = 0; k < N; ++k) {
          if (k != p && k != q) {
            const double akp = SSB_YS(k, p), akq = SSB_YS(k, q);
            SSB_YS(k, p) = fma(c, akp, -(s * akq));
            SSB_YS(k, q) = fma(s, akp, c * akq);
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
#undef SSB_YS
  return done;
}

// Eigen-system of P = D S: Cholesky of -N D, Y = L^T K L, Jacobi.  Leaves U, L, L^-1 in
// the stack slice and returns lambda = sqrt(eigenvalue), e = exp(-lambda dz) and the
// eigenvalues themselves.
template <int NR, int NS>
SSB_HDI bool layer_eigen(const LayerCoef<NR, NS> &c, double dz, const StateMem &st, double *lam2, double *lam,
                         double *e) {
  typedef LayerStack<NR, NS> Stk;
  constexpr int N = NR * NS;
  double U[N * N];
  bool converged;
  {
    double Y[N * N];  // lower triangle
    {
      double L[N * N], Ldinv[N];  // same-stream lower entries only
      SSB_UNROLL
      for (int j = 0; j < N; ++j) {
        double d = -c.nsc(j) * c.D(j, j);
        SSB_UNROLL
        for (int k = j % NS; k < j; k += NS) d = fma(-L[j + N * k], L[j + N * k], d);
        const double l = sqrt(d), inv = 1.0 / l;
        L[j + N * j] = l;
        Ldinv[j] = inv;
        SSB_UNROLL
        for (int i = j + NS; i < N; i += NS) {
          double s = -c.nsc(i) * c.D(i, j);
          SSB_UNROLL
          for (int k = j % NS; k < j; k += NS) s = fma(-L[i + N * k], L[j + N * k], s);
          L[i + N * j] = s * inv;
        }
      }
      // L^-1 (same sparsity)
      SSB_UNROLL
      for (int j = 0; j < N; ++j) {
        double Li[N];  // column j of L^-1, rows j, j+NS, ...
        Li[j] = Ldinv[j];
        st(Stk::oLi + lslot<NR, NS>(j, j)) = Li[j];
        SSB_UNROLL
        for (int i = j + NS; i < N; i += NS) {
          double s = 0.0;
          SSB_UNROLL
          for (int k = j; k < i; k += NS) s = fma(L[i + N * k], Li[k], s);
          Li[i] = -s * Ldinv[i];
          st(Stk::oLi + lslot<NR, NS>(i, j)) = Li[i];
        }
      }
      // K = -S / N (symmetric; lower triangle kept), Y = L^T K L column by column
      double K[N * N];
      SSB_UNROLL
      for (int k = 0; k < N; ++k) {
        SSB_UNROLL
        for (int i = k; i < N; ++i) K[i + N * k] = -c.S(i, k) * c.ninv(k);
      }
      SSB_UNROLL
      for (int j = 0; j < N; ++j) {
        double t1[N];  // (K L)(:, j)
        SSB_UNROLL
        for (int i = 0; i < N; ++i) {
          double s = 0.0;
          SSB_UNROLL
          for (int k = j; k < N; k += NS) s = fma((i >= k) ? K[i + N * k] : K[k + N * i], L[k + N * j], s);
          t1[i] = s;
        }
        SSB_UNROLL
        for (int i = j; i < N; ++i) {
          double s = 0.0;
          SSB_UNROLL
          for (int k = i; k < N; k += NS) s = fma(L[k + N * i], t1[k], s);
          Y[i + N * j] = s;
        }
      }
      SSB_UNROLL
      for (int j = 0; j < N; ++j) {
        SSB_UNROLL
        for (int i = j; i < N; i += NS) st(Stk::oL + lslot<NR, NS>(i, j)) = L[i + N * j];
      }
    }
    converged = sm_jacobi_sym<N>(Y, U, kJacobiSweeps(N));
    SSB_UNROLL
    for (int k = 0; k < N; ++k) lam2[k] = Y[k + N * k];
  }
  SSB_UNROLL
  for (int i = 0; i < N * N; ++i) st(Stk::oU + i) = U[i];
  SSB_UNROLL
  for (int k = 0; k < N; ++k) {
    lam[k] = sqrt(dmax(0.0, lam2[k]));
    e[k] = exp(-lam[k] * dz);
  }
  return converged;
}

// rows i of V = N^-1 L U and M = L^-T U diag(lam) from the stack slice
template <int NR, int NS>
SSB_HDI void layer_vm_row(const LayerCoef<NR, NS> &c, const StateMem &st, int i, const double *lam, double *v,
                          double *m) {
  typedef LayerStack<NR, NS> Stk;
  constexpr int N = NR * NS;
  SSB_UNROLL
  for (int k = 0; k < N; ++k) v[k] = m[k] = 0.0;
  SSB_UNROLL
  for (int j = i % NS; j < N; j += NS) {
    // j <= i contributes to V through L(i,j), j >= i to M through L^-1(j,i)
    const double lv = (j <= i) ? st(Stk::oL + lslot<NR, NS>(i, j)) : 0.0;
    const double lm = (j >= i) ? st(Stk::oLi + lslot<NR, NS>(j, i)) : 0.0;
    SSB_UNROLL
    for (int k = 0; k < N; ++k) {
      const double u = st(Stk::oU + j + N * k);
      if (j <= i) v[k] = fma(lv, u, v[k]);
      if (j >= i) m[k] = fma(lm, u, m[k]);
    }
  }
  const double ni = c.ninv(i);
  SSB_UNROLL
  for (int k = 0; k < N; ++k) {
    v[k] *= ni;
    m[k] *= lam[k];
  }
}

// addressing of the layer scratch of one problem: element e at P[e * kScratchTile]
#define SSB_OUT(P, e) (P)[(size_t)(e) * kScratchTile]
// final value of an element: not read again in this kernel, consumed by the sweeps hundreds of
// megabytes later - a streaming (evict-first) store keeps it from displacing the parked rows
// (column-resident kernels: the tile is private and re-read within microseconds - plain stores)
#if defined(__CUDA_ARCH__) && !defined(SSB_FUSED_TU)
#define SSB_FINAL(P, e, v) __stcs((P) + (size_t)(e) * kScratchTile, (v))
#else
#define SSB_FINAL(P, e, v) ((P)[(size_t)(e) * kScratchTile] = (v))
#endif

// One stage of the two-point solve, X = B A^-1 with A = V(1+s e) + M(1-s e),
// B = V(1+s e) - M(1-s e): builds both row by row, parks the rows of B in the scratch
// (matrix slot eB of the full-size layout, NREG regions, block offset i0) and returns A
// LU-factored; the caller then solves x_i A = b_i one row at a time.  `z` non-null:
// also parks the rows of V diag(z) in slot eC (longwave int_flux).
template <int NREG, int NS, int NR, int R0>
SSB_HDI void layer_stage_factor(const LayerCoef<NR, NS> &c, const StateMem &st, double sigma, const double *lam,
                                const double *e, double *A, double *P, int eB, const double *z, int eC) {
  constexpr int N = NR * NS, n = NREG * NS, i0 = R0 * NS;
  SSB_UNROLL
  for (int i = 0; i < N; ++i) {
    double v[N], m[N];
    layer_vm_row<NR, NS>(c, st, i, lam, v, m);
    SSB_UNROLL
    for (int k = 0; k < N; ++k) {
      const double vv = v[k] * fma(sigma, e[k], 1.0), mm = m[k] * fma(-sigma, e[k], 1.0);
      A[i + N * k] = vv + mm;
      SSB_OUT(P, eB + (i + i0) + n * (k + i0)) = vv - mm;
      if (z != nullptr) SSB_OUT(P, eC + (i + i0) + n * (k + i0)) = v[k] * z[k];
    }
  }
  sm_lu<N>(A);
}
// row i of a parked matrix (rows beyond the block read row N-1 again: the row loops fetch one
// row ahead so that the L2 latency of the parked rows hides behind the solve of the previous one)
template <int NREG, int NS, int NR, int R0>
SSB_HDI void layer_load_row(const double *Pr, int eB, int i, double *x) {
  constexpr int N = NR * NS, n = NREG * NS, i0 = R0 * NS;
  const int ii = (i < N ? i : N - 1) + i0;
  SSB_UNROLL
  for (int k = 0; k < N; ++k) x[k] = SSB_OUT(Pr, eB + ii + n * (k + i0));
}

// Shortwave layer of the solved block (NR regions starting at region R0 of NREG): writes
// R, T, int_diff, S_up, S_dn, int_dir_diff, E, int_dir to the layer scratch at P, at their
// positions in the full-size matrices (radsurf_urban_sw.F90:512-583); the zeros outside the
// block are not written - the sweeps know the segment of the layer and do not load them.  Returns false when
// R or T is not finite.
template <int NREG, int NS, int NR, int R0>
SSB_HDI bool layer_sw_solve(const LayerCoef<NR, NS> &c, double dz, double *P, const StateMem &st_in) {
  typedef LayerStack<NR, NS> Stk;
  constexpr int N = NR * NS, D = NR, n = NREG * NS, d = NREG, i0 = R0 * NS, r0 = R0;
  constexpr int eR = 0, eT = n * n, eIdiff = 2 * n * n, eSup = 3 * n * n, eSdn = eSup + n * d, eIdd = eSdn + n * d,
                eE = eIdd + n * d, eIdir = eE + d * d;
  constexpr int oU0 = Stk::oX, oEps = oU0 + D * D, oSq = oEps + D, oRsq = oSq + D;
  // ---- direct beam: g0 = B0 diag(1/frac) with B0 symmetric -> symmetric Jacobi ----------
  double g0inv[D * D];
  bool jac0;
  {
    double sq[D], rsq[D], Y0[D * D], U0[D * D];
    SSB_UNROLL
    for (int r = 0; r < D; ++r) {
      sq[r] = sqrt(c.frac[r]);
      rsq[r] = 1.0 / sq[r];
    }
    SSB_UNROLL
    for (int j = 0; j < D; ++j) {
      SSB_UNROLL
      for (int i = j; i < D; ++i) Y0[i + D * j] = c.g0(i, j) * sq[j] * rsq[i];
    }
    jac0 = sm_jacobi_sym<D>(Y0, U0, kJacobiSweeps(D));
    double e0[D], reps[D];
    SSB_UNROLL
    for (int k = 0; k < D; ++k) {
      const double eps = Y0[k + D * k];
      e0[k] = exp(eps * dz);
      reps[k] = 1.0 / eps;
      st_in(oEps + k) = eps;
      st_in(oSq + k) = sq[k];
      st_in(oRsq + k) = rsq[k];
      SSB_UNROLL
      for (int i = 0; i < D; ++i) st_in(oU0 + i + D * k) = U0[i + D * k];
    }
    // E = G0 diag(e0) G0^-1, g0^-1 = G0 diag(1/eps) G0^-1 with G0 = sq U0, G0^-1 = U0^T / sq
    SSB_UNROLL
    for (int j = 0; j < D; ++j) {
      SSB_UNROLL
      for (int i = 0; i < D; ++i) {
        double se = 0.0, si = 0.0;
        SSB_UNROLL
        for (int k = 0; k < D; ++k) {
          const double gk = (sq[i] * U0[i + D * k]) * (U0[j + D * k] * rsq[j]);
          se = fma(gk, e0[k], se);
          si = fma(gk, reps[k], si);
        }
        SSB_FINAL(P, eE + (i + r0) + d * (j + r0), se);
        SSB_FINAL(P, eIdir + (i + r0) + d * (j + r0), -si);
        g0inv[i + D * j] = si;
      }
    }
  }
  // ---- integrated-flux matrices: int_diff = -S^-1, int_dir_diff = 2 S^-1 G3 g0^-1 --------
  {
    double LUs[N * N], Idd[N * D];
    SSB_UNROLL
    for (int j = 0; j < N; ++j) {
      SSB_UNROLL
      for (int i = 0; i < N; ++i) LUs[i + N * j] = c.S(i, j);
    }
    sm_lu<N>(LUs);
    SSB_UNROLL
    for (int i = 0; i < N * D; ++i) Idd[i] = 0.0;
    SSB_UNROLL
    for (int j = 0; j < N; ++j) {
      double x[N];  // column j of -S^-1
      SSB_UNROLL
      for (int i = 0; i < N; ++i) x[i] = (i == j) ? -1.0 : 0.0;
      sm_lu_solve_left<N, 1>(LUs, x);
      SSB_UNROLL
      for (int i = 0; i < N; ++i) SSB_FINAL(P, eIdiff + (i + i0) + n * (j + i0), x[i]);
      const double gj = -2.0 * c.g3(j);
      SSB_UNROLL
      for (int jd = 0; jd < D; ++jd) {
        const double w = gj * g0inv[j / NS + D * jd];
        SSB_UNROLL
        for (int i = 0; i < N; ++i) Idd[i + N * jd] = fma(x[i], w, Idd[i + N * jd]);
      }
    }
    SSB_UNROLL
    for (int jd = 0; jd < D; ++jd) {
      SSB_UNROLL
      for (int i = 0; i < N; ++i) SSB_FINAL(P, eIdd + (i + i0) + n * (jd + r0), Idd[i + N * jd]);
    }
  }
  // ---- diffuse eigen-system and the two-point solve ---------------------------------------
  double lam2[N], lam[N], e[N];
  bool bad = !layer_eigen<NR, NS>(c, dz, st_in, lam2, lam, e) || !jac0;
  const StateMem st = opaque(st_in);
  {
    // sum (sigma = +1) then difference (-1) problem; rows of B pass through the T slot, X+ is
    // parked in the R slot.  Rolled loops: the same code serves both stages and all rows.
    double A[N * N];
    SSB_ROLLED
    for (int sg = 0; sg < 2; ++sg) {
      layer_stage_factor<NREG, NS, NR, R0>(c, st, sg ? -1.0 : 1.0, lam, e, A, P, eT, nullptr, 0);
      const double *Pr = opaque_ptr(P);
      double xn[N], xpn[N];
      layer_load_row<NREG, NS, NR, R0>(Pr, eT, 0, xn);
      layer_load_row<NREG, NS, NR, R0>(Pr, eR, 0, xpn);  // X+ (only meaningful in the second stage)
      SSB_ROLLED
      for (int i = 0; i < N; ++i) {
        double x[N], xpr[N];
        SSB_UNROLL
        for (int k = 0; k < N; ++k) {
          x[k] = xn[k];
          xpr[k] = xpn[k];
        }
        layer_load_row<NREG, NS, NR, R0>(Pr, eT, i + 1, xn);
        layer_load_row<NREG, NS, NR, R0>(Pr, eR, i + 1, xpn);
        sm_lu_solve_right<1, N>(A, x);
        if (sg == 0) {
          SSB_UNROLL
          for (int k = 0; k < N; ++k) SSB_OUT(P, eR + (i + i0) + n * (k + i0)) = x[k];
        } else {
          SSB_UNROLL
          for (int k = 0; k < N; ++k) {
            const double xp = xpr[k];
            const double r = 0.5 * (xp + x[k]), t = 0.5 * (xp - x[k]);
            bad = bad || !(fabs(r) < 1.0e300) || !(fabs(t) < 1.0e300);
            SSB_OUT(P, eR + (i + i0) + n * (k + i0)) = r;
            SSB_OUT(P, eT + (i + i0) + n * (k + i0)) = t;
          }
        }
      }
    }
  }
  // ---- particular solutions per direct eigen-mode:
  //   a = g3p + g4p = 2 V (eps^2 - Lambda)^-1 V^-1 D c ,  b = g3p - g4p = -(S a + 2c)/eps,
  // source terms S_up +- S_dn = +-(R+-T)(r1+-r2) + (G3p +- G4p e0) G0^-1
  // (rolled over the modes; G3p, G4p are parked in the S_up / S_dn slots)
  SSB_ROLLED
  for (int jd = 0; jd < D; ++jd) {
    const double eps = st(oEps + jd);
    double a[N], cc[N];
    {
      double w[N], t[N];
      SSB_UNROLL
      for (int i = 0; i < N; ++i) cc[i] = c.g3(i) * (st(oSq + i / NS) * st(oU0 + i / NS + D * jd));
      // w = L^-1 N D c
      SSB_UNROLL
      for (int i = 0; i < N; ++i) {
        double s = 0.0;
        SSB_UNROLL
        for (int k = i % NS; k < N; k += NS) s = fma(c.D(i, k), cc[k], s);
        t[i] = s * c.nsc(i);
      }
      SSB_UNROLL
      for (int i = 0; i < N; ++i) {
        double s = 0.0;
        SSB_UNROLL
        for (int k = i % NS; k <= i; k += NS) s = fma(st(Stk::oLi + lslot<NR, NS>(i, k)), t[k], s);
        w[i] = s;
      }
      // t = 2 (eps^2 - lambda^2)^-1 U^T w ; then a = N^-1 L U t
      SSB_UNROLL
      for (int k = 0; k < N; ++k) {
        double s = 0.0;
        SSB_UNROLL
        for (int i = 0; i < N; ++i) s = fma(st(Stk::oU + i + N * k), w[i], s);
        // eps^2 = lambda_k^2 (a direct-beam mode decaying exactly like a diffuse mode) is a removable
        // singularity of S_up / S_dn, but a pole of this particular solution (the reference's
        // Gamma1 - Q Gamma2 - eps I is singular there too, radtool_calc_matrices_sw_eig.F90:232-253):
        // keep the two apart by 1e-8 relative, which perturbs the sources by that order instead of
        // returning Inf / NaN
        const double e2 = eps * eps;
        double den = e2 - lam2[k];
        const double dmin = 1.0e-8 * (e2 + fabs(lam2[k]));
        if (fabs(den) < dmin) den = (den < 0.0) ? -dmin : dmin;
        t[k] = 2.0 * s / den;
      }
      SSB_UNROLL
      for (int i = 0; i < N; ++i) {
        double s = 0.0;
        SSB_UNROLL
        for (int k = 0; k < N; ++k) s = fma(st(Stk::oU + i + N * k), t[k], s);
        w[i] = s;
      }
      SSB_UNROLL
      for (int i = 0; i < N; ++i) {
        double s = 0.0;
        SSB_UNROLL
        for (int k = i % NS; k <= i; k += NS) s = fma(st(Stk::oL + lslot<NR, NS>(i, k)), w[k], s);
        a[i] = s * c.ninv(i);
      }
    }
    const double mreps = -1.0 / eps;
    SSB_UNROLL
    for (int i = 0; i < N; ++i) {
      double s = 2.0 * cc[i];
      SSB_UNROLL
      for (int k = 0; k < N; ++k) {
        if (k % NS == i % NS || k / NS == i / NS) s = fma(c.S(i, k), a[k], s);
      }
      const double b = s * mreps;
      SSB_OUT(P, eSup + (i + i0) + n * (jd + r0)) = 0.5 * (a[i] + b);
      SSB_OUT(P, eSdn + (i + i0) + n * (jd + r0)) = 0.5 * (a[i] - b);
    }
  }
  // S_up + S_dn = (R+T) rp + qp ; S_up - S_dn = (T-R) rm + qm, row by row from the scratch, with
  //   rp = -(G3p e0 + G4p) G0^-1, rm = -(G3p e0 - G4p) G0^-1, qp = (G3p + G4p e0) G0^-1, qm = (G3p - G4p e0) G0^-1
  {
    double G0i[D * D], e0[D], rp[N * D], rm[N * D], g3p[N * D], g4p[N * D];
    const double *Ps = opaque_ptr(P);
    SSB_UNROLL
    for (int jd = 0; jd < D; ++jd) {
      SSB_UNROLL
      for (int i = 0; i < N; ++i) {
        g3p[i + N * jd] = SSB_OUT(Ps, eSup + (i + i0) + n * (jd + r0));
        g4p[i + N * jd] = SSB_OUT(Ps, eSdn + (i + i0) + n * (jd + r0));
      }
      e0[jd] = exp(st(oEps + jd) * dz);
      SSB_UNROLL
      for (int j = 0; j < D; ++j) G0i[jd + D * j] = st(oU0 + j + D * jd) * st(oRsq + j);
    }
    SSB_UNROLL
    for (int j = 0; j < D; ++j) {
      SSB_UNROLL
      for (int i = 0; i < N; ++i) {
        double sp = 0.0, sm = 0.0;
        SSB_UNROLL
        for (int jd = 0; jd < D; ++jd) {
          const double g3e = g3p[i + N * jd] * e0[jd];
          sp = fma(-(g3e + g4p[i + N * jd]), G0i[jd + D * j], sp);
          sm = fma(-(g3e - g4p[i + N * jd]), G0i[jd + D * j], sm);
        }
        rp[i + N * j] = sp;
        rm[i + N * j] = sm;
      }
    }
    // rows are independent: rolled loop, G3p / G4p of the row re-read from their parking slots
    SSB_ROLLED
    for (int i = 0; i < N; ++i) {
      double xp[N], xm[N];
      SSB_UNROLL
      for (int k = 0; k < N; ++k) {
        const double r = SSB_OUT(Ps, eR + (i + i0) + n * (k + i0)), t = SSB_OUT(Ps, eT + (i + i0) + n * (k + i0));
        xp[k] = r + t;
        xm[k] = t - r;
      }
      double g3r[D], g4r[D];
      SSB_UNROLL
      for (int jd = 0; jd < D; ++jd) {
        g3r[jd] = SSB_OUT(Ps, eSup + (i + i0) + n * (jd + r0));
        g4r[jd] = SSB_OUT(Ps, eSdn + (i + i0) + n * (jd + r0)) * e0[jd];
      }
      SSB_UNROLL
      for (int j = 0; j < D; ++j) {
        double sum = 0.0, dif = 0.0;
        SSB_UNROLL
        for (int jd = 0; jd < D; ++jd) {
          sum = fma(g3r[jd] + g4r[jd], G0i[jd + D * j], sum);
          dif = fma(g3r[jd] - g4r[jd], G0i[jd + D * j], dif);
        }
        SSB_UNROLL
        for (int k = 0; k < N; ++k) {
          sum = fma(xp[k], rp[k + N * j], sum);
          dif = fma(xm[k], rm[k + N * j], dif);
        }
        SSB_FINAL(P, eSup + (i + i0) + n * (j + r0), 0.5 * (sum + dif));
        SSB_FINAL(P, eSdn + (i + i0) + n * (j + r0), 0.5 * (sum - dif));
      }
    }
  }
  return !bad;
}

// Longwave layer: R, T, int_flux, source, int_flux_source (calc_matrices_lw_eig) with
//   y = -S^-1 b, source = (I - R - T) y, int_flux = 2 V Z A+^-1, Z = diag((1-e)/lambda),
//   int_flux_source = 2 (y dz - int_flux y).   `brate(i)` is the emission rate vector b.
template <int NREG, int NS, int NR, int R0>
SSB_HDI bool layer_lw_solve(const LayerCoef<NR, NS> &c, const double *brate, double dz, double *P,
                            const StateMem &st_in) {
  typedef LayerStack<NR, NS> Stk;
  constexpr int N = NR * NS, n = NREG * NS, i0 = R0 * NS;
  constexpr int eR = 0, eT = n * n, eIF = 2 * n * n, eSrc = 3 * n * n, eIsrc = eSrc + n;
  {
    double LUs[N * N], y[N];
    SSB_UNROLL
    for (int j = 0; j < N; ++j) {
      SSB_UNROLL
      for (int i = 0; i < N; ++i) LUs[i + N * j] = c.S(i, j);
    }
    sm_lu<N>(LUs);
    SSB_UNROLL
    for (int i = 0; i < N; ++i) y[i] = -brate[i];
    sm_lu_solve_left<N, 1>(LUs, y);
    SSB_UNROLL
    for (int i = 0; i < N; ++i) st_in(Stk::oX + i) = y[i];
  }
  double lam2[N], lam[N], e[N];
  bool bad = !layer_eigen<NR, NS>(c, dz, st_in, lam2, lam, e);
  const StateMem st = opaque(st_in);
  {
    double A[N * N], z[N], y[N];
    SSB_UNROLL
    for (int k = 0; k < N; ++k) {
      z[k] = 2.0 * (1.0 - e[k]) / lam[k];
      y[k] = st(Stk::oX + k);
    }
    // sum then difference problem (rolled); in the sum stage also int_flux = (2 V Z) A+^-1, whose
    // rows pass through the IF slot, and the two source vectors
    SSB_ROLLED
    for (int sg = 0; sg < 2; ++sg) {
      layer_stage_factor<NREG, NS, NR, R0>(c, st, sg ? -1.0 : 1.0, lam, e, A, P, eT, sg ? nullptr : z, eIF);
      const double *Pr = opaque_ptr(P);
      double xn[N], xpn[N];  // next row of B; next row of 2VZ (first stage) or of X+ (second stage)
      layer_load_row<NREG, NS, NR, R0>(Pr, eT, 0, xn);
      layer_load_row<NREG, NS, NR, R0>(Pr, sg ? eR : eIF, 0, xpn);
      SSB_ROLLED
      for (int i = 0; i < N; ++i) {
        double x[N], xpr[N];
        SSB_UNROLL
        for (int k = 0; k < N; ++k) {
          x[k] = xn[k];
          xpr[k] = xpn[k];
        }
        layer_load_row<NREG, NS, NR, R0>(Pr, eT, i + 1, xn);
        layer_load_row<NREG, NS, NR, R0>(Pr, sg ? eR : eIF, i + 1, xpn);
        sm_lu_solve_right<1, N>(A, x);
        if (sg == 0) {
          const double yi = st(Stk::oX + i);
          double s = 0.0;
          SSB_UNROLL
          for (int k = 0; k < N; ++k) {
            s = fma(x[k], y[k], s);
            SSB_OUT(P, eR + (i + i0) + n * (k + i0)) = x[k];
          }
          SSB_FINAL(P, eSrc + i + i0, yi - s);
          sm_lu_solve_right<1, N>(A, xpr);
          double f = 0.0;
          SSB_UNROLL
          for (int k = 0; k < N; ++k) {
            f = fma(xpr[k], y[k], f);
            SSB_FINAL(P, eIF + (i + i0) + n * (k + i0), xpr[k]);
          }
          SSB_FINAL(P, eIsrc + i + i0, 2.0 * (yi * dz - f));
        } else {
          SSB_UNROLL
          for (int k = 0; k < N; ++k) {
            const double xp = xpr[k];
            const double r = 0.5 * (xp + x[k]), t = 0.5 * (xp - x[k]);
            bad = bad || !(fabs(r) < 1.0e300) || !(fabs(t) < 1.0e300);
            SSB_FINAL(P, eR + (i + i0) + n * (k + i0), r);
            SSB_FINAL(P, eT + (i + i0) + n * (k + i0), t);
          }
        }
      }
    }
  }
  return !bad;
}

}  // namespace ssb
