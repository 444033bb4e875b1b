// ssb_fused.cuh - column-resident solver: one thread owns one (column, interval) from the
// ground to the canopy top and back.
//
// The split path (ssb_fast_layer.cuh + ssb_fast_sweeps.cuh) hands every layer's matrices
// (R, T, S, E, integrated-flux matrices: 200 doubles at 2 streams, 3 regions) and the
// interface state of the upward sweep (a_above, d_above, LU factors: 90 doubles) through HBM:
// 17 KB of DRAM traffic per (column, layer) against 540 B of inputs and outputs.  Here
//
//  * a thread loops over the layers of its column: geometry -> layer matrices (the same
//    layer_sw_solve / layer_lw_solve) -> upward adding step, with the layer matrices and the
//    carried state in a PRIVATE scratch tile of the thread block that is re-used for every
//    layer of every column the block solves (persistent grid): it lives in L2 and never
//    needs to reach DRAM;
//  * the only thing the downward pass needs from the upward one is, per layer, the LINEAR
//    MAP from what enters the layer top (diffuse flux vector, direct flux vector) to what
//    the downward pass produces there: the fluxes leaving the layer base and the handful of
//    flux functionals the outputs are made of (absorption by clear air / air in vegetation /
//    vegetation, wall interception, their direct-beam parts).  That map - the "operator
//    record" - is formed in the upward step, where every ingredient is at hand, and is all
//    that goes through HBM: 4n + 6d + n^2 + nd + d^2 + 1 = 106 doubles per full layer at
//    n = 6, d = 3 (35 for a layer that solves the clear region only) instead of 290.
//    Columns of the record that belong to regions the layer does not solve are exactly zero
//    and are neither stored nor loaded.
//
// The recurrences are the reference's (radsurf_urban_sw.F90:603-984, radsurf_urban_lw.F90:
// 548-858 and the forest equivalents), including the solve with (I - a_above R) in the
// downward pass; with D = I - a_above R, H = I - a_above:
//     P = D^-1 T                       x_above = P x_below + Q dir_below
//     Q = D^-1 (R d_above E + S_dn)    up_above = a_above x_above + d_above E dir_below
//     a_below = R + T D^-1 a_above T   d_below = S_up + T D^-1 (d_above E + a_above S_dn)
//     x_below - x_above - up_below + up_above = (I - a_below - H P) x_below
//                                               + (d_above E - d_below - H Q) dir_below
// and the longwave analogue with the emission source as a constant column.
#pragma once
#include "ssb_fast_sweeps.cuh"

// unroll factor of the per-column loops of the upward steps (1 = rolled: smallest code; 2 or 3 let
// the two triangular solves of neighbouring columns overlap)
#ifndef SSB_REC_UNROLL
#define SSB_REC_UNROLL 1
#endif
#if defined(__CUDACC__)
#define SSB_PRAGMA_(x) _Pragma(#x)
#define SSB_PRAGMA(x) SSB_PRAGMA_(x)
#define SSB_REC_LOOP SSB_PRAGMA(unroll SSB_REC_UNROLL)
#else
#define SSB_REC_LOOP
#endif

namespace ssb {

// Phase alignment of the warps of a block (device only; a.fused bit 1): the loop body of a column
// is ~250 KB of straight-line code, far more than the instruction caches hold, so warps that
// drift apart each stream it from L2 on their own (ncu: 3 stalled warps per issue on
// `no_instruction`); four warps that enter a phase together fetch it once.
SSB_HDI void phase_sync(const ClassArgs &a) {
#if defined(__CUDA_ARCH__)
  if (a.fused & 2) __syncthreads();
#else
  (void)a;
#endif
}

SSB_HDI bool region_solved(int seg, int r) { return seg == 0 || (seg == 1 ? r == 0 : r > 0); }
// class of an element (see seg_class) with a run-time column region
SSB_HDI bool keep_rc(int seg, int ri, int rj) {
  return seg == 0 || (seg == 1 ? (ri == 0 && rj == 0) : (ri > 0 && rj > 0));
}

// The private tile of the calling thread (one level; re-used for every layer: default cache
// policy, unlike the streamed operator records).
struct Tile {
  double *base;
  size_t lev_stride;  // 0: one private level (column-resident kernels); else the layer scratch of the split path
  SSB_HDI Tile(const ClassArgs &a, int q)
      : base(a.layer + layer_sidx(a, 0, 0, q)), lev_stride(a.fused ? 0 : (size_t)a.ne_layer * kScratchTile) {}
  SSB_HDI Tile at(int lev) const {
    Tile t = *this;
    t.base += (size_t)lev * lev_stride;
    return t;
  }
  SSB_HDI double ld(int e, int) const { return base[(size_t)e * kScratchTile]; }
  SSB_HDI double ldp(int e, int, bool keep) const {  // structural zeros are not loaded
    double v = 0.0;
    if (keep) v = base[(size_t)e * kScratchTile];
    return v;
  }
  SSB_HDI void st(int e, int, double v) const { base[(size_t)e * kScratchTile] = v; }
};

// ===========================================================================
// Shortwave
// ===========================================================================
template <int NREG, int NS, bool URBAN>
struct SwFused {
  typedef SwSweepLayout<NREG, NS, URBAN> L;
  static constexpr int n = NREG * NS, d = NREG;
  // private tile: layer elements, geometry block, carried state (a_below and d_below are parked
  // over R and S_up)
  static constexpr int oState = L::oGeo + kGeoElems;  // a_above (n x n), d_above (n x d)
  static constexpr int private_elems = oState + n * n + n * d;
  // operator record of one level
  static constexpr int NF = 4;   // clear_air_abs, veg_air_abs, veg_abs, wall_in (diffuse part)
  static constexpr int NFD = 6;  // ... and veg_abs_dir, wall_in_dir
  static constexpr int mFx = 0;                  // NF x n
  static constexpr int mFd = mFx + NF * n;       // NFD x d
  static constexpr int mP = mFd + NFD * d;       // n x n
  static constexpr int mQ = mP + n * n;          // n x d
  static constexpr int mE = mQ + n * d;          // d x d
  static constexpr int mScal = mE + d * d;       // f_wall_dir_clear, segment
  static constexpr int mProf = mScal + 2;        // 2 x (n + d): sum(up_below), sum(up_above)
  static constexpr int op_elems = mProf + 2 * (n + d);
  // shared-memory slice: the layer stack during the layer solve; a_above, T, d_above in the step
  static constexpr int sAa = 0, sDa = n * n, sT = n * n + n * d;  // [a_above | d_above] = the carried state
  static constexpr int step_doubles = 2 * n * n + n * d;
  static constexpr int ls = LayerStack<NREG, NS>::sw_doubles;
  static constexpr int smem_doubles = ls > step_doubles ? ls : step_doubles;
};

// One upward adding step of layer jl: reads the layer matrices and the state from the private
// tile, writes the operator record of the level and the state above the next interface.
// REC: the carried state stays in the shared-memory slice between the steps (split layer kernels
// + record sweeps, where the tile is the layer scratch of level jl) instead of in the private tile.
template <int NREG, int NS, bool URBAN, bool REC = false>
SSB_HD inline void fused_up_step_sw(const ClassArgs &a, const Tile &Lp, const Scr &Mo, int jl, int il, int il1,
                                    int nlay, int g, const StateMem &sm, double zcos, double sin0) {
  const int ls = layer_step(a);
  typedef SwFused<NREG, NS, URBAN> F;
  typedef SwSweepLayout<NREG, NS, URBAN> Lay;
  constexpr int n = NREG * NS, d = NREG, NRB = URBAN ? NREG + 1 : NREG, NF = F::NF, NFD = F::NFD;
  const int nspec = a.cfg.nspec;
  const int seg = (int)Lp.ld(Lay::oGeo + 7, 0);
  const SegKeep sk = seg_keep(seg);
  // ---- operands that every column needs, into shared memory --------------------------------
  if (!REC) {
    SSB_UNROLL
    for (int i = 0; i < n * n; ++i) sm(F::sAa + i) = Lp.ld(F::oState + i, 0);
    SSB_UNROLL
    for (int i = 0; i < n * d; ++i) sm(F::sDa + i) = Lp.ld(F::oState + n * n + i, 0);
  }
  SSB_UNROLL
  for (int j = 0; j < n; ++j) {
    SSB_UNROLL
    for (int i = 0; i < n; ++i) sm(F::sT + i + n * j) = Lp.ldp(Lay::oT + i + n * j, 0, sk.k[seg_class<0, NS>(i, j)]);
  }
  // ---- D = I - a_above R, factored ----------------------------------------------------------
  double LU[n * n];
  SSB_UNROLL
  for (int j = 0; j < n; ++j) {
    SSB_UNROLL
    for (int i = 0; i < n; ++i) LU[i + n * j] = (i == j) ? 1.0 : 0.0;
    SSB_UNROLL
    for (int k = 0; k < n; ++k) {
      const double r = Lp.ldp(Lay::oR + k + n * j, 0, sk.k[seg_class<0, NS>(k, j)]);
      SSB_UNROLL
      for (int i = 0; i < n; ++i) LU[i + n * j] = fma(-sm(F::sAa + i + n * k), r, LU[i + n * j]);
    }
  }
  sm_lu<n>(LU);
  // ---- flux functionals: weights x integrated-flux matrices ---------------------------------
  double W[NF * n], G[NFD * d];
  {
    double od[3], fw[3];
    SSB_UNROLL
    for (int r = 0; r < 3; ++r) {
      fw[r] = Lp.ld(Lay::oGeo + r, 0);
      od[r] = Lp.ld(Lay::oGeo + 3 + r, 0);
    }
    const bool veg = NREG > 1 || !URBAN;
    const double ve = (veg && a.cp.veg_ext) ? a.cp.veg_ext[il] : 0.0;
    const double air_abs = SSB_LAY(a.sw.air_ext, g, il) * (1.0 - SSB_LAY(a.sw.air_ssa, g, il));
    const double vabs = NREG > 1 ? ve * (1.0 - SSB_LAY(a.sw.veg_ssa, g, il)) : 0.0;
    double wt[NF * n];  // [f + NF * i]
    SSB_UNROLL
    for (int i = 0; i < n; ++i) {
      const int r = i / NS, js = i % NS;
      const double rmu = 1.0 / a.lg.mu[js];
      wt[0 + NF * i] = (r == 0) ? air_abs * rmu : 0.0;
      wt[1 + NF * i] = (r > 0) ? air_abs * rmu : 0.0;
      wt[2 + NF * i] = (r > 0) ? vabs * od[r] * rmu : 0.0;
      wt[3 + NF * i] = URBAN ? fw[r] * a.lg.tan_ang[js] : 0.0;
    }
    SSB_UNROLL
    for (int i = 0; i < NF * n; ++i) W[i] = 0.0;
    SSB_UNROLL
    for (int j = 0; j < n; ++j) {
      SSB_UNROLL
      for (int i = 0; i < n; ++i) {
        const double v = Lp.ldp(Lay::oIdiff + i + n * j, 0, sk.k[seg_class<0, NS>(i, j)]);
        SSB_UNROLL
        for (int f = 0; f < NF; ++f) W[f + NF * j] = fma(wt[f + NF * i], v, W[f + NF * j]);
      }
    }
    // G: what multiplies (I - E) dir_below: rows 0..3 from int_dir_diff (+ int_dir), 4..5 from int_dir
    SSB_UNROLL
    for (int i = 0; i < NFD * d; ++i) G[i] = 0.0;
    SSB_UNROLL
    for (int j = 0; j < d; ++j) {
      SSB_UNROLL
      for (int i = 0; i < n; ++i) {
        const double v = Lp.ldp(Lay::oIdd + i + n * j, 0, sk.k[seg_class<1, NS>(i, j)]);
        SSB_UNROLL
        for (int f = 0; f < NF; ++f) G[f + NFD * j] = fma(wt[f + NF * i], v, G[f + NFD * j]);
      }
      SSB_UNROLL
      for (int r = 0; r < d; ++r) {
        const double v = Lp.ldp(Lay::oIdir + r + d * j, 0, sk.k[seg_class<2, NS>(r, j)]);
        if (r == 0) G[0 + NFD * j] = fma(air_abs, v, G[0 + NFD * j]);
        if (r > 0) {
          G[1 + NFD * j] = fma(air_abs, v, G[1 + NFD * j]);
          G[2 + NFD * j] = fma(vabs * od[r], v, G[2 + NFD * j]);
          G[4 + NFD * j] = fma(vabs * od[r], v, G[4 + NFD * j]);
        }
        if (URBAN) G[5 + NFD * j] = fma(fw[r] * sin0, v, G[5 + NFD * j]);
      }
    }
  }
  const bool prof = a.save_profile != 0;
  // ---- columns of the direct part (first: they read all of R, which the diffuse columns then
  //      overwrite with a_below, column by column) ----------------------------------------------
  SSB_REC_LOOP
  for (int j = 0; j < d; ++j) {
    if (!region_solved(seg, j)) {
      SSB_UNROLL
      for (int i = 0; i < n; ++i) Lp.st(Lay::oSup + i + n * j, 0, 0.0);
      continue;
    }
    double ecol[d], v[n], q[n], w[n];
    SSB_UNROLL
    for (int k = 0; k < d; ++k) ecol[k] = Lp.ldp(Lay::oE + k + d * j, 0, keep_rc(seg, k, j));
    SSB_UNROLL
    for (int i = 0; i < n; ++i) {
      double s = 0.0;
      SSB_UNROLL
      for (int k = 0; k < d; ++k) s = fma(sm(F::sDa + i + n * k), ecol[k], s);
      v[i] = s;  // (d_above E)(:, j)
    }
    // q = S_dn_j + R v, w = v + a_above S_dn_j
    {
      double sdn[n];
      SSB_UNROLL
      for (int i = 0; i < n; ++i) {
        sdn[i] = Lp.ldp(Lay::oSdn + i + n * j, 0, keep_rc(seg, i / NS, j));
        q[i] = sdn[i];
        w[i] = v[i];
      }
      SSB_UNROLL
      for (int k = 0; k < n; ++k) {
        SSB_UNROLL
        for (int i = 0; i < n; ++i) {
          w[i] = fma(sm(F::sAa + i + n * k), sdn[k], w[i]);
          const double r = Lp.ldp(Lay::oR + i + n * k, 0, keep_rc(seg, i / NS, k / NS));
          q[i] = fma(r, v[k], q[i]);
        }
      }
    }
    sm_lu_solve_left<n, 1>(LU, q);
    sm_lu_solve_left<n, 1>(LU, w);
    double db[n], aq[n];
    SSB_UNROLL
    for (int i = 0; i < n; ++i) {
      db[i] = Lp.ldp(Lay::oSup + i + n * j, 0, keep_rc(seg, i / NS, j));
      aq[i] = 0.0;
    }
    SSB_UNROLL
    for (int k = 0; k < n; ++k) {
      SSB_UNROLL
      for (int i = 0; i < n; ++i) {
        db[i] = fma(sm(F::sT + i + n * k), w[k], db[i]);
        aq[i] = fma(sm(F::sAa + i + n * k), q[k], aq[i]);
      }
    }
    double sdb = 0.0, sua = 0.0;
    SSB_UNROLL
    for (int i = 0; i < n; ++i) {
      Lp.st(Lay::oSup + i + n * j, 0, db[i]);  // d_below over S_up (read above)
      Mo.st(F::mQ + i + n * j, jl, q[i]);
      sdb += db[i];
      sua += aq[i] + v[i];
    }
    SSB_UNROLL
    for (int k = 0; k < d; ++k) Mo.st(F::mE + k + d * j, jl, ecol[k]);
    if (prof) {
      Mo.st(F::mProf + n + j, jl, sdb);
      Mo.st(F::mProf + (n + d) + n + j, jl, sua);
    }
    double fd[NFD];
    SSB_UNROLL
    for (int f = 0; f < NFD; ++f) fd[f] = 0.0;
    SSB_UNROLL
    for (int i = 0; i < n; ++i) {
      const double c = v[i] - db[i] - q[i] + aq[i];
      SSB_UNROLL
      for (int f = 0; f < NF; ++f) fd[f] = fma(W[f + NF * i], c, fd[f]);
    }
    SSB_UNROLL
    for (int k = 0; k < d; ++k) {
      const double ie = ((k == j) ? 1.0 : 0.0) - ecol[k];
      SSB_UNROLL
      for (int f = 0; f < NFD; ++f) fd[f] = fma(G[f + NFD * k], ie, fd[f]);
    }
    SSB_UNROLL
    for (int f = 0; f < NFD; ++f) Mo.st(F::mFd + f + NFD * j, jl, fd[f]);
  }
  // ---- columns of the diffuse part (rolled: the same code for every column) -----------------
  SSB_REC_LOOP
  for (int j = 0; j < n; ++j) {
    const int rj = j / NS;
    if (!region_solved(seg, rj)) {
      SSB_UNROLL
      for (int i = 0; i < n; ++i) Lp.st(Lay::oR + i + n * j, 0, 0.0);
      continue;
    }
    double p[n], x[n];
    SSB_UNROLL
    for (int i = 0; i < n; ++i) {
      p[i] = sm(F::sT + i + n * j);
      x[i] = 0.0;
    }
    SSB_UNROLL
    for (int k = 0; k < n; ++k) {
      SSB_UNROLL
      for (int i = 0; i < n; ++i) x[i] = fma(sm(F::sAa + i + n * k), p[k], x[i]);
    }
    sm_lu_solve_left<n, 1>(LU, p);
    sm_lu_solve_left<n, 1>(LU, x);
    double ab[n], ap[n];
    SSB_UNROLL
    for (int i = 0; i < n; ++i) {
      ab[i] = Lp.ldp(Lay::oR + i + n * j, 0, keep_rc(seg, i / NS, rj));
      ap[i] = 0.0;
    }
    SSB_UNROLL
    for (int k = 0; k < n; ++k) {
      SSB_UNROLL
      for (int i = 0; i < n; ++i) {
        ab[i] = fma(sm(F::sT + i + n * k), x[k], ab[i]);
        ap[i] = fma(sm(F::sAa + i + n * k), p[k], ap[i]);
      }
    }
    double sab = 0.0, sap = 0.0;
    SSB_UNROLL
    for (int i = 0; i < n; ++i) {
      Lp.st(Lay::oR + i + n * j, 0, ab[i]);  // a_below over R (column j is not read again)
      Mo.st(F::mP + i + n * j, jl, p[i]);
      sab += ab[i];
      sap += ap[i];
    }
    if (prof) {
      Mo.st(F::mProf + j, jl, sab);
      Mo.st(F::mProf + (n + d) + j, jl, sap);
    }
    // c = e_j - a_below_j - (p - a_above p); functionals W c
    double fx[NF];
    SSB_UNROLL
    for (int f = 0; f < NF; ++f) fx[f] = 0.0;
    SSB_UNROLL
    for (int i = 0; i < n; ++i) {
      const double c = ((i == j) ? 1.0 : 0.0) - ab[i] - p[i] + ap[i];
      SSB_UNROLL
      for (int f = 0; f < NF; ++f) fx[f] = fma(W[f + NF * i], c, fx[f]);
    }
    SSB_UNROLL
    for (int f = 0; f < NF; ++f) Mo.st(F::mFx + f + NF * j, jl, fx[f]);
  }
  Mo.st(F::mScal, jl, Lp.ld(Lay::oGeo + 6, 0));
  Mo.st(F::mScal + 1, jl, (double)seg);
  // ---- roofs, overlap: state above the next interface (into the private tile) ---------------
  double hw[NS];
  SSB_UNROLL
  for (int js = 0; js < NS; ++js) hw[js] = a.lg.hweight[js];
  double rb[NS], rd[NS];
  SSB_UNROLL
  for (int js = 0; js < NS; ++js) rb[js] = rd[js] = 0.0;
  if (URBAN) {
    const double ralb = SSB_LAY(a.sw.roof_albedo, g, il);
    const double ralb_dir = a.sw.roof_albedo_dir ? SSB_LAY(a.sw.roof_albedo_dir, g, il) : ralb;
    SSB_UNROLL
    for (int js = 0; js < NS; ++js) {
      rb[js] = ralb * hw[js];
      rd[js] = zcos * ralb_dir * hw[js];
    }
  }
  double U[12], V[12];
  overlap_above<NREG, URBAN, false>(a, il1, nlay, jl, U, V);
  const StateMem stt = REC ? StateMem{&sm(F::sAa), sm.stride}
                           : StateMem{Lp.base + (size_t)F::oState * kScratchTile, kScratchTile};
  {
    double Ab[n * n];
    SSB_UNROLL
    for (int i = 0; i < n * n; ++i) Ab[i] = Lp.ld(Lay::oR + i, 0);
    overlap_matrix<NREG, NRB, NS>(Ab, rb, U, V, stt, 0);
  }
  {
    double Db[n * d];
    SSB_UNROLL
    for (int i = 0; i < n * d; ++i) Db[i] = Lp.ld(Lay::oSup + i, 0);
    SSB_UNROLL
    for (int jt = 0; jt < NS; ++jt) {
      double DV[NREG * NREG];
      SSB_UNROLL
      for (int up = 0; up < NREG; ++up) {
        SSB_UNROLL
        for (int lo = 0; lo < NREG; ++lo) {
          double s = 0.0;
          SSB_UNROLL
          for (int l2 = 0; l2 < NREG; ++l2) s = fma(Db[(lo * NS + jt) + n * l2], V[l2 + NRB * up], s);
          DV[lo + NREG * up] = s;
        }
      }
      SSB_UNROLL
      for (int up = 0; up < NREG; ++up) {
        SSB_UNROLL
        for (int u = 0; u < NREG; ++u) {
          double s = 0.0;
          SSB_UNROLL
          for (int lo = 0; lo < NREG; ++lo) s = fma(U[u + NREG * lo], DV[lo + NREG * up], s);
          if (URBAN) s = fma(U[u + NREG * NREG] * rd[jt], V[NREG + NRB * up], s);
          stt(n * n + (u * NS + jt) + n * up) = s;
        }
      }
    }
  }
  (void)nspec;
}

// one layer of the column: geometry, layer matrices into the private tile
template <int NREG, int NS>
SSB_HD inline void fused_layer_sw(const ClassArgs &a, int q, int lev, const StateMem &st) {
  if (fast_prepare_level(a, q, lev) < 0) return;
  fast_layer_problem_sw<NREG, NS>(a, q, lev, st);
}

// MODE 0: the whole column in one thread (column-resident kernels).  MODE 1 / 2: the two halves of the
// "record sweeps" that follow the split layer kernels - 1: upward pass over the layer scratch, writes the
// operator records and the boundary conditions; 2: downward pass through the records, writes the fluxes.
template <int NREG, int NS, bool URBAN, int MODE = 0>
SSB_HD inline void fused_column_sw(const ClassArgs &a, int q_in, bool active, const StateMem &st) {
  constexpr bool REC = MODE != 0;
  typedef SwFused<NREG, NS, URBAN> F;
  constexpr int n = NREG * NS, d = NREG, NRB = URBAN ? NREG + 1 : NREG, NF = F::NF, NFD = F::NFD;
  const SolveCfg &c = a.cfg;
  const int nspec = c.nspec;
  const int q = active ? q_in : 0;  // (idle threads of the last tile only keep the barriers company)
  const int ic = q / nspec, g = q % nspec;
  const int col = a.cols[ic];
  const int nlay = a.nlay[col], il1 = layer_index(a, ic, col, 0), ls = layer_step(a);
  const ssb200_canopy_flux &fdir = a.f1, &fdif = a.f2;
  const double cos_sza = a.cp.cos_sza[col];
  int itransp = 0;
  if (nspec > 1) {
    double best = 0.0;
    for (int gg = 0; gg < nspec; ++gg) {
      double od = 0.0;
      for (int l = 0; l < nlay; ++l) od += a.sw.air_ext[(size_t)gg + (size_t)nspec * (il1 + l * ls)] * a.cp.dz[il1 + l * ls];
      if (gg == 0 || od < best) {
        best = od;
        itransp = gg;
      }
    }
  }
  const bool own = (g == itransp);
  const bool live = active && (cos_sza > 0.0);
  if (MODE != 2 && active && !live) {  // night: every member of the column is zero (radsurf_interface.F90:193-196)
    zero_column(fdir, nspec, g, col, il1, nlay, own, ls);
    zero_column(fdif, nspec, g, col, il1, nlay, own, ls);
  }
  if (!live) {
    if (!REC && (a.fused & 2))
      for (int jl = 0; jl < a.lmax; ++jl) {
        phase_sync(a);
        phase_sync(a);
      }
    return;
  }
  if (MODE != 2) {
    zero_unwritten_sw<NREG, URBAN>(fdir, nspec, g, col, il1, nlay, own, true, ls);
    zero_unwritten_sw<NREG, URBAN>(fdif, nspec, g, col, il1, nlay, own, false, ls);
  }
  const double zcos = URBAN ? dmax(cos_sza, 1.0e-6) : cos_sza;
  const double sin0 = URBAN ? sqrt(1.0 - zcos * zcos) : 0.0;
  double hw[NS], tang[NS];
  SSB_UNROLL
  for (int js = 0; js < NS; ++js) {
    hw[js] = a.lg.hweight[js];
    tang[js] = a.lg.tan_ang[js];
  }
  const Tile Lp(a, q);
  const Scr Mo(a.sweep, a.lmax, a.ne_sweep, q);
  // ---- upward: state = [a_above | d_above] in the private tile (REC: in the shared-memory slice) ----
  double talb_diff = 0.0, talb_dir = 0.0;
  if (MODE != 2) {
    // element e of the carried state
    auto state_st = [&](int e, double v) {
      if (REC)
        st(F::sAa + e) = v;
      else
        Lp.st(F::oState + e, 0, v);
    };
    auto state_ld = [&](int e) -> double { return REC ? st(F::sAa + e) : Lp.ld(F::oState + e, 0); };
    const double galb = a.sw.ground_albedo[(size_t)g + (size_t)nspec * col];
    const double galb_dir =
        (a.use_sw_direct_albedo ? a.sw.ground_albedo_dir : a.sw.ground_albedo)[(size_t)g + (size_t)nspec * col];
    SSB_UNROLL
    for (int i = 0; i < n * n + n * d; ++i) state_st(i, 0.0);
    SSB_UNROLL
    for (int r = 0; r < NREG; ++r) {
      SSB_UNROLL
      for (int jt = 0; jt < NS; ++jt) {
        state_st(n * n + (jt + r * NS) + n * r, zcos * galb_dir * hw[jt]);
        SSB_UNROLL
        for (int jf = 0; jf < NS; ++jf) state_st((jt + r * NS) + n * (jf + r * NS), galb * hw[jt]);
      }
    }
    if (REC) {
      for (int jl = 0; jl < nlay; ++jl)
        fused_up_step_sw<NREG, NS, URBAN, REC>(a, Lp.at(jl), Mo, jl, il1 + jl * ls, il1, nlay, g, st, zcos, sin0);
    } else {
      const int nloop = (a.fused & 2) ? a.lmax : nlay;
      for (int jl = 0; jl < nloop; ++jl) {
        phase_sync(a);
        if (jl < nlay) fused_layer_sw<NREG, NS>(a, q, jl, st);
        phase_sync(a);
        if (jl < nlay) fused_up_step_sw<NREG, NS, URBAN>(a, Lp, Mo, jl, il1 + jl * ls, il1, nlay, g, st, zcos, sin0);
      }
    }
    SSB_UNROLL
    for (int i = 0; i < NS; ++i) {
      double s = 0.0;
      SSB_UNROLL
      for (int j = 0; j < NS; ++j) s = fma(state_ld(i + n * j), hw[j], s);
      talb_diff += s;
    }
    double s = 0.0;
    SSB_UNROLL
    for (int js = 0; js < NS; ++js) s += state_ld(n * n + js);
    talb_dir = s / zcos;
    a.bc.sw_albedo[(size_t)g + (size_t)nspec * col] = talb_diff;
    a.bc.sw_albedo_dir[(size_t)g + (size_t)nspec * col] = talb_dir;
  }
  if (MODE == 1) return;
  if (MODE == 2) {
    talb_diff = a.bc.sw_albedo[(size_t)g + (size_t)nspec * col];
    talb_dir = a.bc.sw_albedo_dir[(size_t)g + (size_t)nspec * col];
  }
  // ---- downward: direct (suffix d) and diffuse (suffix f) sources through the records --------
  double dir_above[d], xa_d[n], xa_f[n];
  SSB_UNROLL
  for (int i = 0; i < d; ++i) dir_above[i] = 0.0;
  SSB_UNROLL
  for (int i = 0; i < n; ++i) xa_d[i] = xa_f[i] = 0.0;
  dir_above[0] = 1.0 / zcos;
  SSB_UNROLL
  for (int js = 0; js < NS; ++js) xa_f[js] = hw[js];
  SSB_FC(fdir, top_dn_dir) = 1.0;
  SSB_FC(fdir, top_dn) = 1.0;
  SSB_FC(fdir, top_net) = 1.0 * (1.0 - talb_dir);
  SSB_FC(fdif, top_dn_dir) = 0.0;
  SSB_FC(fdif, top_dn) = 1.0;
  SSB_FC(fdif, top_net) = 1.0 - talb_diff;
  if (URBAN && own && fdir.roof_sunlit_frac && nlay > 0) fdir.roof_sunlit_frac[il1 + (nlay - 1) * ls] = 1.0;
  double flux_dn_dir_clear = 1.0 / zcos;
  double ua_sum_d = 0.0, ua_sum_f = 0.0, vt_d = 0.0, vt_f = 0.0;  // at the ground, from the last record
  const bool prof = a.save_profile != 0;
  for (int jl = nlay - 1; jl >= 0; --jl) {
    const int il = il1 + jl * ls;
    const int seg = (int)Mo.ld(F::mScal + 1, jl);
    const double f_wall_dir_clear = Mo.ld(F::mScal, jl);
    const bool veg = NREG > 1 || !URBAN;
    const double bf = URBAN ? a.cp.building_fraction[il] : 0.0;
    const double vf = (veg && a.cp.veg_fraction) ? a.cp.veg_fraction[il] : 0.0;
    const double ve = (veg && a.cp.veg_ext) ? a.cp.veg_ext[il] : 0.0;
    double xb_d[NRB * NS], xb_f[NRB * NS], dir_below[NRB];
    {
      double U[12], V[12];
      overlap_above<NREG, URBAN, false>(a, il1, nlay, jl, U, V);
      expand_down<NREG, NRB, NS>(V, xa_d, xb_d);
      expand_down<NREG, NRB, NS>(V, xa_f, xb_f);
      SSB_UNROLL
      for (int lo = 0; lo < NRB; ++lo) {
        double s = 0.0;
        SSB_UNROLL
        for (int up = 0; up < NREG; ++up) s = fma(V[lo + NRB * up], dir_above[up], s);
        dir_below[lo] = s;
      }
    }
    double fn_d[NFD], fn_f[NF], sp_d[2] = {0.0, 0.0}, sp_f[2] = {0.0, 0.0};
    SSB_UNROLL
    for (int f = 0; f < NFD; ++f) fn_d[f] = 0.0;
    SSB_UNROLL
    for (int f = 0; f < NF; ++f) fn_f[f] = 0.0;
    SSB_UNROLL
    for (int i = 0; i < n; ++i) xa_d[i] = xa_f[i] = 0.0;
    SSB_UNROLL
    for (int i = 0; i < d; ++i) dir_above[i] = 0.0;
    SSB_UNROLL
    for (int j = 0; j < n; ++j) {
      if (region_solved(seg, j / NS)) {
        SSB_UNROLL
        for (int i = 0; i < n; ++i) {
          const double p = Mo.ld(F::mP + i + n * j, jl);
          xa_d[i] = fma(p, xb_d[j], xa_d[i]);
          xa_f[i] = fma(p, xb_f[j], xa_f[i]);
        }
        SSB_UNROLL
        for (int f = 0; f < NF; ++f) {
          const double w = Mo.ld(F::mFx + f + NF * j, jl);
          fn_d[f] = fma(w, xb_d[j], fn_d[f]);
          fn_f[f] = fma(w, xb_f[j], fn_f[f]);
        }
        if (prof) {
          const double s0 = Mo.ld(F::mProf + j, jl), s1 = Mo.ld(F::mProf + (n + d) + j, jl);
          sp_d[0] = fma(s0, xb_d[j], sp_d[0]);
          sp_f[0] = fma(s0, xb_f[j], sp_f[0]);
          sp_d[1] = fma(s1, xb_d[j], sp_d[1]);
          sp_f[1] = fma(s1, xb_f[j], sp_f[1]);
        }
      }
    }
    SSB_UNROLL
    for (int j = 0; j < d; ++j) {
      if (region_solved(seg, j)) {
        SSB_UNROLL
        for (int i = 0; i < n; ++i) xa_d[i] = fma(Mo.ld(F::mQ + i + n * j, jl), dir_below[j], xa_d[i]);
        SSB_UNROLL
        for (int f = 0; f < NFD; ++f) fn_d[f] = fma(Mo.ld(F::mFd + f + NFD * j, jl), dir_below[j], fn_d[f]);
        SSB_UNROLL
        for (int k = 0; k < d; ++k) dir_above[k] = fma(Mo.ld(F::mE + k + d * j, jl), dir_below[j], dir_above[k]);
        if (prof) {
          sp_d[0] = fma(Mo.ld(F::mProf + n + j, jl), dir_below[j], sp_d[0]);
          sp_d[1] = fma(Mo.ld(F::mProf + (n + d) + n + j, jl), dir_below[j], sp_d[1]);
        }
      }
    }
    if (URBAN) {
      const double ralb = SSB_LAY(a.sw.roof_albedo, g, il);
      const double ralb_dir = a.sw.roof_albedo_dir ? SSB_LAY(a.sw.roof_albedo_dir, g, il) : ralb;
      double sroof_d = 0.0, sroof_f = 0.0, rup_d = 0.0, rup_f = 0.0;
      SSB_UNROLL
      for (int js = 0; js < NS; ++js) {
        sroof_d += xb_d[n + js];
        sroof_f += xb_f[n + js];
      }
      SSB_UNROLL
      for (int js = 0; js < NS; ++js) {
        rup_d += ralb * hw[js] * sroof_d + zcos * ralb_dir * hw[js] * dir_below[NREG];
        rup_f += ralb * hw[js] * sroof_f;
      }
      SSB_FL(fdir, roof_in_dir, il) = zcos * dir_below[NREG];
      SSB_FL(fdir, roof_in, il) = SSB_FL(fdir, roof_in_dir, il) + sroof_d;
      SSB_FL(fdir, roof_net, il) = SSB_FL(fdir, roof_in, il) - rup_d;
      SSB_FL(fdif, roof_in, il) = sroof_f;
      SSB_FL(fdif, roof_net, il) = sroof_f - rup_f;
    }
    if (fdir.flux_dn_layer_top || fdif.flux_dn_layer_top) {
      double s0 = 0.0, s1 = 0.0, s4 = 0.0, s5 = 0.0, sdb = 0.0, sda = 0.0;
      SSB_UNROLL
      for (int i = 0; i < n; ++i) {
        s0 += xb_d[i];
        s1 += xa_d[i];
        s4 += xb_f[i];
        s5 += xa_f[i];
      }
      SSB_UNROLL
      for (int i = 0; i < d; ++i) {
        sdb += dir_below[i];
        sda += dir_above[i];
      }
      if (fdir.flux_dn_layer_top) {
        SSB_FL(fdir, flux_dn_dir_layer_top, il) = zcos * sdb;
        SSB_FL(fdir, flux_dn_layer_top, il) = zcos * sdb + s0;
        SSB_FL(fdir, flux_dn_dir_layer_base, il) = zcos * sda;
        SSB_FL(fdir, flux_dn_layer_base, il) = zcos * sda + s1;
        SSB_FL(fdir, flux_up_layer_top, il) = sp_d[0];
        SSB_FL(fdir, flux_up_layer_base, il) = sp_d[1];
      }
      if (fdif.flux_dn_layer_top) {
        SSB_FL(fdif, flux_dn_layer_top, il) = s4;
        SSB_FL(fdif, flux_dn_layer_base, il) = s5;
        SSB_FL(fdif, flux_up_layer_top, il) = sp_f[0];
        SSB_FL(fdif, flux_up_layer_base, il) = sp_f[1];
      }
    }
    SSB_FL(fdir, clear_air_abs, il) = fn_d[0];
    SSB_FL(fdif, clear_air_abs, il) = fn_f[0];
    if (NREG > 1) {
      SSB_FL(fdir, veg_air_abs, il) = fn_d[1];
      SSB_FL(fdir, veg_abs, il) = fn_d[2];
      SSB_FL(fdir, veg_abs_dir, il) = fn_d[4];
      SSB_FL(fdif, veg_air_abs, il) = fn_f[1];
      SSB_FL(fdif, veg_abs, il) = fn_f[2];
    }
    if (URBAN) {
      const double walb = SSB_LAY(a.sw.wall_albedo, g, il);
      SSB_FL(fdir, wall_in_dir, il) = fn_d[5];
      SSB_FL(fdir, wall_in, il) = fn_d[5] + fn_d[3];
      SSB_FL(fdir, wall_net, il) = (fn_d[5] + fn_d[3]) * (1.0 - walb);
      SSB_FL(fdif, wall_in, il) = fn_f[3];
      SSB_FL(fdif, wall_net, il) = fn_f[3] * (1.0 - walb);
    }
    {
      // spectrally independent sunlit fractions from the most transparent interval (urban_sw:805-848)
      const double nonb_here = URBAN ? 1.0 - bf : 1.0;
      double nonb_above = 1.0;
      if (URBAN && jl + 1 < nlay) nonb_above = 1.0 - a.cp.building_fraction[il + ls];
      if (URBAN) {
        const double roof_fraction = (jl == nlay - 1) ? bf : dmax(0.0, bf - a.cp.building_fraction[il + ls]);
        if (own && fdir.roof_sunlit_frac)
          fdir.roof_sunlit_frac[il] = (zcos * dir_below[NREG]) * nonb_above /
                                      (zcos * flux_dn_dir_clear * dmax(c.min_bld, roof_fraction));
        flux_dn_dir_clear = flux_dn_dir_clear * nonb_here / nonb_above;
      }
      const double air_ext_t = a.sw.air_ext[(size_t)itransp + (size_t)nspec * il];
      const double trans_dir_clear = exp(-air_ext_t * a.cp.dz[il] / zcos);
      const double int_flux_dir_clear = (air_ext_t > 0.0)
                                            ? flux_dn_dir_clear * (1.0 - trans_dir_clear) * zcos / air_ext_t
                                            : flux_dn_dir_clear * a.cp.dz[il];
      if (own) {
        if ((URBAN ? NREG > 1 : true) && fdir.veg_sunlit_frac && a.cp.veg_ext && a.cp.veg_fraction && a.sw.veg_ssa) {
          const double veg_abs_dir_clear = int_flux_dir_clear * ve * (1.0 - SSB_LAY(a.sw.veg_ssa, g, il)) * vf;
          fdir.veg_sunlit_frac[il] = (NREG > 1 ? fn_d[4] : 0.0) / dmax(SSB_EPS, veg_abs_dir_clear);
        }
        if (URBAN && fdir.wall_sunlit_frac)
          fdir.wall_sunlit_frac[il] = 0.5 * fn_d[5] / dmax(SSB_EPS, (f_wall_dir_clear * sin0 * int_flux_dir_clear));
      }
      flux_dn_dir_clear = flux_dn_dir_clear * trans_dir_clear;
    }
    if (jl == 0) {
      // up_above at the ground = a_above(ground) x_above + d_above(ground) dir_above: the ground albedo
      const double galb = a.sw.ground_albedo[(size_t)g + (size_t)nspec * col];
      const double galb_dir =
          (a.use_sw_direct_albedo ? a.sw.ground_albedo_dir : a.sw.ground_albedo)[(size_t)g + (size_t)nspec * col];
      SSB_UNROLL
      for (int r = 0; r < NREG; ++r) {
        double sd = 0.0, sf = 0.0;
        SSB_UNROLL
        for (int js = 0; js < NS; ++js) {
          sd += xa_d[r * NS + js];
          sf += xa_f[r * NS + js];
        }
        SSB_UNROLL
        for (int js = 0; js < NS; ++js) {
          const double ud = fma(galb * hw[js], sd, zcos * galb_dir * hw[js] * dir_above[r]);
          const double uf = galb * hw[js] * sf;
          ua_sum_d += ud;
          ua_sum_f += uf;
          vt_d += (xa_d[r * NS + js] + ud) * tang[js] / SSB_PI;
          vt_f += (xa_f[r * NS + js] + uf) * tang[js] / SSB_PI;
        }
      }
    }
  }
  {
    double s_dir = 0.0, dn_d = 0.0, dn_f = 0.0;
    SSB_UNROLL
    for (int i = 0; i < d; ++i) s_dir += dir_above[i];
    SSB_UNROLL
    for (int i = 0; i < n; ++i) {
      dn_d += xa_d[i];
      dn_f += xa_f[i];
    }
    SSB_FC(fdir, ground_dn_dir) = zcos * s_dir;
    SSB_FC(fdir, ground_dn) = zcos * s_dir + dn_d;
    SSB_FC(fdir, ground_net) = SSB_FC(fdir, ground_dn) - ua_sum_d;
    SSB_FC(fdir, ground_vertical_diff) = vt_d;
    if (own && fdir.ground_sunlit_frac)
      fdir.ground_sunlit_frac[col] = SSB_FC(fdir, ground_dn_dir) / (zcos * flux_dn_dir_clear);
    SSB_FC(fdif, ground_dn_dir) = 0.0;
    SSB_FC(fdif, ground_dn) = dn_f;
    SSB_FC(fdif, ground_net) = dn_f - ua_sum_f;
    SSB_FC(fdif, ground_vertical_diff) = vt_f;
  }
  (void)NF;
}

// ===========================================================================
// Longwave
// ===========================================================================
template <int NREG, int NS, bool URBAN>
struct LwFused {
  typedef LwSweepLayout<NREG, NS, URBAN> L;
  static constexpr int n = NREG * NS, d = NREG;
  static constexpr int oState = L::oGeo + kGeoElems;  // a_above (n x n), source_above (n)
  static constexpr int private_elems = oState + n * n + n;
  static constexpr int NF = 4;  // clear_air_abs, veg_air_abs, veg_abs, wall_in
  static constexpr int mFx = 0;                // NF x n
  static constexpr int mF0 = mFx + NF * n;     // NF constants (internal emission pass) + emitted wall power
  static constexpr int mP = mF0 + NF + 1;      // n x n
  static constexpr int mP0 = mP + n * n;       // n
  static constexpr int mScal = mP0 + n;        // segment
  static constexpr int mProf = mScal + 1;      // 2 x (n + 1)
  static constexpr int op_elems = mProf + 2 * (n + 1);
  static constexpr int sAa = 0, sSa = n * n, sT = n * n + n;  // [a_above | source_above] = the carried state
  static constexpr int step_doubles = 2 * n * n + n;
  static constexpr int ls = LayerStack<NREG, NS>::lw_doubles;
  static constexpr int smem_doubles = ls > step_doubles ? ls : step_doubles;
};

template <int NREG, int NS, bool URBAN, bool REC = false>
SSB_HD inline void fused_up_step_lw(const ClassArgs &a, const Tile &Lp, const Scr &Mo, int jl, int il, int il1,
                                    int nlay, int g, const StateMem &sm) {
  const int ls = layer_step(a);
  typedef LwFused<NREG, NS, URBAN> F;
  typedef LwSweepLayout<NREG, NS, URBAN> Lay;
  constexpr int n = NREG * NS, d = NREG, NRB = URBAN ? NREG + 1 : NREG, NF = F::NF;
  const int nspec = a.cfg.nspec;
  const int seg = (int)Lp.ld(Lay::oGeo + 7, 0);
  const SegKeep sk = seg_keep(seg);
  if (!REC) {
    SSB_UNROLL
    for (int i = 0; i < n * n; ++i) sm(F::sAa + i) = Lp.ld(F::oState + i, 0);
    SSB_UNROLL
    for (int i = 0; i < n; ++i) sm(F::sSa + i) = Lp.ld(F::oState + n * n + i, 0);
  }
  SSB_UNROLL
  for (int j = 0; j < n; ++j) {
    SSB_UNROLL
    for (int i = 0; i < n; ++i) sm(F::sT + i + n * j) = Lp.ldp(Lay::oT + i + n * j, 0, sk.k[seg_class<0, NS>(i, j)]);
  }
  double LU[n * n];
  SSB_UNROLL
  for (int j = 0; j < n; ++j) {
    SSB_UNROLL
    for (int i = 0; i < n; ++i) LU[i + n * j] = (i == j) ? 1.0 : 0.0;
    SSB_UNROLL
    for (int k = 0; k < n; ++k) {
      const double r = Lp.ldp(Lay::oR + k + n * j, 0, sk.k[seg_class<0, NS>(k, j)]);
      SSB_UNROLL
      for (int i = 0; i < n; ++i) LU[i + n * j] = fma(-sm(F::sAa + i + n * k), r, LU[i + n * j]);
    }
  }
  sm_lu<n>(LU);
  double W[NF * n], f0[NF + 1];
  const double dz = a.cp.dz[il];
  {
    double od[3], fw[3];
    SSB_UNROLL
    for (int r = 0; r < 3; ++r) {
      fw[r] = Lp.ld(Lay::oGeo + r, 0);
      od[r] = Lp.ld(Lay::oGeo + 3 + r, 0);
    }
    const bool veg = NREG > 1 || !URBAN;
    const double ve = (veg && a.cp.veg_ext) ? a.cp.veg_ext[il] : 0.0;
    const double air_abs = SSB_LAY(a.lw.air_ext, g, il) * (1.0 - SSB_LAY(a.lw.air_ssa, g, il));
    const double vabs = NREG > 1 ? ve * (1.0 - SSB_LAY(a.lw.veg_ssa, g, il)) : 0.0;
    double wt[NF * n];
    SSB_UNROLL
    for (int i = 0; i < n; ++i) {
      const int r = i / NS, js = i % NS;
      const double rmu = 1.0 / a.lg.mu[js];
      wt[0 + NF * i] = (r == 0) ? air_abs * rmu : 0.0;
      wt[1 + NF * i] = (r > 0) ? air_abs * rmu : 0.0;
      wt[2 + NF * i] = (r > 0) ? vabs * od[r] * rmu : 0.0;
      wt[3 + NF * i] = URBAN ? fw[r] * a.lg.tan_ang[js] : 0.0;
    }
    SSB_UNROLL
    for (int i = 0; i < NF * n; ++i) W[i] = 0.0;
    SSB_UNROLL
    for (int j = 0; j < n; ++j) {
      SSB_UNROLL
      for (int i = 0; i < n; ++i) {
        const double v = Lp.ldp(Lay::oIF + i + n * j, 0, sk.k[seg_class<0, NS>(i, j)]);
        SSB_UNROLL
        for (int f = 0; f < NF; ++f) W[f + NF * j] = fma(wt[f + NF * i], v, W[f + NF * j]);
      }
    }
    // constants: weights x int_flux_source minus the emitted power (urban_lw:676-698)
    SSB_UNROLL
    for (int f = 0; f < NF; ++f) f0[f] = 0.0;
    SSB_UNROLL
    for (int i = 0; i < n; ++i) {
      const double v = Lp.ldp(Lay::oIsrc + i, 0, sk.k[seg_class<3, NS>(i, 0)]);
      SSB_UNROLL
      for (int f = 0; f < NF; ++f) f0[f] = fma(wt[f + NF * i], v, f0[f]);
    }
    f0[0] -= Lp.ld(Lay::oBook + 0, 0) * dz;
    SSB_UNROLL
    for (int r = 1; r < NREG; ++r) {
      f0[1] -= Lp.ld(Lay::oBook + d + r, 0) * dz;
      f0[2] -= Lp.ld(Lay::oBook + 2 * d + r, 0) * dz;
    }
    f0[NF] = Lp.ld(Lay::oBook + 3 * d, 0) * dz;
  }
  const bool prof = a.save_profile != 0;
  {
    // the emission column: p0 = D^-1 (R s_above + src), source_below = src + T D^-1 (s_above + a_above src)
    double src[n], p0[n], w0[n];
    SSB_UNROLL
    for (int i = 0; i < n; ++i) {
      src[i] = Lp.ldp(Lay::oSrc + i, 0, sk.k[seg_class<3, NS>(i, 0)]);
      p0[i] = src[i];
      w0[i] = sm(F::sSa + i);
    }
    SSB_UNROLL
    for (int k = 0; k < n; ++k) {
      const double sak = sm(F::sSa + k);
      SSB_UNROLL
      for (int i = 0; i < n; ++i) {
        w0[i] = fma(sm(F::sAa + i + n * k), src[k], w0[i]);
        const double r = Lp.ldp(Lay::oR + i + n * k, 0, sk.k[seg_class<0, NS>(i, k)]);
        p0[i] = fma(r, sak, p0[i]);
      }
    }
    sm_lu_solve_left<n, 1>(LU, p0);
    sm_lu_solve_left<n, 1>(LU, w0);
    double sb[n], ua0[n];
    SSB_UNROLL
    for (int i = 0; i < n; ++i) {
      sb[i] = src[i];
      ua0[i] = sm(F::sSa + i);
    }
    SSB_UNROLL
    for (int k = 0; k < n; ++k) {
      SSB_UNROLL
      for (int i = 0; i < n; ++i) {
        sb[i] = fma(sm(F::sT + i + n * k), w0[k], sb[i]);
        ua0[i] = fma(sm(F::sAa + i + n * k), p0[k], ua0[i]);
      }
    }
    double ssb_ = 0.0, sua = 0.0;
    SSB_UNROLL
    for (int i = 0; i < n; ++i) {
      Lp.st(Lay::oSrc + i, 0, sb[i]);  // source_below over the layer source
      Mo.st(F::mP0 + i, jl, p0[i]);
      ssb_ += sb[i];
      sua += ua0[i];
      SSB_UNROLL
      for (int f = 0; f < NF; ++f) f0[f] = fma(W[f + NF * i], ua0[i], f0[f]);
    }
    SSB_UNROLL
    for (int f = 0; f < NF + 1; ++f) Mo.st(F::mF0 + f, jl, f0[f]);
    if (prof) {
      Mo.st(F::mProf + n, jl, ssb_);
      Mo.st(F::mProf + (n + 1) + n, jl, sua);
    }
  }
  SSB_REC_LOOP
  for (int j = 0; j < n; ++j) {
    const int rj = j / NS;
    if (!region_solved(seg, rj)) {
      SSB_UNROLL
      for (int i = 0; i < n; ++i) Lp.st(Lay::oR + i + n * j, 0, 0.0);
      continue;
    }
    double p[n], x[n];
    SSB_UNROLL
    for (int i = 0; i < n; ++i) {
      p[i] = sm(F::sT + i + n * j);
      x[i] = 0.0;
    }
    SSB_UNROLL
    for (int k = 0; k < n; ++k) {
      SSB_UNROLL
      for (int i = 0; i < n; ++i) x[i] = fma(sm(F::sAa + i + n * k), p[k], x[i]);
    }
    sm_lu_solve_left<n, 1>(LU, p);
    sm_lu_solve_left<n, 1>(LU, x);
    double ab[n], ap[n];
    SSB_UNROLL
    for (int i = 0; i < n; ++i) {
      ab[i] = Lp.ldp(Lay::oR + i + n * j, 0, keep_rc(seg, i / NS, rj));
      ap[i] = 0.0;
    }
    SSB_UNROLL
    for (int k = 0; k < n; ++k) {
      SSB_UNROLL
      for (int i = 0; i < n; ++i) {
        ab[i] = fma(sm(F::sT + i + n * k), x[k], ab[i]);
        ap[i] = fma(sm(F::sAa + i + n * k), p[k], ap[i]);
      }
    }
    double sab = 0.0, sap = 0.0, fx[NF];
    SSB_UNROLL
    for (int f = 0; f < NF; ++f) fx[f] = 0.0;
    SSB_UNROLL
    for (int i = 0; i < n; ++i) {
      Lp.st(Lay::oR + i + n * j, 0, ab[i]);  // a_below over R
      Mo.st(F::mP + i + n * j, jl, p[i]);
      sab += ab[i];
      sap += ap[i];
      const double c = ((i == j) ? 1.0 : 0.0) + ap[i];  // int_flux acts on x_below + up_above
      SSB_UNROLL
      for (int f = 0; f < NF; ++f) fx[f] = fma(W[f + NF * i], c, fx[f]);
    }
    SSB_UNROLL
    for (int f = 0; f < NF; ++f) Mo.st(F::mFx + f + NF * j, jl, fx[f]);
    if (prof) {
      Mo.st(F::mProf + j, jl, sab);
      Mo.st(F::mProf + (n + 1) + j, jl, sap);
    }
  }
  Mo.st(F::mScal, jl, (double)seg);
  // ---- roofs, overlap ---------------------------------------------------------------------------
  double hw[NS], rb[NS], rs[NS];
  SSB_UNROLL
  for (int js = 0; js < NS; ++js) {
    hw[js] = a.lg.hweight[js];
    rb[js] = rs[js] = 0.0;
  }
  if (URBAN) {
    const double bfj = a.cp.building_fraction[il];
    const double exposed = (jl < nlay - 1) ? dmax(0.0, bfj - a.cp.building_fraction[il + ls]) : bfj;
    const double remis = SSB_LAY(a.lw.roof_emissivity, g, il), remission = SSB_LAY(a.lw.roof_emission, g, il);
    SSB_UNROLL
    for (int js = 0; js < NS; ++js) {
      rb[js] = (1.0 - remis) * hw[js];
      rs[js] = hw[js] * remission * exposed;
    }
  }
  double U[12], V[12];
  overlap_above<NREG, URBAN, true>(a, il1, nlay, jl, U, V);
  const StateMem stt = REC ? StateMem{&sm(F::sAa), sm.stride}
                           : StateMem{Lp.base + (size_t)F::oState * kScratchTile, kScratchTile};
  {
    double Ab[n * n];
    SSB_UNROLL
    for (int i = 0; i < n * n; ++i) Ab[i] = Lp.ld(Lay::oR + i, 0);
    overlap_matrix<NREG, NRB, NS>(Ab, rb, U, V, stt, 0);
  }
  {
    double Sb[n];
    SSB_UNROLL
    for (int i = 0; i < n; ++i) Sb[i] = Lp.ld(Lay::oSrc + i, 0);
    SSB_UNROLL
    for (int u = 0; u < NREG; ++u) {
      SSB_UNROLL
      for (int jt = 0; jt < NS; ++jt) {
        double s = 0.0;
        SSB_UNROLL
        for (int lo = 0; lo < NREG; ++lo) s = fma(U[u + NREG * lo], Sb[lo * NS + jt], s);
        if (URBAN) s = fma(U[u + NREG * NREG], rs[jt], s);
        stt(n * n + u * NS + jt) = s;
      }
    }
  }
  (void)nspec;
}

template <int NREG, int NS>
SSB_HD inline void fused_layer_lw(const ClassArgs &a, int q, int lev, const StateMem &st) {
  if (fast_prepare_level(a, q, lev) < 0) return;
  fast_layer_problem_lw<NREG, NS>(a, q, lev, st);
}

template <int NREG, int NS, bool URBAN, int MODE = 0>
SSB_HD inline void fused_column_lw(const ClassArgs &a, int q_in, bool active, const StateMem &st) {
  constexpr bool REC = MODE != 0;
  typedef LwFused<NREG, NS, URBAN> F;
  constexpr int n = NREG * NS, NRB = URBAN ? NREG + 1 : NREG, NF = F::NF;
  const SolveCfg &c = a.cfg;
  const int nspec = c.nspec;
  const int q = active ? q_in : 0;
  const int ic = q / nspec, g = q % nspec;
  const int col = a.cols[ic];
  const int nlay = a.nlay[col], il1 = layer_index(a, ic, col, 0), ls = layer_step(a);
  const ssb200_canopy_flux &fint = a.f1, &fnorm = a.f2;
  if (!active) {
    if (!REC && (a.fused & 2))
      for (int jl = 0; jl < a.lmax; ++jl) {
        phase_sync(a);
        phase_sync(a);
      }
    return;
  }
  if (MODE != 2) {
    zero_unwritten_lw<NREG, URBAN>(fint, nspec, g, col, il1, nlay, ls);
    zero_unwritten_lw<NREG, URBAN>(fnorm, nspec, g, col, il1, nlay, ls);
  }
  double hw[NS], tang[NS];
  SSB_UNROLL
  for (int js = 0; js < NS; ++js) {
    hw[js] = a.lg.hweight[js];
    tang[js] = a.lg.tan_ang[js];
  }
  const Tile Lp(a, q);
  const Scr Mo(a.sweep, a.lmax, a.ne_sweep, q);
  const double gemis = a.lw.ground_emissivity[(size_t)g + (size_t)nspec * col];
  const double gemission = a.lw.ground_emission[(size_t)g + (size_t)nspec * col];
  double frac0[3] = {1.0, 0.0, 0.0};
  if (nlay > 0) {
    const bool veg = NREG > 1 || !URBAN;
    region_fractions(c, URBAN ? a.cp.building_fraction[il1] : 0.0,
                     (veg && a.cp.veg_fraction) ? a.cp.veg_fraction[il1] : 0.0, frac0);
  }
  double top_emissivity = 0.0, top_emission = 0.0;
  if (MODE != 2) {
    auto state_st = [&](int e, double v) {
      if (REC)
        st(F::sAa + e) = v;
      else
        Lp.st(F::oState + e, 0, v);
    };
    auto state_ld = [&](int e) -> double { return REC ? st(F::sAa + e) : Lp.ld(F::oState + e, 0); };
    SSB_UNROLL
    for (int i = 0; i < n * n + n; ++i) state_st(i, 0.0);
    SSB_UNROLL
    for (int r = 0; r < NREG; ++r) {
      SSB_UNROLL
      for (int jt = 0; jt < NS; ++jt) {
        SSB_UNROLL
        for (int jf = 0; jf < NS; ++jf) state_st((jt + r * NS) + n * (jf + r * NS), (1.0 - gemis) * hw[jt]);
        state_st(n * n + jt + r * NS, (hw[jt] * frac0[r]) * gemission);
      }
    }
    if (REC) {
      for (int jl = 0; jl < nlay; ++jl)
        fused_up_step_lw<NREG, NS, URBAN, REC>(a, Lp.at(jl), Mo, jl, il1 + jl * ls, il1, nlay, g, st);
    } else {
      const int nloop = (a.fused & 2) ? a.lmax : nlay;
      for (int jl = 0; jl < nloop; ++jl) {
        phase_sync(a);
        if (jl < nlay) fused_layer_lw<NREG, NS>(a, q, jl, st);
        phase_sync(a);
        if (jl < nlay) fused_up_step_lw<NREG, NS, URBAN>(a, Lp, Mo, jl, il1 + jl * ls, il1, nlay, g, st);
      }
    }
    double sAll = 0.0;
    SSB_UNROLL
    for (int i = 0; i < NS; ++i) {
      double s = 0.0;
      SSB_UNROLL
      for (int j = 0; j < NS; ++j) s = fma(state_ld(i + n * j), hw[j], s);
      sAll += s;
      top_emission += state_ld(n * n + i);
    }
    top_emissivity = 1.0 - sAll;
    a.bc.lw_emissivity[(size_t)g + (size_t)nspec * col] = top_emissivity;
    a.bc.lw_emission[(size_t)g + (size_t)nspec * col] = top_emission;
  }
  if (MODE == 1) return;
  if (MODE == 2) {
    top_emissivity = a.bc.lw_emissivity[(size_t)g + (size_t)nspec * col];
    top_emission = a.bc.lw_emission[(size_t)g + (size_t)nspec * col];
  }
  // ---- downward: internal emission (suffix i) and incoming flux (suffix f) through the records ---
  double xa_i[n], xa_f[n];
  SSB_UNROLL
  for (int i = 0; i < n; ++i) xa_i[i] = xa_f[i] = 0.0;
  SSB_UNROLL
  for (int js = 0; js < NS; ++js) xa_f[js] = hw[js];
  SSB_FC(fint, top_dn) = 0.0;
  SSB_FC(fint, top_net) = -top_emission;
  SSB_FC(fnorm, top_dn) = 1.0;
  SSB_FC(fnorm, top_net) = top_emissivity;
  const bool prof = a.save_profile != 0;
  for (int jl = nlay - 1; jl >= 0; --jl) {
    const int il = il1 + jl * ls;
    const int seg = (int)Mo.ld(F::mScal, jl);
    double xb_i[NRB * NS], xb_f[NRB * NS];
    {
      double U[12], V[12];
      overlap_above<NREG, URBAN, true>(a, il1, nlay, jl, U, V);
      expand_down<NREG, NRB, NS>(V, xa_i, xb_i);
      expand_down<NREG, NRB, NS>(V, xa_f, xb_f);
    }
    double fn_i[NF], fn_f[NF], sp_i[2], sp_f[2] = {0.0, 0.0};
    SSB_UNROLL
    for (int f = 0; f < NF; ++f) {
      fn_i[f] = Mo.ld(F::mF0 + f, jl);
      fn_f[f] = 0.0;
    }
    sp_i[0] = prof ? Mo.ld(F::mProf + n, jl) : 0.0;
    sp_i[1] = prof ? Mo.ld(F::mProf + (n + 1) + n, jl) : 0.0;
    SSB_UNROLL
    for (int i = 0; i < n; ++i) {
      xa_i[i] = Mo.ld(F::mP0 + i, jl);
      xa_f[i] = 0.0;
    }
    SSB_UNROLL
    for (int j = 0; j < n; ++j) {
      if (region_solved(seg, j / NS)) {
        SSB_UNROLL
        for (int i = 0; i < n; ++i) {
          const double p = Mo.ld(F::mP + i + n * j, jl);
          xa_i[i] = fma(p, xb_i[j], xa_i[i]);
          xa_f[i] = fma(p, xb_f[j], xa_f[i]);
        }
        SSB_UNROLL
        for (int f = 0; f < NF; ++f) {
          const double w = Mo.ld(F::mFx + f + NF * j, jl);
          fn_i[f] = fma(w, xb_i[j], fn_i[f]);
          fn_f[f] = fma(w, xb_f[j], fn_f[f]);
        }
        if (prof) {
          const double s0 = Mo.ld(F::mProf + j, jl), s1 = Mo.ld(F::mProf + (n + 1) + j, jl);
          sp_i[0] = fma(s0, xb_i[j], sp_i[0]);
          sp_f[0] = fma(s0, xb_f[j], sp_f[0]);
          sp_i[1] = fma(s1, xb_i[j], sp_i[1]);
          sp_f[1] = fma(s1, xb_f[j], sp_f[1]);
        }
      }
    }
    if (URBAN) {
      const double bfj = a.cp.building_fraction[il];
      const double exposed = (jl < nlay - 1) ? dmax(0.0, bfj - a.cp.building_fraction[il + ls]) : bfj;
      const double remis = SSB_LAY(a.lw.roof_emissivity, g, il), remission = SSB_LAY(a.lw.roof_emission, g, il);
      double sroof_i = 0.0, sroof_f = 0.0, rup_i = 0.0, rup_f = 0.0;
      SSB_UNROLL
      for (int js = 0; js < NS; ++js) {
        sroof_i += xb_i[n + js];
        sroof_f += xb_f[n + js];
      }
      SSB_UNROLL
      for (int js = 0; js < NS; ++js) {
        rup_i += (1.0 - remis) * hw[js] * sroof_i + hw[js] * remission * exposed;
        rup_f += (1.0 - remis) * hw[js] * sroof_f;
      }
      SSB_FL(fint, roof_in, il) = sroof_i;
      SSB_FL(fint, roof_net, il) = sroof_i - rup_i;
      SSB_FL(fnorm, roof_in, il) = sroof_f;
      SSB_FL(fnorm, roof_net, il) = sroof_f - rup_f;
    }
    if (fint.flux_dn_layer_top || fnorm.flux_dn_layer_top) {
      double s0 = 0.0, s1 = 0.0, s4 = 0.0, s5 = 0.0;
      SSB_UNROLL
      for (int i = 0; i < n; ++i) {
        s0 += xb_i[i];
        s1 += xa_i[i];
        s4 += xb_f[i];
        s5 += xa_f[i];
      }
      if (fint.flux_dn_layer_top) {
        SSB_FL(fint, flux_dn_layer_top, il) = s0;
        SSB_FL(fint, flux_dn_layer_base, il) = s1;
        SSB_FL(fint, flux_up_layer_top, il) = sp_i[0];
        SSB_FL(fint, flux_up_layer_base, il) = sp_i[1];
      }
      if (fnorm.flux_dn_layer_top) {
        SSB_FL(fnorm, flux_dn_layer_top, il) = s4;
        SSB_FL(fnorm, flux_dn_layer_base, il) = s5;
        SSB_FL(fnorm, flux_up_layer_top, il) = sp_f[0];
        SSB_FL(fnorm, flux_up_layer_base, il) = sp_f[1];
      }
    }
    SSB_FL(fint, clear_air_abs, il) = fn_i[0];
    SSB_FL(fnorm, clear_air_abs, il) = fn_f[0];
    if (NREG > 1) {
      SSB_FL(fint, veg_air_abs, il) = fn_i[1];
      SSB_FL(fint, veg_abs, il) = fn_i[2];
      SSB_FL(fnorm, veg_air_abs, il) = fn_f[1];
      SSB_FL(fnorm, veg_abs, il) = fn_f[2];
    }
    if (URBAN) {
      const double wemis = SSB_LAY(a.lw.wall_emissivity, g, il);
      SSB_FL(fint, wall_in, il) = fn_i[3];
      SSB_FL(fint, wall_net, il) = fn_i[3] * wemis - Mo.ld(F::mF0 + NF, jl);
      SSB_FL(fnorm, wall_in, il) = fn_f[3];
      SSB_FL(fnorm, wall_net, il) = fn_f[3] * wemis;
    }
  }
  {
    // at the ground: up_above = (1 - emissivity) hweight sum(x_above) + emission (urban_lw:554-565)
    double dn_i = 0.0, up_i = 0.0, vt_i = 0.0, dn_f = 0.0, up_f = 0.0, vt_f = 0.0;
    SSB_UNROLL
    for (int r = 0; r < NREG; ++r) {
      double si = 0.0, sf = 0.0;
      SSB_UNROLL
      for (int js = 0; js < NS; ++js) {
        si += xa_i[r * NS + js];
        sf += xa_f[r * NS + js];
      }
      SSB_UNROLL
      for (int js = 0; js < NS; ++js) {
        const int i = js + r * NS;
        double ui = 0.0, uf = 0.0;
        if (nlay > 0) {
          ui = fma((1.0 - gemis) * hw[js], si, (hw[js] * frac0[r]) * gemission);
          uf = (1.0 - gemis) * hw[js] * sf;
        }
        dn_i += xa_i[i];
        up_i += ui;
        vt_i += (xa_i[i] + ui) * tang[js] / SSB_PI;
        dn_f += xa_f[i];
        up_f += uf;
        vt_f += (xa_f[i] + uf) * tang[js] / SSB_PI;
      }
    }
    SSB_FC(fint, ground_dn) = dn_i;
    SSB_FC(fint, ground_net) = dn_i - up_i;
    SSB_FC(fnorm, ground_dn) = dn_f;
    SSB_FC(fnorm, ground_net) = dn_f - up_f;
    // forest_lw:687-694 accumulates the normalised pass into lw_internal as well
    if (URBAN) {
      SSB_FC(fint, ground_vertical_diff) = vt_i;
      SSB_FC(fnorm, ground_vertical_diff) = vt_f;
    } else {
      SSB_FC(fint, ground_vertical_diff) = vt_i + vt_f;
      SSB_FC(fnorm, ground_vertical_diff) = 0.0;
    }
  }
}

}  // namespace ssb
