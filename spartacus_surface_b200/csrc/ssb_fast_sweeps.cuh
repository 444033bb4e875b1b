// ssb_fast_sweeps.cuh - register-resident adding sweeps: one thread per
// (column, interval) with compile-time (NREG regions, NS streams, URBAN).
//
// Same recurrences and outputs as column_sweeps_sw/_lw of ssb_solver.cuh
// (radsurf_urban_sw.F90:591-984, radsurf_urban_lw.F90:552-858 and the forest
// equivalents) with two changes that cut the inter-sweep state:
//  * only a_above, d_above (or source_above) and the LU factors of the
//    denominator I - a_above R are kept per interface; the "below" albedo
//    matrices (which carry the extra roof region) are never formed in the
//    downward passes - their action on a flux vector is evaluated as
//      a_below x = R x + T D^-1 (a_above (T x))   (+ the roof block),
//    and likewise for d_below and source_below;
//  * the overlap matrices U, V are recomputed from the region fractions of the
//    two adjacent layers instead of being stored.
#pragma once
#include "ssb_fast_layer.cuh"

namespace ssb {

template <int N, int C>
SSB_HDI void sload(const double *S, int e0, int lev, int nlev, int width, int q, double *dst) {
  SSB_UNROLL
  for (int i = 0; i < N * C; ++i) dst[i] = S[sidx(e0 + i, lev, nlev, width, q)];
}
template <int N, int C>
SSB_HDI void sstore(double *S, int e0, int lev, int nlev, int width, int q, const double *src) {
  SSB_UNROLL
  for (int i = 0; i < N * C; ++i) S[sidx(e0 + i, lev, nlev, width, q)] = src[i];
}
// y += A x with A (R x C) read straight from scratch
template <int R, int C>
SSB_HDI void smv_acc(const double *S, int e0, int lev, int nlev, int width, int q, const double *x, double *y) {
  SSB_UNROLL
  for (int j = 0; j < C; ++j) {
    SSB_UNROLL
    for (int i = 0; i < R; ++i) y[i] = fma(S[sidx(e0 + i + R * j, lev, nlev, width, q)], x[j], y[i]);
  }
}

// (V (x) I_NS) x : below-interface vector (NRB*NS) from the above-interface one (NREG*NS)
template <int NREG, int NRB, int NS>
SSB_HDI void expand_down(const double *V, const double *x, double *y) {
  SSB_UNROLL
  for (int lo = 0; lo < NRB; ++lo) {
    SSB_UNROLL
    for (int js = 0; js < NS; ++js) {
      double s = 0.0;
      SSB_UNROLL
      for (int up = 0; up < NREG; ++up) s = fma(V[lo + NRB * up], x[up * NS + js], s);
      y[lo * NS + js] = s;
    }
  }
}

template <int NREG, int NS, bool URBAN>
SSB_HD inline void fast_column_sweeps_sw(const ClassArgs &a, int q) {
  constexpr int n = NREG * NS, d = NREG, NRB = URBAN ? NREG + 1 : NREG, m = NRB * NS;
  const SolveCfg &c = a.cfg;
  const int nspec = c.nspec;
  const int ic = q / nspec, g = q % nspec;
  const int col = a.cols[ic];
  const int nlay = a.nlay[col], il1 = a.istartlay[col] - 1;
  const int width = a.ncols * nspec;
  const ssb200_canopy_flux &fdir = a.f1, &fdif = a.f2;
  const double cos_sza = a.cp.cos_sza[col];
  int itransp = 0;
  if (nspec > 1) {
    double best = 0.0;
    for (int gg = 0; gg < nspec; ++gg) {
      double od = 0.0;
      for (int l = 0; l < nlay; ++l) od += a.sw.air_ext[(size_t)gg + (size_t)nspec * (il1 + l)] * a.cp.dz[il1 + l];
      if (gg == 0 || od < best) {
        best = od;
        itransp = gg;
      }
    }
  }
  const bool own = (g == itransp);
  zero_column(fdir, nspec, g, col, il1, nlay, own);
  zero_column(fdif, nspec, g, col, il1, nlay, own);
  if (!(cos_sza > 0.0)) return;
  const double zcos = URBAN ? dmax(cos_sza, 1.0e-6) : cos_sza;
  const double sin0 = URBAN ? sqrt(1.0 - zcos * zcos) : 0.0;
  const double galb = a.sw.ground_albedo[(size_t)g + (size_t)nspec * col];
  const double galb_dir =
      (a.use_sw_direct_albedo ? a.sw.ground_albedo_dir : a.sw.ground_albedo)[(size_t)g + (size_t)nspec * col];
  double hw[NS], mu_inv[NS], tang[NS];
  SSB_UNROLL
  for (int js = 0; js < NS; ++js) {
    hw[js] = a.lg.hweight[js];
    mu_inv[js] = 1.0 / a.lg.mu[js];
    tang[js] = a.lg.tan_ang[js];
  }
  const double *LS = a.layer;
  double *WS = a.sweep;
  const int nlev = a.lmax, wlev = a.lmax + 1;
  constexpr int oR = 0, oT = n * n, oIdiff = 2 * n * n, oSup = 3 * n * n, oSdn = oSup + n * d, oIdd = oSdn + n * d,
                oE = oIdd + n * d, oIdir = oE + d * d;
  constexpr int oAa = 0, oDa = n * n, oLU = oDa + n * d;  // fast-path sweep scratch

  double Aa[n * n], Da[n * d];
  // ---- upward sweep ---------------------------------------------------------
  SSB_UNROLL
  for (int i = 0; i < n * n; ++i) Aa[i] = 0.0;
  SSB_UNROLL
  for (int i = 0; i < n * d; ++i) Da[i] = 0.0;
  SSB_UNROLL
  for (int r = 0; r < NREG; ++r) {
    SSB_UNROLL
    for (int jt = 0; jt < NS; ++jt) {
      Da[(jt + r * NS) + n * r] = zcos * galb_dir * hw[jt];
      SSB_UNROLL
      for (int jf = 0; jf < NS; ++jf) Aa[(jt + r * NS) + n * (jf + r * NS)] = galb * hw[jt];
    }
  }
  sstore<n, n>(WS, oAa, 0, wlev, width, q, Aa);
  sstore<n, d>(WS, oDa, 0, wlev, width, q, Da);
  for (int jl = 0; jl < nlay; ++jl) {
    const int il = il1 + jl;
    double R[n * n], T[n * n], LU[n * n], X[n * n];
    sload<n, n>(LS, oR, jl, nlev, width, q, R);
    sload<n, n>(LS, oT, jl, nlev, width, q, T);
    sm_mul<n, n, n>(Aa, R, LU);
    SSB_UNROLL
    for (int j = 0; j < n; ++j) {
      SSB_UNROLL
      for (int i = 0; i < n; ++i) LU[i + n * j] = (i == j ? 1.0 : 0.0) - LU[i + n * j];
    }
    sm_lu<n>(LU);
    sstore<n, n>(WS, oLU, jl, wlev, width, q, LU);
    sm_mul<n, n, n>(Aa, T, X);
    sm_lu_solve_left<n, n>(LU, X);
    double Ab[n * n];
    sm_mul<n, n, n>(T, X, Ab);
    SSB_UNROLL
    for (int i = 0; i < n * n; ++i) Ab[i] += R[i];
    // d_below (street part)
    double Db[n * d];
    {
      double Su[n * d], Sd[n * d], E[d * d], W1[n * d], W2[n * d];
      sload<n, d>(LS, oSup, jl, nlev, width, q, Su);
      sload<n, d>(LS, oSdn, jl, nlev, width, q, Sd);
      sload<d, d>(LS, oE, jl, nlev, width, q, E);
      sm_mul<n, d, d>(Da, E, W1);
      sm_mul<n, n, d>(Aa, Sd, W2);
      SSB_UNROLL
      for (int i = 0; i < n * d; ++i) W1[i] += W2[i];
      sm_lu_solve_left<n, d>(LU, W1);
      sm_mul<n, n, d>(T, W1, Db);
      SSB_UNROLL
      for (int i = 0; i < n * d; ++i) Db[i] += Su[i];
    }
    double rb[NS], rd[NS];  // roof rows of a_below / d_below
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
    overlap_at(a, il1, nlay, jl + 1, U, V);
    // a_above(next) = (U (x) I) a_below (V (x) I), d_above(next) = (U (x) I) d_below V
    {
      double AV[n * n];  // street rows of a_below (V (x) I)
      SSB_UNROLL
      for (int up = 0; up < NREG; ++up) {
        SSB_UNROLL
        for (int js = 0; js < NS; ++js) {
          SSB_UNROLL
          for (int i = 0; i < n; ++i) {
            double s = 0.0;
            SSB_UNROLL
            for (int lo = 0; lo < NREG; ++lo) s = fma(Ab[i + n * (lo * NS + js)], V[lo + NRB * up], s);
            AV[i + n * (up * NS + js)] = s;
          }
        }
      }
      SSB_UNROLL
      for (int up = 0; up < NREG; ++up) {
        SSB_UNROLL
        for (int js = 0; js < NS; ++js) {
          SSB_UNROLL
          for (int u = 0; u < NREG; ++u) {
            SSB_UNROLL
            for (int jt = 0; jt < NS; ++jt) {
              double s = 0.0;
              SSB_UNROLL
              for (int lo = 0; lo < NREG; ++lo) s = fma(U[u + NREG * lo], AV[(lo * NS + jt) + n * (up * NS + js)], s);
              if (URBAN) s = fma(U[u + NREG * NREG] * rb[jt], V[NREG + NRB * up], s);
              Aa[(u * NS + jt) + n * (up * NS + js)] = s;
            }
          }
        }
      }
      double DV[n * d];
      SSB_UNROLL
      for (int up = 0; up < NREG; ++up) {
        SSB_UNROLL
        for (int i = 0; i < n; ++i) {
          double s = 0.0;
          SSB_UNROLL
          for (int lo = 0; lo < NREG; ++lo) s = fma(Db[i + n * lo], V[lo + NRB * up], s);
          DV[i + n * up] = s;
        }
      }
      SSB_UNROLL
      for (int up = 0; up < NREG; ++up) {
        SSB_UNROLL
        for (int u = 0; u < NREG; ++u) {
          SSB_UNROLL
          for (int jt = 0; jt < NS; ++jt) {
            double s = 0.0;
            SSB_UNROLL
            for (int lo = 0; lo < NREG; ++lo) s = fma(U[u + NREG * lo], DV[(lo * NS + jt) + n * up], s);
            if (URBAN) s = fma(U[u + NREG * NREG] * rd[jt], V[NREG + NRB * up], s);
            Da[(u * NS + jt) + n * up] = s;
          }
        }
      }
    }
    sstore<n, n>(WS, oAa, jl + 1, wlev, width, q, Aa);
    sstore<n, d>(WS, oDa, jl + 1, wlev, width, q, Da);
  }
  double talb_diff = 0.0, talb_dir = 0.0;
  {
    SSB_UNROLL
    for (int i = 0; i < NS; ++i) {
      double s = 0.0;
      SSB_UNROLL
      for (int j = 0; j < NS; ++j) s = fma(Aa[i + n * j], hw[j], s);
      talb_diff += s;
    }
    double s = 0.0;
    SSB_UNROLL
    for (int js = 0; js < NS; ++js) s += Da[js];
    talb_dir = s / zcos;
    a.bc.sw_albedo[(size_t)g + (size_t)nspec * col] = talb_diff;
    a.bc.sw_albedo_dir[(size_t)g + (size_t)nspec * col] = talb_dir;
  }

  // ---- two downward passes ----------------------------------------------------
  for (int pass = 0; pass < 2; ++pass) {
    const bool direct = (pass == 0);
    const ssb200_canopy_flux &f = direct ? fdir : fdif;
    double dir_above[d], diff_above[n], up_above[n];
    SSB_UNROLL
    for (int i = 0; i < d; ++i) dir_above[i] = 0.0;
    SSB_UNROLL
    for (int i = 0; i < n; ++i) {
      diff_above[i] = 0.0;
      up_above[i] = 0.0;
    }
    double flux_dn_dir_clear = 1.0 / zcos;
    if (direct) {
      dir_above[0] = 1.0 / zcos;
      SSB_FC(f, top_dn_dir) = 1.0;
      SSB_FC(f, top_dn) = 1.0;
      SSB_FC(f, top_net) = 1.0 * (1.0 - talb_dir);
      if (URBAN && own && f.roof_sunlit_frac && nlay > 0) f.roof_sunlit_frac[il1 + nlay - 1] = 1.0;
    } else {
      SSB_UNROLL
      for (int js = 0; js < NS; ++js) diff_above[js] = hw[js];
      SSB_FC(f, top_dn_dir) = 0.0;
      SSB_FC(f, top_dn) = 1.0;
      SSB_FC(f, top_net) = 1.0 - talb_diff;
    }
    for (int jl = nlay - 1; jl >= 0; --jl) {
      const int il = il1 + jl;
      LayerGeom gm;
      double bf, vf, ve;
      geometry_of_layer(a, il, 1.0, gm, bf, vf, ve);
      double U[12], V[12];
      overlap_at(a, il1, nlay, jl + 1, U, V);
      double diff_below[m], dir_below[NRB], up_below[m];
      expand_down<NREG, NRB, NS>(V, diff_above, diff_below);
      SSB_UNROLL
      for (int lo = 0; lo < NRB; ++lo) {
        double s = 0.0;
        SSB_UNROLL
        for (int up = 0; up < NREG; ++up) s = fma(V[lo + NRB * up], dir_above[up], s);
        dir_below[lo] = direct ? s : 0.0;
      }
      double LU[n * n];
      sload<n, n>(WS, oLU, jl, wlev, width, q, LU);
      // y = T x + Sdn dirb ; refl = d_above (E dirb)
      double y[n], refl[n], ddir[d];
      SSB_UNROLL
      for (int i = 0; i < n; ++i) {
        y[i] = 0.0;
        refl[i] = 0.0;
      }
      smv_acc<n, n>(LS, oT, jl, nlev, width, q, diff_below, y);
      if (direct) {
        smv_acc<n, d>(LS, oSdn, jl, nlev, width, q, dir_below, y);
        double da_new[d];
        SSB_UNROLL
        for (int i = 0; i < d; ++i) da_new[i] = 0.0;
        smv_acc<d, d>(LS, oE, jl, nlev, width, q, dir_below, da_new);
        SSB_UNROLL
        for (int i = 0; i < d; ++i) {
          ddir[i] = dir_below[i] - da_new[i];
          dir_above[i] = da_new[i];
        }
        smv_acc<n, d>(WS, oDa, jl, wlev, width, q, dir_above, refl);
      } else {
        SSB_UNROLL
        for (int i = 0; i < d; ++i) ddir[i] = 0.0;
      }
      // z1 = D^-1 (A y + refl) for up_below ; diff_above = D^-1 (y + R refl)
      double z1[n], z2[n];
      SSB_UNROLL
      for (int i = 0; i < n; ++i) {
        z1[i] = refl[i];
        z2[i] = y[i];
      }
      smv_acc<n, n>(WS, oAa, jl, wlev, width, q, y, z1);
      if (direct) smv_acc<n, n>(LS, oR, jl, nlev, width, q, refl, z2);
      sm_lu_solve_left<n, 1>(LU, z1);
      sm_lu_solve_left<n, 1>(LU, z2);
      // up_below (street part) = R x + Sup dirb + T z1
      SSB_UNROLL
      for (int i = 0; i < m; ++i) up_below[i] = 0.0;
      smv_acc<n, n>(LS, oR, jl, nlev, width, q, diff_below, up_below);
      smv_acc<n, n>(LS, oT, jl, nlev, width, q, z1, up_below);
      if (direct) smv_acc<n, d>(LS, oSup, jl, nlev, width, q, dir_below, up_below);
      if (URBAN) {
        const double ralb = SSB_LAY(a.sw.roof_albedo, g, il);
        const double ralb_dir = a.sw.roof_albedo_dir ? SSB_LAY(a.sw.roof_albedo_dir, g, il) : ralb;
        double sroof = 0.0;
        SSB_UNROLL
        for (int js = 0; js < NS; ++js) sroof += diff_below[n + js];
        double roof_up = 0.0;
        SSB_UNROLL
        for (int js = 0; js < NS; ++js) {
          up_below[n + js] = ralb * hw[js] * sroof + (direct ? zcos * ralb_dir * hw[js] * dir_below[NREG] : 0.0);
          roof_up += up_below[n + js];
        }
        if (direct) {
          SSB_FL(f, roof_in_dir, il) = zcos * dir_below[NREG];
          SSB_FL(f, roof_in, il) = SSB_FL(f, roof_in_dir, il) + sroof;
        } else {
          SSB_FL(f, roof_in, il) = sroof;
        }
        SSB_FL(f, roof_net, il) = SSB_FL(f, roof_in, il) - roof_up;
      }
      SSB_UNROLL
      for (int i = 0; i < n; ++i) {
        diff_above[i] = z2[i];
        up_above[i] = refl[i];
      }
      smv_acc<n, n>(WS, oAa, jl, wlev, width, q, diff_above, up_above);

      if (f.flux_dn_layer_top) {
        double s_db = 0.0, s_da = 0.0, s_ub = 0.0, s_ua = 0.0, s_dirb = 0.0, s_dira = 0.0;
        SSB_UNROLL
        for (int i = 0; i < n; ++i) {
          s_db += diff_below[i];
          s_da += diff_above[i];
          s_ub += up_below[i];
          s_ua += up_above[i];
        }
        SSB_UNROLL
        for (int i = 0; i < d; ++i) {
          s_dirb += dir_below[i];
          s_dira += dir_above[i];
        }
        if (direct) {
          SSB_FL(f, flux_dn_dir_layer_top, il) = zcos * s_dirb;
          SSB_FL(f, flux_dn_layer_top, il) = zcos * s_dirb + s_db;
          SSB_FL(f, flux_dn_dir_layer_base, il) = zcos * s_dira;
          SSB_FL(f, flux_dn_layer_base, il) = zcos * s_dira + s_da;
        } else {
          SSB_FL(f, flux_dn_layer_top, il) = s_db;
          SSB_FL(f, flux_dn_layer_base, il) = s_da;
        }
        SSB_FL(f, flux_up_layer_top, il) = s_ub;
        SSB_FL(f, flux_up_layer_base, il) = s_ua;
      }
      // integrated fluxes
      double conv[n], iflux_diff[n], iflux_dir[d];
      SSB_UNROLL
      for (int i = 0; i < n; ++i) {
        conv[i] = diff_below[i] - diff_above[i] - up_below[i] + up_above[i];
        iflux_diff[i] = 0.0;
      }
      SSB_UNROLL
      for (int i = 0; i < d; ++i) iflux_dir[i] = 0.0;
      smv_acc<n, n>(LS, oIdiff, jl, nlev, width, q, conv, iflux_diff);
      if (direct) {
        smv_acc<d, d>(LS, oIdir, jl, nlev, width, q, ddir, iflux_dir);
        smv_acc<n, d>(LS, oIdd, jl, nlev, width, q, ddir, iflux_diff);
      }
      double smu[NREG], stan[NREG];
      SSB_UNROLL
      for (int r = 0; r < NREG; ++r) {
        smu[r] = 0.0;
        stan[r] = 0.0;
        SSB_UNROLL
        for (int js = 0; js < NS; ++js) {
          smu[r] = fma(iflux_diff[r * NS + js], mu_inv[js], smu[r]);
          stan[r] = fma(iflux_diff[r * NS + js], tang[js], stan[r]);
        }
      }
      const double air_ext = SSB_LAY(a.sw.air_ext, g, il);
      const double air_abs = air_ext * (1.0 - SSB_LAY(a.sw.air_ssa, g, il));
      SSB_FL(f, clear_air_abs, il) = air_abs * (iflux_dir[0] + smu[0]);
      if (NREG > 1) {
        const double vabs = ve * (1.0 - SSB_LAY(a.sw.veg_ssa, g, il));
        double s_air = 0.0, s_veg = 0.0, s_vdir = 0.0;
        SSB_UNROLL
        for (int r = 1; r < NREG; ++r) {
          s_air += air_abs * (iflux_dir[r] + smu[r]);
          s_vdir += vabs * iflux_dir[r] * gm.od_scaling[r];
          s_veg += vabs * (iflux_dir[r] + smu[r]) * gm.od_scaling[r];
        }
        SSB_FL(f, veg_air_abs, il) = s_air;
        SSB_FL(f, veg_abs, il) = s_veg;
        if (direct) SSB_FL(f, veg_abs_dir, il) = s_vdir;
      }
      if (URBAN) {
        const double walb = SSB_LAY(a.sw.wall_albedo, g, il);
        double win_dir = 0.0, win = 0.0;
        SSB_UNROLL
        for (int r = 0; r < NREG; ++r) {
          win_dir += gm.f_wall[r] * sin0 * iflux_dir[r];
          win += gm.f_wall[r] * stan[r];
        }
        if (direct) SSB_FL(f, wall_in_dir, il) = win_dir;
        SSB_FL(f, wall_in, il) = (direct ? win_dir : 0.0) + win;
        SSB_FL(f, wall_net, il) = SSB_FL(f, wall_in, il) * (1.0 - walb);
      }
      if (direct) {
        const double nonb_here = URBAN ? 1.0 - bf : 1.0;
        double nonb_above = 1.0;
        if (URBAN && jl + 1 < nlay) nonb_above = 1.0 - a.cp.building_fraction[il + 1];
        if (URBAN) {
          const double roof_fraction = (jl == nlay - 1) ? bf : dmax(0.0, bf - a.cp.building_fraction[il + 1]);
          if (own && f.roof_sunlit_frac)
            f.roof_sunlit_frac[il] = SSB_FL(f, roof_in_dir, il) * nonb_above /
                                     (zcos * flux_dn_dir_clear * dmax(c.min_bld, roof_fraction));
          flux_dn_dir_clear = flux_dn_dir_clear * nonb_here / nonb_above;
        }
        const double air_ext_t = a.sw.air_ext[(size_t)itransp + (size_t)nspec * il];
        const double trans_dir_clear = exp(-air_ext_t * a.cp.dz[il] / zcos);
        const double int_flux_dir_clear = (air_ext_t > 0.0)
                                              ? flux_dn_dir_clear * (1.0 - trans_dir_clear) * zcos / air_ext_t
                                              : flux_dn_dir_clear * a.cp.dz[il];
        if (own) {
          if ((URBAN ? NREG > 1 : true) && f.veg_sunlit_frac && a.cp.veg_ext && a.cp.veg_fraction && a.sw.veg_ssa) {
            const double veg_abs_dir_clear = int_flux_dir_clear * ve * (1.0 - SSB_LAY(a.sw.veg_ssa, g, il)) * vf;
            f.veg_sunlit_frac[il] = SSB_FL(f, veg_abs_dir, il) / dmax(SSB_EPS, veg_abs_dir_clear);
          }
          if (URBAN && f.wall_sunlit_frac)
            f.wall_sunlit_frac[il] =
                0.5 * SSB_FL(f, wall_in_dir, il) / dmax(SSB_EPS, (gm.f_wall_dir_clear * sin0 * int_flux_dir_clear));
        }
        flux_dn_dir_clear = flux_dn_dir_clear * trans_dir_clear;
      }
    }
    double s_dir = 0.0, s_dn = 0.0, s_up = 0.0, s_vert = 0.0;
    SSB_UNROLL
    for (int i = 0; i < d; ++i) s_dir += dir_above[i];
    SSB_UNROLL
    for (int r = 0; r < NREG; ++r) {
      SSB_UNROLL
      for (int js = 0; js < NS; ++js) {
        const int i = js + r * NS;
        s_dn += diff_above[i];
        s_up += up_above[i];
        s_vert += (diff_above[i] + up_above[i]) * tang[js] / SSB_PI;
      }
    }
    SSB_FC(f, ground_dn_dir) = direct ? zcos * s_dir : 0.0;
    SSB_FC(f, ground_dn) = SSB_FC(f, ground_dn_dir) + s_dn;
    SSB_FC(f, ground_net) = SSB_FC(f, ground_dn) - s_up;
    SSB_FC(f, ground_vertical_diff) = s_vert;
    if (direct && own && f.ground_sunlit_frac)
      f.ground_sunlit_frac[col] = SSB_FC(f, ground_dn_dir) / (zcos * flux_dn_dir_clear);
  }
}

template <int NREG, int NS, bool URBAN>
SSB_HD inline void fast_column_sweeps_lw(const ClassArgs &a, int q) {
  constexpr int n = NREG * NS, d = NREG, NRB = URBAN ? NREG + 1 : NREG, m = NRB * NS;
  const SolveCfg &c = a.cfg;
  const int nspec = c.nspec;
  const int ic = q / nspec, g = q % nspec;
  const int col = a.cols[ic];
  const int nlay = a.nlay[col], il1 = a.istartlay[col] - 1;
  const int width = a.ncols * nspec;
  const ssb200_canopy_flux &fint = a.f1, &fnorm = a.f2;
  zero_column(fint, nspec, g, col, il1, nlay, g == 0);
  zero_column(fnorm, nspec, g, col, il1, nlay, g == 0);
  double hw[NS], mu_inv[NS], tang[NS];
  SSB_UNROLL
  for (int js = 0; js < NS; ++js) {
    hw[js] = a.lg.hweight[js];
    mu_inv[js] = 1.0 / a.lg.mu[js];
    tang[js] = a.lg.tan_ang[js];
  }
  const double *LS = a.layer;
  double *WS = a.sweep;
  const int nlev = a.lmax, wlev = a.lmax + 1;
  constexpr int oR = 0, oT = n * n, oIF = 2 * n * n, oSrc = 3 * n * n, oIsrc = oSrc + n, oBook = oIsrc + n;
  constexpr int oAa = 0, oSa = n * n, oLU = oSa + n;

  const double gemis = a.lw.ground_emissivity[(size_t)g + (size_t)nspec * col];
  const double gemission = a.lw.ground_emission[(size_t)g + (size_t)nspec * col];
  double Aa[n * n], Sa[n];
  {
    double frac0[3] = {1.0, 0.0, 0.0};
    if (nlay > 0) {
      const bool veg = NREG > 1 || !URBAN;
      region_fractions(c, URBAN ? a.cp.building_fraction[il1] : 0.0,
                       (veg && a.cp.veg_fraction) ? a.cp.veg_fraction[il1] : 0.0, frac0);
    }
    SSB_UNROLL
    for (int i = 0; i < n * n; ++i) Aa[i] = 0.0;
    SSB_UNROLL
    for (int r = 0; r < NREG; ++r) {
      SSB_UNROLL
      for (int jt = 0; jt < NS; ++jt) {
        SSB_UNROLL
        for (int jf = 0; jf < NS; ++jf) Aa[(jt + r * NS) + n * (jf + r * NS)] = (1.0 - gemis) * hw[jt];
        Sa[jt + r * NS] = (hw[jt] * frac0[r]) * gemission;
      }
    }
  }
  sstore<n, n>(WS, oAa, 0, wlev, width, q, Aa);
  sstore<n, 1>(WS, oSa, 0, wlev, width, q, Sa);
  for (int jl = 0; jl < nlay; ++jl) {
    const int il = il1 + jl;
    double R[n * n], T[n * n], LU[n * n], X[n * n], src[n];
    sload<n, n>(LS, oR, jl, nlev, width, q, R);
    sload<n, n>(LS, oT, jl, nlev, width, q, T);
    sload<n, 1>(LS, oSrc, jl, nlev, width, q, src);
    sm_mul<n, n, n>(Aa, R, LU);
    SSB_UNROLL
    for (int j = 0; j < n; ++j) {
      SSB_UNROLL
      for (int i = 0; i < n; ++i) LU[i + n * j] = (i == j ? 1.0 : 0.0) - LU[i + n * j];
    }
    sm_lu<n>(LU);
    sstore<n, n>(WS, oLU, jl, wlev, width, q, LU);
    sm_mul<n, n, n>(Aa, T, X);
    sm_lu_solve_left<n, n>(LU, X);
    double Ab[n * n];
    sm_mul<n, n, n>(T, X, Ab);
    SSB_UNROLL
    for (int i = 0; i < n * n; ++i) Ab[i] += R[i];
    // source_below (street part) = src + T D^-1 (Sa + Aa src)
    double Sb[n], v1[n];
    sm_mulvec<n, n>(Aa, src, v1);
    SSB_UNROLL
    for (int i = 0; i < n; ++i) v1[i] += Sa[i];
    sm_lu_solve_left<n, 1>(LU, v1);
    sm_mulvec<n, n>(T, v1, Sb);
    SSB_UNROLL
    for (int i = 0; i < n; ++i) Sb[i] += src[i];
    double rb[NS], rs[NS];
    SSB_UNROLL
    for (int js = 0; js < NS; ++js) rb[js] = rs[js] = 0.0;
    if (URBAN) {
      const double bfj = a.cp.building_fraction[il];
      const double exposed = (jl < nlay - 1) ? dmax(0.0, bfj - a.cp.building_fraction[il + 1]) : bfj;
      const double remis = SSB_LAY(a.lw.roof_emissivity, g, il), remission = SSB_LAY(a.lw.roof_emission, g, il);
      SSB_UNROLL
      for (int js = 0; js < NS; ++js) {
        rb[js] = (1.0 - remis) * hw[js];
        rs[js] = hw[js] * remission * exposed;
      }
    }
    double U[12], V[12];
    overlap_at(a, il1, nlay, jl + 1, U, V);
    {
      double AV[n * n];
      SSB_UNROLL
      for (int up = 0; up < NREG; ++up) {
        SSB_UNROLL
        for (int js = 0; js < NS; ++js) {
          SSB_UNROLL
          for (int i = 0; i < n; ++i) {
            double s = 0.0;
            SSB_UNROLL
            for (int lo = 0; lo < NREG; ++lo) s = fma(Ab[i + n * (lo * NS + js)], V[lo + NRB * up], s);
            AV[i + n * (up * NS + js)] = s;
          }
        }
      }
      SSB_UNROLL
      for (int up = 0; up < NREG; ++up) {
        SSB_UNROLL
        for (int js = 0; js < NS; ++js) {
          SSB_UNROLL
          for (int u = 0; u < NREG; ++u) {
            SSB_UNROLL
            for (int jt = 0; jt < NS; ++jt) {
              double s = 0.0;
              SSB_UNROLL
              for (int lo = 0; lo < NREG; ++lo) s = fma(U[u + NREG * lo], AV[(lo * NS + jt) + n * (up * NS + js)], s);
              if (URBAN) s = fma(U[u + NREG * NREG] * rb[jt], V[NREG + NRB * up], s);
              Aa[(u * NS + jt) + n * (up * NS + js)] = s;
            }
          }
        }
      }
      SSB_UNROLL
      for (int u = 0; u < NREG; ++u) {
        SSB_UNROLL
        for (int jt = 0; jt < NS; ++jt) {
          double s = 0.0;
          SSB_UNROLL
          for (int lo = 0; lo < NREG; ++lo) s = fma(U[u + NREG * lo], Sb[lo * NS + jt], s);
          if (URBAN) s = fma(U[u + NREG * NREG], rs[jt], s);
          Sa[u * NS + jt] = s;
        }
      }
    }
    sstore<n, n>(WS, oAa, jl + 1, wlev, width, q, Aa);
    sstore<n, 1>(WS, oSa, jl + 1, wlev, width, q, Sa);
  }
  double top_emissivity, top_emission = 0.0;
  {
    double sAll = 0.0;
    SSB_UNROLL
    for (int i = 0; i < NS; ++i) {
      double s = 0.0;
      SSB_UNROLL
      for (int j = 0; j < NS; ++j) s = fma(Aa[i + n * j], hw[j], s);
      sAll += s;
      top_emission += Sa[i];
    }
    top_emissivity = 1.0 - sAll;
    a.bc.lw_emissivity[(size_t)g + (size_t)nspec * col] = top_emissivity;
    a.bc.lw_emission[(size_t)g + (size_t)nspec * col] = top_emission;
  }

  double gvd_internal = 0.0;
  for (int pass = 0; pass < 2; ++pass) {
    const bool internal = (pass == 0);
    const ssb200_canopy_flux &f = internal ? fint : fnorm;
    double dn_above[n], up_above[n];
    SSB_UNROLL
    for (int i = 0; i < n; ++i) {
      dn_above[i] = 0.0;
      up_above[i] = 0.0;
    }
    if (internal) {
      SSB_FC(f, top_dn) = 0.0;
      SSB_FC(f, top_net) = -top_emission;
    } else {
      SSB_UNROLL
      for (int js = 0; js < NS; ++js) dn_above[js] = hw[js];
      SSB_FC(f, top_dn) = 1.0;
      SSB_FC(f, top_net) = top_emissivity;
    }
    for (int jl = nlay - 1; jl >= 0; --jl) {
      const int il = il1 + jl;
      LayerGeom gm;
      double bf, vf, ve;
      geometry_of_layer(a, il, a.lg.vadjustment2, gm, bf, vf, ve);
      double U[12], V[12];
      overlap_at(a, il1, nlay, jl + 1, U, V);
      double dn_below[m], up_below[m];
      expand_down<NREG, NRB, NS>(V, dn_above, dn_below);
      double LU[n * n], src[n], sa[n];
      sload<n, n>(WS, oLU, jl, wlev, width, q, LU);
      SSB_UNROLL
      for (int i = 0; i < n; ++i) {
        src[i] = 0.0;
        sa[i] = 0.0;
      }
      if (internal) {
        sload<n, 1>(LS, oSrc, jl, nlev, width, q, src);
        sload<n, 1>(WS, oSa, jl, wlev, width, q, sa);
      }
      // y = T x + src
      double y[n];
      SSB_UNROLL
      for (int i = 0; i < n; ++i) y[i] = 0.0;
      smv_acc<n, n>(LS, oT, jl, nlev, width, q, dn_below, y);
      SSB_UNROLL
      for (int i = 0; i < n; ++i) y[i] += src[i];
      // z1 = D^-1 (Aa y + Sa) ; z2 = D^-1 (y + R Sa)
      double z1[n], z2[n];
      SSB_UNROLL
      for (int i = 0; i < n; ++i) {
        z1[i] = sa[i];
        z2[i] = y[i];
      }
      smv_acc<n, n>(WS, oAa, jl, wlev, width, q, y, z1);
      if (internal) smv_acc<n, n>(LS, oR, jl, nlev, width, q, sa, z2);
      sm_lu_solve_left<n, 1>(LU, z1);
      sm_lu_solve_left<n, 1>(LU, z2);
      SSB_UNROLL
      for (int i = 0; i < m; ++i) up_below[i] = 0.0;
      SSB_UNROLL
      for (int i = 0; i < n; ++i) up_below[i] = src[i];
      smv_acc<n, n>(LS, oR, jl, nlev, width, q, dn_below, up_below);
      smv_acc<n, n>(LS, oT, jl, nlev, width, q, z1, up_below);
      if (URBAN) {
        const double bfj = a.cp.building_fraction[il];
        const double exposed = (jl < nlay - 1) ? dmax(0.0, bfj - a.cp.building_fraction[il + 1]) : bfj;
        const double remis = SSB_LAY(a.lw.roof_emissivity, g, il), remission = SSB_LAY(a.lw.roof_emission, g, il);
        double sroof = 0.0, roof_up = 0.0;
        SSB_UNROLL
        for (int js = 0; js < NS; ++js) sroof += dn_below[n + js];
        SSB_UNROLL
        for (int js = 0; js < NS; ++js) {
          up_below[n + js] = (1.0 - remis) * hw[js] * sroof + (internal ? hw[js] * remission * exposed : 0.0);
          roof_up += up_below[n + js];
        }
        SSB_FL(f, roof_in, il) = sroof;
        SSB_FL(f, roof_net, il) = sroof - roof_up;
      }
      SSB_UNROLL
      for (int i = 0; i < n; ++i) {
        dn_above[i] = z2[i];
        up_above[i] = sa[i];
      }
      smv_acc<n, n>(WS, oAa, jl, wlev, width, q, dn_above, up_above);
      if (f.flux_dn_layer_top) {
        double s_db = 0.0, s_da = 0.0, s_ub = 0.0, s_ua = 0.0;
        SSB_UNROLL
        for (int i = 0; i < n; ++i) {
          s_db += dn_below[i];
          s_da += dn_above[i];
          s_ub += up_below[i];
          s_ua += up_above[i];
        }
        SSB_FL(f, flux_dn_layer_top, il) = s_db;
        SSB_FL(f, flux_up_layer_top, il) = s_ub;
        SSB_FL(f, flux_dn_layer_base, il) = s_da;
        SSB_FL(f, flux_up_layer_base, il) = s_ua;
      }
      double tv[n], iflux[n], book[3 * d + 1];
      SSB_UNROLL
      for (int i = 0; i < n; ++i) {
        tv[i] = dn_below[i] + up_above[i];
        iflux[i] = 0.0;
      }
      SSB_UNROLL
      for (int i = 0; i < 3 * d + 1; ++i) book[i] = 0.0;
      smv_acc<n, n>(LS, oIF, jl, nlev, width, q, tv, iflux);
      if (internal) {
        double isrc[n];
        sload<n, 1>(LS, oIsrc, jl, nlev, width, q, isrc);
        SSB_UNROLL
        for (int i = 0; i < n; ++i) iflux[i] += isrc[i];
        sload<3 * d + 1, 1>(LS, oBook, jl, nlev, width, q, book);
      }
      double smu[NREG], stan[NREG];
      SSB_UNROLL
      for (int r = 0; r < NREG; ++r) {
        smu[r] = 0.0;
        stan[r] = 0.0;
        SSB_UNROLL
        for (int js = 0; js < NS; ++js) {
          smu[r] = fma(iflux[r * NS + js], mu_inv[js], smu[r]);
          stan[r] = fma(iflux[r * NS + js], tang[js], stan[r]);
        }
      }
      const double dz = a.cp.dz[il];
      const double air_abs = SSB_LAY(a.lw.air_ext, g, il) * (1.0 - SSB_LAY(a.lw.air_ssa, g, il));
      SSB_FL(f, clear_air_abs, il) = air_abs * smu[0] - book[0] * dz;
      if (NREG > 1) {
        const double vabs = ve * (1.0 - SSB_LAY(a.lw.veg_ssa, g, il));
        double s_air = 0.0, s_veg = 0.0;
        SSB_UNROLL
        for (int r = 1; r < NREG; ++r) {
          s_air += air_abs * smu[r] - book[d + r] * dz;
          s_veg += vabs * smu[r] * gm.od_scaling[r] - book[2 * d + r] * dz;
        }
        SSB_FL(f, veg_air_abs, il) = s_air;
        SSB_FL(f, veg_abs, il) = s_veg;
      }
      if (URBAN) {
        double win = 0.0;
        SSB_UNROLL
        for (int r = 0; r < NREG; ++r) win += gm.f_wall[r] * stan[r];
        const double wemis = SSB_LAY(a.lw.wall_emissivity, g, il);
        SSB_FL(f, wall_in, il) = win;
        SSB_FL(f, wall_net, il) = win * wemis - book[3 * d] * dz;
      }
    }
    double s_dn = 0.0, s_up = 0.0, s_vert = 0.0;
    SSB_UNROLL
    for (int r = 0; r < NREG; ++r) {
      SSB_UNROLL
      for (int js = 0; js < NS; ++js) {
        const int i = js + r * NS;
        s_dn += dn_above[i];
        s_up += up_above[i];
        s_vert += (dn_above[i] + up_above[i]) * tang[js] / SSB_PI;
      }
    }
    SSB_FC(f, ground_dn) = s_dn;
    SSB_FC(f, ground_net) = s_dn - s_up;
    // forest_lw:687-694 accumulates the normalised pass into lw_internal as well
    if (internal) {
      gvd_internal = s_vert;
      SSB_FC(f, ground_vertical_diff) = s_vert;
    } else if (URBAN) {
      SSB_FC(f, ground_vertical_diff) = s_vert;
    } else {
      SSB_FC(fint, ground_vertical_diff) = gvd_internal + s_vert;
    }
  }
}

}  // namespace ssb
