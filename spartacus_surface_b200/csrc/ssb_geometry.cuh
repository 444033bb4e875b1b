// ssb_geometry.cuh - spectrally independent canopy geometry of one layer /
// one interface, evaluated on the fly inside the solver kernels.
//
// Replaces radsurf/radsurf_overlap.F90:28-73,85-171,178-280,289-394
// (max-random overlap U,V per interface, urban variant with the exposed-roof
// pseudo-region and the overhang rescaling), radsurf/radsurf_norm_perim.F90:
// 30-116,131-281 (perimeter lengths) and radsurf/radsurf_view_factor.F90:
// 28-70,76-138, plus the per-layer exchange / wall rates of
// radsurf_urban_sw.F90:284-298,342-418 and radsurf_urban_lw.F90:303-392.
//
// The reference computes these per column into (nreg,nlay) arrays before the
// layer loop; every quantity depends on one layer (or on the two layers
// either side of an interface), so here they are recomputed where needed and
// never stored per column.
#pragma once
#include "ssb_math.cuh"

namespace ssb {

struct SolveCfg {
  int urban;      // 1: spartacus_urban_*, 0: spartacus_forest_*
  int lw;         // 1: longwave variant
  int nreg, ns, nspec;
  int symmetric_scale;
  double isolation, min_veg, min_bld;
};

struct LayerGeom {
  double frac[3];
  double norm_perim[3], norm_perim_wall[3];
  double f_exchange[9];  // [to + 3*from]
  double f_wall[3];
  double f_wall_dir_clear;
  double od_scaling[3];
  int r0, nr;  // first region and number of regions actually solved in this layer
};

// Region fractions of one layer (urban_sw:284-291, forest_sw:244-248, forest_lw:233-236).
SSB_HDI void region_fractions(const SolveCfg &c, double bf, double vf, double *frac) {
  const int nreg = c.nreg;
  if (c.urban) {
    frac[0] = 1.0 - bf;
    if (nreg > 1) {
      frac[0] = dmax(0.0, frac[0] - vf);
      const double fv = dmax(0.0, 1.0 - bf - frac[0]) / (double)(nreg - 1);
      for (int r = 1; r < nreg; ++r) frac[r] = fv;
    }
  } else {
    frac[0] = 1.0 - vf;
    for (int r = 1; r < nreg; ++r)
      frac[r] = (c.lw ? (1.0 - frac[0]) : vf) / (double)(nreg - 1);
  }
  for (int r = nreg; r < 3; ++r) frac[r] = 0.0;
}

SSB_HDI double sum_frac(int nreg, const double *frac) {
  double s = 0.0;
  for (int r = 0; r < nreg; ++r) s += frac[r];
  return s;
}

// wall_adj = 1 for SW, lg%vadjustment2 for LW (urban_lw:379).
SSB_HD inline void layer_geometry(const SolveCfg &c, double bf, double bs, double vf, double vs, double vcf,
                                  double vfsd, double wall_adj, LayerGeom &g) {
  const int nreg = c.nreg;
  region_fractions(c, bf, vf, g.frac);
  for (int r = 0; r < 3; ++r) {
    g.norm_perim[r] = 0.0;
    g.norm_perim_wall[r] = 0.0;
    g.f_wall[r] = 0.0;
    g.od_scaling[r] = 0.0;
  }
  for (int i = 0; i < 9; ++i) g.f_exchange[i] = 0.0;
  // -- perimeter lengths -----------------------------------------------------
  if (nreg > 1 && vf > c.min_veg) {
    if (c.urban) {
      if (c.symmetric_scale)
        g.norm_perim[0] = 4.0 * vf * dmax(0.0, 1.0 - vf - bf) / (dmax(c.min_bld, 1.0 - bf) * vs);
      else
        g.norm_perim[0] = 4.0 * vf / vs;
    } else {
      if (c.symmetric_scale)
        g.norm_perim[0] = 4.0 * vf * dmax(0.0, 1.0 - vf) / vs;
      else
        g.norm_perim[0] = 4.0 * vf / vs;
    }
    if (nreg > 2) {
      g.norm_perim[nreg - 1] = 0.5 * c.isolation * g.norm_perim[0];
      g.norm_perim[0] = (1.0 - 0.5 * c.isolation) * g.norm_perim[0];
      if (c.symmetric_scale) {
        if (c.urban)
          g.norm_perim[1] = (1.0 - c.isolation) * 4.0 * (0.5 * vf) * (1.0 - (0.5 * vf) - bf) /
                            (dmax(c.min_bld, 1.0 - bf) * vs);
        else
          g.norm_perim[1] = (1.0 - c.isolation) * 4.0 * (0.5 * vf) * (1.0 - (0.5 * vf)) / vs;
      } else {
        g.norm_perim[1] = (1.0 - c.isolation) * 4.0 * vf / (sqrt(2.0) * vs);
      }
    }
  }
  if (c.urban && bf > c.min_bld) {
    double *w = g.norm_perim_wall;
    w[0] = 4.0 * bf / bs;
    if (nreg > 1) {
      if (1.0 - vf - bf <= c.min_veg) {
        if (nreg == 2) {
          w[1] = w[0];
        } else {
          w[1] = w[0] * (1.0 - c.isolation);
          w[2] = w[0] * c.isolation;
        }
        w[0] = 0.0;
      } else if (vf > c.min_veg) {
        if (vcf > 0.0) {
          if (nreg == 2) {
            w[1] = w[0] * vcf;
          } else {
            w[1] = w[0] * vcf * (1.0 - c.isolation);
            w[2] = w[0] * vcf * c.isolation;
          }
          w[0] = w[0] * (1.0 - vcf);
        }
      }
    }
  }
  // -- exchange and wall rates (urban_sw:373-410) ----------------------------
  for (int r = 0; r < nreg - 1; ++r) {
    if (!(g.frac[r] <= c.min_veg || g.frac[r + 1] <= c.min_veg)) {
      g.f_exchange[(r + 1) + 3 * r] = g.norm_perim[r] / (SSB_PI * g.frac[r]);
      g.f_exchange[r + 3 * (r + 1)] = g.norm_perim[r] / (SSB_PI * g.frac[r + 1]);
    }
  }
  if (nreg > 2 && g.norm_perim[nreg - 1] > 0.0) {
    if (!(g.frac[2] <= c.min_veg || g.frac[0] <= c.min_veg)) {
      g.f_exchange[0 + 3 * 2] = g.norm_perim[nreg - 1] / (SSB_PI * g.frac[2]);
      g.f_exchange[2 + 3 * 0] = g.norm_perim[nreg - 1] / (SSB_PI * g.frac[0]);
    }
  }
  g.f_wall_dir_clear = 0.0;
  if (c.urban) {
    for (int r = 0; r < nreg; ++r) {
      if (g.frac[r] <= c.min_veg)
        g.f_wall[r] = 0.0;
      else if (c.lw)
        g.f_wall[r] = g.norm_perim_wall[r] * wall_adj / (SSB_PI * g.frac[r]);
      else
        g.f_wall[r] = g.norm_perim_wall[r] / (SSB_PI * g.frac[r]);
    }
    const double nonb = 1.0 - bf;
    if (!(nonb <= c.min_bld)) {
      double s = 0.0;
      for (int r = 0; r < nreg; ++r) s += g.norm_perim_wall[r];
      g.f_wall_dir_clear = s / (SSB_PI * nonb);
    }
  }
  if (nreg == 2) {
    g.od_scaling[1] = 1.0;
  } else if (nreg == 3) {
    g.od_scaling[1] = exp(-vfsd * (1.0 + 0.5 * vfsd * (1.0 + 0.5 * vfsd)));
    g.od_scaling[2] = 2.0 - g.od_scaling[1];
  }
  // -- which regions are solved (urban_sw:512-583) ---------------------------
  const bool veg_branching = c.urban ? (nreg > 1) : true;
  if (veg_branching) {
    if (vf <= c.min_veg) {
      g.r0 = 0;
      g.nr = 1;
    } else if (g.frac[0] <= c.min_veg) {
      g.r0 = 1;
      g.nr = nreg - 1;
    } else {
      g.r0 = 0;
      g.nr = nreg;
    }
  } else {
    g.r0 = 0;
    g.nr = 1;
  }
}

// Overlap matrices at one interface.  U is (nreg x nrb) at [up + nreg*lo],
// V is (nrb x nreg) at [lo + nrb*up]; nrb = nreg (forest) or nreg+1 (urban,
// last lower "region" = exposed roof).  fu = fractions just above, fl = just
// below (fl[nreg] = roof fraction for urban).
SSB_HD inline void overlap_interface(const SolveCfg &c, const double *fu, const double *fl, double *U,
                                     double *V) {
  const int nreg = c.nreg;
  const int nrb = c.urban ? nreg + 1 : nreg;
  double O[12];  // [up + nreg*lo]
  for (int i = 0; i < 12; ++i) O[i] = 0.0;
#define OV(up, lo) O[(up) + nreg * (lo)]
  if (!c.urban) {
    const double f_upper = 1.0 - fu[0], f_lower = 1.0 - fl[0];
    const double pair_cover = dmax(f_upper, f_lower);
    OV(0, 0) = 1.0 - pair_cover;
    if (nreg == 2) {
      OV(0, 1) = pair_cover - f_upper;
      OV(1, 0) = pair_cover - f_lower;
      OV(1, 1) = f_upper + f_lower - pair_cover;
    } else if (nreg == 3) {
      OV(0, 1) = 0.5 * (pair_cover - f_upper);
      OV(0, 2) = OV(0, 1);
      OV(1, 0) = 0.5 * (pair_cover - f_lower);
      OV(2, 0) = OV(1, 0);
      OV(1, 1) = 0.5 * (f_upper + f_lower - pair_cover);
      OV(2, 2) = OV(1, 1);
    }
  } else if (nreg == 1) {
    OV(0, 0) = fl[0];
    OV(0, 1) = fl[1];
  } else if (nreg == 2) {
    const double pair_cover = dmax(fu[1], fl[1]);
    if (pair_cover <= fl[0] + fl[1]) {
      OV(0, 2) = fl[2];
      OV(0, 0) = fl[0] + fl[1] - pair_cover;
      OV(0, 1) = pair_cover - fu[1];
      OV(1, 0) = pair_cover - fl[1];
      OV(1, 1) = fu[1] + fl[1] - pair_cover;
    } else {
      OV(1, 0) = fl[0];
      OV(1, 1) = fl[1];
      OV(1, 2) = fu[1] - fl[0] - fl[1];
      OV(0, 2) = fu[0];
    }
  } else {
    const double pair_cover = dmax(fu[1] + fu[2], fl[1] + fl[2]);
    if (pair_cover <= fl[0] + fl[1] + fl[2]) {
      OV(0, 3) = fl[3];
      OV(0, 0) = fl[0] + fl[1] + fl[2] - pair_cover;
      if (pair_cover > fu[1] + fu[2]) {
        OV(1, 1) = fu[1];
        OV(2, 2) = fu[2];
        OV(0, 1) = fl[1] - fu[1];
        OV(0, 2) = fl[2] - fu[2];
      } else {
        OV(1, 1) = fl[1];
        OV(2, 2) = fl[2];
        OV(1, 0) = fu[1] - fl[1];
        OV(2, 0) = fu[2] - fl[2];
      }
    } else {
      // overhang: vegetation above extends over the roof below
      OV(1, 1) = fl[1];
      OV(2, 2) = fl[2];
      OV(1, 0) = fl[0] * 0.5;
      OV(2, 0) = OV(0, 1);  // reference assigns O(3,1) = O(1,2), which is zero here (overlap:268)
      OV(1, 3) = (fl[3] - fu[0]) * 0.5;
      OV(2, 3) = OV(1, 3);
      OV(0, 3) = fu[0];
    }
  }
  for (int up = 0; up < nreg; ++up)
    for (int lo = 0; lo < nrb; ++lo) {
      U[up + nreg * lo] = (fl[lo] >= c.min_veg) ? OV(up, lo) / fl[lo] : 0.0;
      V[lo + nrb * up] = (fu[up] >= c.min_veg) ? OV(up, lo) / fu[up] : 0.0;
    }
#undef OV
}

// Fractions either side of interface k (0 = ground .. nlay = canopy top) from
// the region fractions of the layer below (fb, layer k-1) and above (fa, layer k).
SSB_HD inline void interface_fractions(const SolveCfg &c, int k, int nlay, const double *fb, const double *fa,
                                       double *fu, double *fl) {
  const int nreg = c.nreg;
  for (int r = 0; r < 4; ++r) fl[r] = 0.0;
  for (int r = 0; r < 3; ++r) fu[r] = 0.0;
  if (k < nlay) {
    for (int r = 0; r < nreg; ++r) fu[r] = fa[r];
  } else {
    fu[0] = 1.0;
  }
  if (!c.urban) {
    if (k == 0)
      fl[0] = 1.0;
    else
      for (int r = 0; r < nreg; ++r) fl[r] = fb[r];
    return;
  }
  if (k == 0) {
    fl[nreg] = sum_frac(nreg, fa);  // the ground is one "roof" region (overlap:333-334)
    return;
  }
  for (int r = 0; r < nreg; ++r) fl[r] = fb[r];
  if (k < nlay) {
    const double sa = sum_frac(nreg, fa), sb = sum_frac(nreg, fb);
    fl[nreg] = sa - sb;
    if (fl[nreg] < 0.0) {  // overhanging building (overlap:376-388)
      for (int r = 0; r < nreg; ++r) fl[r] = fl[r] * sa / sb;
      fl[nreg] = 0.0;
    }
  } else {
    fl[nreg] = 1.0 - sum_frac(nreg, fb);
  }
}

// Compile-time variant of region_fractions + interface_fractions + overlap_interface for the
// register-resident sweeps (interfaces k = 1..nlay): same case analysis, orders known at
// compile time, the divisions by the interface fractions replaced by 7 reciprocals.
// `fb`, `fa`: region fractions of the layers below and above the interface (fa unused at the
// canopy top).  U is (NREG x NRB) at [up + NREG*lo], V is (NRB x NREG) at [lo + NRB*up].
template <int NREG, bool URBAN, bool LW>
SSB_HDI void region_fractions_t(double bf, double vf, double *frac) {
  if (URBAN) {
    frac[0] = 1.0 - bf;
    if (NREG > 1) {
      frac[0] = dmax(0.0, frac[0] - vf);
      const double fv = dmax(0.0, 1.0 - bf - frac[0]) * (NREG == 3 ? 0.5 : 1.0);
      SSB_UNROLL
      for (int r = 1; r < NREG; ++r) frac[r] = fv;
    }
  } else {
    frac[0] = 1.0 - vf;
    SSB_UNROLL
    for (int r = 1; r < NREG; ++r) frac[r] = (LW ? (1.0 - frac[0]) : vf) * (NREG == 3 ? 0.5 : 1.0);
  }
}

template <int NREG, bool URBAN>
SSB_HDI void overlap_fast(const double *fb, const double *fa, bool top, double min_veg, double *U, double *V) {
  constexpr int NRB = URBAN ? NREG + 1 : NREG;
  double fu[NREG], fl[NRB], O[NREG * NRB];
  SSB_UNROLL
  for (int r = 0; r < NREG; ++r) {
    fu[r] = top ? (r == 0 ? 1.0 : 0.0) : fa[r];
    fl[r] = fb[r];
  }
  if (URBAN) {
    double sa = 0.0, sb = 0.0;
    SSB_UNROLL
    for (int r = 0; r < NREG; ++r) {
      sa += fa[r];
      sb += fb[r];
    }
    if (top) {
      fl[NREG] = 1.0 - sb;
    } else {
      fl[NREG] = sa - sb;
      if (fl[NREG] < 0.0) {  // overhanging building (radsurf_overlap.F90:376-388)
        const double sc = sa / sb;
        SSB_UNROLL
        for (int r = 0; r < NREG; ++r) fl[r] = fl[r] * sc;
        fl[NREG] = 0.0;
      }
    }
  }
  SSB_UNROLL
  for (int i = 0; i < NREG * NRB; ++i) O[i] = 0.0;
#define OV(up, lo) O[(up) + NREG * (lo)]
  if (!URBAN) {
    const double f_upper = 1.0 - fu[0], f_lower = 1.0 - fl[0];
    const double pair_cover = dmax(f_upper, f_lower);
    OV(0, 0) = 1.0 - pair_cover;
    if (NREG == 2) {
      OV(0, 1) = pair_cover - f_upper;
      OV(1, 0) = pair_cover - f_lower;
      OV(1, 1) = f_upper + f_lower - pair_cover;
    } else if (NREG == 3) {
      OV(0, 1) = 0.5 * (pair_cover - f_upper);
      OV(0, 2) = OV(0, 1);
      OV(1, 0) = 0.5 * (pair_cover - f_lower);
      OV(2, 0) = OV(1, 0);
      OV(1, 1) = 0.5 * (f_upper + f_lower - pair_cover);
      OV(2, 2) = OV(1, 1);
    }
  } else if (NREG == 1) {
    OV(0, 0) = fl[0];
    OV(0, 1) = fl[1];
  } else if (NREG == 2) {
    const double pair_cover = dmax(fu[1], fl[1]);
    if (pair_cover <= fl[0] + fl[1]) {
      OV(0, 2) = fl[2];
      OV(0, 0) = fl[0] + fl[1] - pair_cover;
      OV(0, 1) = pair_cover - fu[1];
      OV(1, 0) = pair_cover - fl[1];
      OV(1, 1) = fu[1] + fl[1] - pair_cover;
    } else {
      OV(1, 0) = fl[0];
      OV(1, 1) = fl[1];
      OV(1, 2) = fu[1] - fl[0] - fl[1];
      OV(0, 2) = fu[0];
    }
  } else {
    constexpr int r1 = NREG > 1 ? 1 : 0, r2 = NREG > 2 ? 2 : 0, rr = NRB - 1;  // (valid indices for every NREG)
    const double pair_cover = dmax(fu[r1] + fu[r2], fl[r1] + fl[r2]);
    if (pair_cover <= fl[0] + fl[r1] + fl[r2]) {
      OV(0, rr) = fl[rr];
      OV(0, 0) = fl[0] + fl[r1] + fl[r2] - pair_cover;
      if (pair_cover > fu[r1] + fu[r2]) {
        OV(r1, r1) = fu[r1];
        OV(r2, r2) = fu[r2];
        OV(0, r1) = fl[r1] - fu[r1];
        OV(0, r2) = fl[r2] - fu[r2];
      } else {
        OV(r1, r1) = fl[r1];
        OV(r2, r2) = fl[r2];
        OV(r1, 0) = fu[r1] - fl[r1];
        OV(r2, 0) = fu[r2] - fl[r2];
      }
    } else {
      // overhang: vegetation above extends over the roof below
      OV(r1, r1) = fl[r1];
      OV(r2, r2) = fl[r2];
      OV(r1, 0) = fl[0] * 0.5;
      OV(r2, 0) = OV(0, r1);  // the reference assigns O(3,1) = O(1,2), zero here (radsurf_overlap.F90:268)
      OV(r1, rr) = (fl[rr] - fu[0]) * 0.5;
      OV(r2, rr) = OV(r1, rr);
      OV(0, rr) = fu[0];
    }
  }
  double rfl[NRB], rfu[NREG];
  SSB_UNROLL
  for (int lo = 0; lo < NRB; ++lo) rfl[lo] = (fl[lo] >= min_veg) ? 1.0 / fl[lo] : 0.0;
  SSB_UNROLL
  for (int up = 0; up < NREG; ++up) rfu[up] = (fu[up] >= min_veg) ? 1.0 / fu[up] : 0.0;
  SSB_UNROLL
  for (int up = 0; up < NREG; ++up) {
    SSB_UNROLL
    for (int lo = 0; lo < NRB; ++lo) {
      U[up + NREG * lo] = OV(up, lo) * rfl[lo];
      V[lo + NRB * up] = OV(up, lo) * rfu[up];
    }
  }
#undef OV
}

// View factors of the single-layer urban models (radsurf_view_factor.F90).
SSB_HD inline void view_factors(bool infinite_street, double h, bool with_sun, double cos_sza,
                                double &view_ground_sky, double &view_wall_wall, double &view_dir_ground) {
  view_dir_ground = 0.0;
  if (infinite_street) {
    view_ground_sky = sqrt(h * h + 1.0) - h;
    view_wall_wall = sqrt(1.0 / (h * h) + 1.0) - 1.0 / h;
    if (with_sun) {
      const double norm_x0 = (SSB_PI * 0.5) * h * sqrt(1.0 / (cos_sza * cos_sza) - 1.0);
      const double y_over_w = sqrt(dmax(norm_x0 * norm_x0 - 1.0, 0.0));
      if (y_over_w > 0.0)
        view_dir_ground = (2.0 / SSB_PI) * (y_over_w - norm_x0 + atan(1.0 / y_over_w));
      else
        view_dir_ground = 1.0 - 2.0 * norm_x0 / SSB_PI;
    }
  } else {
    const double weights[8] = {0.0506142681451884, 0.111190517226687, 0.156853322938944, 0.181341891689181,
                               0.181341891689181,  0.156853322938944, 0.111190517226687, 0.0506142681451884};
    const double nodes[8] = {0.0198550717512319, 0.101666761293187, 0.237233795041836, 0.408282678752175,
                             0.591717321247825,  0.762766204958164, 0.898333238706813, 0.980144928248768};
    double hw[8], vw[8], tk[8], ek[8];
    double sh = 0.0, sv = 0.0;
    for (int i = 0; i < 8; ++i) sh += weights[i] * nodes[i];
    for (int i = 0; i < 8; ++i) hw[i] = weights[i] * nodes[i] / sh;
    for (int i = 0; i < 8; ++i) vw[i] = weights[i] * sqrt(1.0 - nodes[i] * nodes[i]);
    for (int i = 0; i < 8; ++i) sv += vw[i];
    for (int i = 0; i < 8; ++i) vw[i] = vw[i] / sv;
    for (int i = 0; i < 8; ++i) {
      tk[i] = h * sqrt(1.0 / (nodes[i] * nodes[i]) - 1.0);
      ek[i] = exp(-tk[i]);
    }
    double s1 = 0.0, s2 = 0.0;
    for (int i = 0; i < 8; ++i) s1 += hw[i] * ek[i];
    for (int i = 0; i < 8; ++i) s2 += vw[i] * (1.0 - ek[i]) / tk[i];
    view_ground_sky = s1;
    view_wall_wall = 1.0 - s2;
    if (with_sun) view_dir_ground = exp(-(h * sqrt(1.0 / (cos_sza * cos_sza) - 1.0)));
  }
}

}  // namespace ssb
