// oracle_radsurf.cpp - CPU restatement of the reference's radsurf/ solver layer
// behind the same C structs as the product ABI (include/spartacus_b200.h).
// TEST INFRASTRUCTURE ONLY (see oracle_radtool.hpp header): "parity unpinned".
#include "oracle_radtool.hpp"
#include "../../include/spartacus_b200.h"
#include <cstdio>
#include <cstdlib>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace orc {

static const real kPi = 3.14159265358979323846; // radiation_constants.F90:24
static const real kEps = std::numeric_limits<double>::epsilon(); // epsilon(1.0_jprb): an algorithm constant, FP64 in every build

static LegendreGauss lg_from_c(const ssb200_legendre_gauss &c) {
  LegendreGauss lg;
  lg.nstream = c.nstream;
  const int n = c.nstream;
  lg.mu.assign(c.mu, c.mu + n);
  lg.sin_ang.assign(c.sin_ang, c.sin_ang + n);
  lg.tan_ang.assign(c.tan_ang, c.tan_ang + n);
  lg.weight.assign(c.weight, c.weight + n);
  lg.hweight.assign(c.hweight, c.hweight + n);
  lg.vweight.assign(c.vweight, c.vweight + n);
  lg.vadjustment = c.vadjustment;
  lg.vadjustment2 = c.vadjustment2;
  return lg;
}

// ---------------------------------------------------------------------------
// radsurf_overlap.F90
// ---------------------------------------------------------------------------

// calc_overlap_matrix_max_ran: radsurf_overlap.F90:28-73 (nreg 2 or 3).
static Mat overlap_max_ran(int nreg, real f_upper, real f_lower) {
  Mat O(nreg, nreg);
  const real pair_cover = rmax(f_upper, f_lower);
  O(0, 0) = 1.0 - pair_cover;
  if (nreg == 2) {
    O(0, 1) = pair_cover - f_upper;
    O(1, 0) = pair_cover - f_lower;
    O(1, 1) = f_upper + f_lower - pair_cover;
  } else if (nreg == 3) {
    O(0, 1) = 0.5 * (pair_cover - f_upper);
    O(0, 2) = O(0, 1);
    O(1, 0) = 0.5 * (pair_cover - f_lower);
    O(2, 0) = O(1, 0);
    O(1, 1) = 0.5 * (f_upper + f_lower - pair_cover);
    O(2, 2) = O(1, 1);
    O(1, 2) = 0.0;
    O(2, 1) = 0.0;
  }
  return O;
}

// calc_overlap_matrices: radsurf_overlap.F90:85-171.  frac is (nreg, nlay+1);
// outputs are per interface jlay = 0..nlay (0 = ground).
static void calc_overlap_matrices(int nlay, int nreg, const Mat &frac, std::vector<Mat> &u,
                                  std::vector<Mat> &v, real frac_threshold) {
  u.assign(nlay + 1, Mat(nreg, nreg));
  v.assign(nlay + 1, Mat(nreg, nreg));
  Vec frac_upper(nreg, 0.0), frac_lower(nreg, 0.0);
  frac_lower[0] = 1.0;
  for (int jlay = 0; jlay <= nlay; ++jlay) {
    if (jlay > nlay - 1) {
      frac_upper.assign(nreg, 0.0);
      frac_upper[0] = 1.0;
    } else {
      for (int r = 0; r < nreg; ++r) frac_upper[r] = frac(r, jlay);
    }
    Mat O = overlap_max_ran(nreg, 1.0 - frac_upper[0], 1.0 - frac_lower[0]);
    for (int jupper = 0; jupper < nreg; ++jupper)
      for (int jlower = 0; jlower < nreg; ++jlower) {
        if (frac_lower[jlower] >= frac_threshold)
          u[jlay](jupper, jlower) = O(jupper, jlower) / frac_lower[jlower];
        else
          u[jlay](jupper, jlower) = 0.0;
        if (frac_upper[jupper] >= frac_threshold)
          v[jlay](jlower, jupper) = O(jupper, jlower) / frac_upper[jupper];
        else
          v[jlay](jlower, jupper) = 0.0;
      }
    frac_lower = frac_upper;
  }
}

// calc_overlap_matrix_max_ran_urban: radsurf_overlap.F90:178-280.
static Mat overlap_max_ran_urban(int nreg, const Vec &fu, const Vec &fl) {
  Mat O(nreg, nreg + 1);
  if (nreg == 1) {
    O(0, 0) = fl[0];
    O(0, 1) = fl[1];
  } else if (nreg == 2) {
    const real pair_cover = rmax(fu[1], fl[1]);
    if (pair_cover <= fl[0] + fl[1]) {
      O(1, 2) = 0.0;
      O(0, 2) = fl[2];
      O(0, 0) = fl[0] + fl[1] - pair_cover;
      O(0, 1) = pair_cover - fu[1];
      O(1, 0) = pair_cover - fl[1];
      O(1, 1) = fu[1] + fl[1] - pair_cover;
    } else {
      O(0, 0) = 0.0;
      O(0, 1) = 0.0;
      O(1, 0) = fl[0];
      O(1, 1) = fl[1];
      O(1, 2) = fu[1] - fl[0] - fl[1];
      O(0, 2) = fu[0];
    }
  } else {
    O(1, 2) = 0.0;
    O(2, 1) = 0.0;
    const real pair_cover = rmax(fu[1] + fu[2], fl[1] + fl[2]);
    if (pair_cover <= fl[0] + fl[1] + fl[2]) {
      O(1, 3) = 0.0;
      O(2, 3) = 0.0;
      O(0, 3) = fl[3];
      O(0, 0) = fl[0] + fl[1] + fl[2] - pair_cover;
      if (pair_cover > fu[1] + fu[2]) {
        O(1, 0) = 0.0;
        O(2, 0) = 0.0;
        O(1, 1) = fu[1];
        O(2, 2) = fu[2];
        O(0, 1) = fl[1] - fu[1];
        O(0, 2) = fl[2] - fu[2];
      } else {
        O(0, 1) = 0.0;
        O(0, 2) = 0.0;
        O(1, 1) = fl[1];
        O(2, 2) = fl[2];
        O(1, 0) = fu[1] - fl[1];
        O(2, 0) = fu[2] - fl[2];
      }
    } else {
      O(0, 0) = 0.0;
      O(0, 1) = 0.0;
      O(0, 2) = 0.0;
      O(1, 1) = fl[1];
      O(2, 2) = fl[2];
      O(1, 0) = fl[0] * 0.5;
      O(2, 0) = O(0, 1); // sic (:268, App. B4)
      O(1, 3) = (fl[3] - fu[0]) * 0.5;
      O(2, 3) = O(1, 3);
      O(0, 3) = fu[0];
    }
  }
  return O;
}

// calc_overlap_matrices_urban: radsurf_overlap.F90:289-394.  frac is
// (nreg, nlay+1) but only columns 0..nlay-1 are read, like the dummy
// argument region_fracs(1:nreg,nlay).
static void calc_overlap_matrices_urban(int nlay, int nreg, const Mat &frac,
                                        std::vector<Mat> &u, std::vector<Mat> &v,
                                        real frac_threshold) {
  u.assign(nlay + 1, Mat(nreg, nreg + 1));
  v.assign(nlay + 1, Mat(nreg + 1, nreg));
  Vec frac_upper(nreg, 0.0), frac_lower(nreg + 1, 0.0);
  auto sumfrac = [&](int jl) {
    real s = 0.0;
    for (int r = 0; r < nreg; ++r) s += frac(r, jl);
    return s;
  };
  frac_lower[nreg] = sumfrac(0);
  for (int jlay = 0; jlay <= nlay; ++jlay) { // jlay+1 is the Fortran interface index
    if (jlay > nlay - 1) {
      frac_upper.assign(nreg, 0.0);
      frac_upper[0] = 1.0;
    } else {
      for (int r = 0; r < nreg; ++r) frac_upper[r] = frac(r, jlay);
    }
    Mat O = overlap_max_ran_urban(nreg, frac_upper, frac_lower);
    for (int jupper = 0; jupper < nreg; ++jupper)
      for (int jlower = 0; jlower < nreg + 1; ++jlower) {
        if (frac_lower[jlower] >= frac_threshold)
          u[jlay](jupper, jlower) = O(jupper, jlower) / frac_lower[jlower];
        else
          u[jlay](jupper, jlower) = 0.0;
        if (frac_upper[jupper] >= frac_threshold)
          v[jlay](jlower, jupper) = O(jupper, jlower) / frac_upper[jupper];
        else
          v[jlay](jlower, jupper) = 0.0;
      }
    for (int r = 0; r < nreg; ++r) frac_lower[r] = frac_upper[r];
    const int fj = jlay + 1; // Fortran jlay
    if (fj < nlay) {
      frac_lower[nreg] = sumfrac(fj) - sumfrac(fj - 1);
      if (frac_lower[nreg] < 0.0) {
        const real sc_num = sumfrac(fj), sc_den = sumfrac(fj - 1);
        for (int r = 0; r < nreg; ++r) frac_lower[r] = frac_lower[r] * sc_num / sc_den;
        frac_lower[nreg] = 0.0;
      }
    } else if (fj == nlay) {
      frac_lower[nreg] = 1.0 - sumfrac(fj - 1);
    }
  }
}

// ---------------------------------------------------------------------------
// radsurf_norm_perim.F90
// ---------------------------------------------------------------------------

// calc_norm_perim_forest: radsurf_norm_perim.F90:30-116.
static void calc_norm_perim_forest(const ssb200_config &cfg, int nlay, int nreg,
                                   const double *veg_fraction, const double *veg_scale,
                                   Mat &norm_perim) {
  norm_perim = Mat(nreg, nlay);
  const real fiso = cfg.vegetation_isolation_factor_forest;
  for (int jlay = 0; jlay < nlay; ++jlay) {
    if (nreg > 1) {
      if (veg_fraction[jlay] > cfg.min_vegetation_fraction) {
        if (cfg.use_symmetric_vegetation_scale_forest)
          norm_perim(0, jlay) = 4.0 * veg_fraction[jlay] *
                                rmax(0.0, 1.0 - veg_fraction[jlay]) / veg_scale[jlay];
        else
          norm_perim(0, jlay) = 4.0 * veg_fraction[jlay] / veg_scale[jlay];
        if (nreg > 2) {
          norm_perim(nreg - 1, jlay) = 0.5 * fiso * norm_perim(0, jlay);
          norm_perim(0, jlay) = (1.0 - 0.5 * fiso) * norm_perim(0, jlay);
          if (cfg.use_symmetric_vegetation_scale_forest)
            norm_perim(1, jlay) = (1.0 - fiso) * 4.0 * (0.5 * veg_fraction[jlay]) *
                                  (1.0 - (0.5 * veg_fraction[jlay])) / veg_scale[jlay];
          else
            norm_perim(1, jlay) =
                (1.0 - fiso) * 4.0 * veg_fraction[jlay] / (std::sqrt(2.0) * veg_scale[jlay]);
        } else {
          for (int r = 1; r < nreg; ++r) norm_perim(r, jlay) = 0.0;
        }
      }
    }
  }
}

// calc_norm_perim_urban: radsurf_norm_perim.F90:131-281.
static void calc_norm_perim_urban(const ssb200_config &cfg, int nlay, int nreg,
                                  const double *building_fraction,
                                  const double *building_scale, const double *veg_fraction,
                                  const double *veg_scale, const double *veg_contact_fraction,
                                  Mat &norm_perim, Mat &norm_perim_wall) {
  norm_perim = Mat(nreg, nlay);
  norm_perim_wall = Mat(nreg, nlay);
  const real fiso = cfg.vegetation_isolation_factor_urban;
  for (int jlay = 0; jlay < nlay; ++jlay) {
    if (nreg > 1) {
      if (veg_fraction[jlay] > cfg.min_vegetation_fraction) {
        if (cfg.use_symmetric_vegetation_scale_urban)
          norm_perim(0, jlay) =
              4.0 * veg_fraction[jlay] *
              rmax(0.0, 1.0 - veg_fraction[jlay] - building_fraction[jlay]) /
              (rmax(cfg.min_building_fraction, 1.0 - building_fraction[jlay]) *
               veg_scale[jlay]);
        else
          norm_perim(0, jlay) = 4.0 * veg_fraction[jlay] / veg_scale[jlay];
        if (nreg > 2) {
          norm_perim(nreg - 1, jlay) = 0.5 * fiso * norm_perim(0, jlay);
          norm_perim(0, jlay) = (1.0 - 0.5 * fiso) * norm_perim(0, jlay);
          if (cfg.use_symmetric_vegetation_scale_urban)
            norm_perim(1, jlay) =
                (1.0 - fiso) * 4.0 * (0.5 * veg_fraction[jlay]) *
                (1.0 - (0.5 * veg_fraction[jlay]) - building_fraction[jlay]) /
                (rmax(cfg.min_building_fraction, 1.0 - building_fraction[jlay]) *
                 veg_scale[jlay]);
          else
            norm_perim(1, jlay) =
                (1.0 - fiso) * 4.0 * veg_fraction[jlay] / (std::sqrt(2.0) * veg_scale[jlay]);
        } else {
          for (int r = 1; r < nreg; ++r) norm_perim(r, jlay) = 0.0;
        }
      }
    }
    if (building_fraction[jlay] > cfg.min_building_fraction) {
      norm_perim_wall(0, jlay) = 4.0 * building_fraction[jlay] / building_scale[jlay];
      if (nreg > 1) {
        if (1.0 - veg_fraction[jlay] - building_fraction[jlay] <= cfg.min_vegetation_fraction) {
          if (nreg == 2) {
            norm_perim_wall(1, jlay) = norm_perim_wall(0, jlay);
          } else {
            norm_perim_wall(1, jlay) = norm_perim_wall(0, jlay) * (1.0 - fiso);
            norm_perim_wall(2, jlay) = norm_perim_wall(0, jlay) * fiso;
          }
          norm_perim_wall(0, jlay) = 0.0;
        } else if (veg_fraction[jlay] > cfg.min_vegetation_fraction) {
          if (veg_contact_fraction[jlay] > 0.0) {
            if (nreg == 2) {
              norm_perim_wall(1, jlay) = norm_perim_wall(0, jlay) * veg_contact_fraction[jlay];
            } else {
              norm_perim_wall(1, jlay) =
                  norm_perim_wall(0, jlay) * veg_contact_fraction[jlay] * (1.0 - fiso);
              norm_perim_wall(2, jlay) =
                  norm_perim_wall(0, jlay) * veg_contact_fraction[jlay] * fiso;
            }
            norm_perim_wall(0, jlay) =
                norm_perim_wall(0, jlay) * (1.0 - veg_contact_fraction[jlay]);
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// radsurf_view_factor.F90
// ---------------------------------------------------------------------------

// calc_view_factors_inf: radsurf_view_factor.F90:28-70.
static void calc_view_factors_inf(real hw, real &view_ground_sky, real &view_wall_wall,
                                  const real *cos_sza, real *view_dir_ground) {
  view_ground_sky = std::sqrt(hw * hw + 1.0) - hw;
  view_wall_wall = std::sqrt(1.0 / (hw * hw) + 1.0) - 1.0 / hw;
  if (cos_sza && view_dir_ground) {
    const real c = *cos_sza;
    const real norm_x0 = (kPi * 0.5) * hw * std::sqrt(1.0 / (c * c) - 1.0);
    const real y_over_w = std::sqrt(rmax(norm_x0 * norm_x0 - 1.0, 0.0));
    if (y_over_w > 0.0)
      *view_dir_ground = (2.0 / kPi) * (y_over_w - norm_x0 + std::atan(1.0 / y_over_w));
    else
      *view_dir_ground = 1.0 - 2.0 * norm_x0 / kPi;
  }
}

// calc_view_factors_exp: radsurf_view_factor.F90:76-138.
static void calc_view_factors_exp(real hx, real &view_ground_sky, real &view_wall_wall,
                                  const real *cos_sza, real *view_dir_ground) {
  static const int nw = 8;
  static const real weights[nw] = {0.0506142681451884, 0.111190517226687, 0.156853322938944,
                                     0.181341891689181,  0.181341891689181, 0.156853322938944,
                                     0.111190517226687,  0.0506142681451884};
  static const real nodes[nw] = {0.0198550717512319, 0.101666761293187, 0.237233795041836,
                                   0.408282678752175,  0.591717321247825, 0.762766204958164,
                                   0.898333238706813,  0.980144928248768};
  real hweight[nw], vweight[nw], tk[nw], exp_tk[nw];
  real sh = 0.0, sv = 0.0;
  for (int i = 0; i < nw; ++i) sh += weights[i] * nodes[i];
  for (int i = 0; i < nw; ++i) hweight[i] = weights[i] * nodes[i] / sh;
  for (int i = 0; i < nw; ++i) vweight[i] = weights[i] * std::sqrt(1.0 - nodes[i] * nodes[i]);
  for (int i = 0; i < nw; ++i) sv += vweight[i];
  for (int i = 0; i < nw; ++i) vweight[i] = vweight[i] / sv;
  for (int i = 0; i < nw; ++i) {
    tk[i] = hx * std::sqrt(1.0 / (nodes[i] * nodes[i]) - 1.0);
    exp_tk[i] = std::exp(-tk[i]);
  }
  real s1 = 0.0, s2 = 0.0;
  for (int i = 0; i < nw; ++i) s1 += hweight[i] * exp_tk[i];
  for (int i = 0; i < nw; ++i) s2 += vweight[i] * (1.0 - exp_tk[i]) / tk[i];
  view_ground_sky = s1;
  view_wall_wall = 1.0 - s2;
  if (cos_sza && view_dir_ground) {
    const real c = *cos_sza;
    const real norm_x0 = hx * std::sqrt(1.0 / (c * c) - 1.0);
    *view_dir_ground = std::exp(-norm_x0);
  }
}

// ---------------------------------------------------------------------------
// canopy_flux_type%zero: radsurf_canopy_flux.F90:286-341
// ---------------------------------------------------------------------------
static void flux_zero(ssb200_canopy_flux *f, int icol, int ilay1, int ilay2) {
  // icol, ilay1, ilay2 are 0-based here; ilay2 inclusive
  if (!f) return;
  const int ns = f->nspec;
  auto zc = [&](double *p) {
    if (p)
      for (int g = 0; g < ns; ++g) p[g + (size_t)ns * icol] = 0.0;
  };
  auto zl = [&](double *p) {
    if (p)
      for (int l = ilay1; l <= ilay2; ++l)
        for (int g = 0; g < ns; ++g) p[g + (size_t)ns * l] = 0.0;
  };
  auto zl1 = [&](double *p) {
    if (p)
      for (int l = ilay1; l <= ilay2; ++l) p[l] = 0.0;
  };
  zc(f->ground_dn);
  zc(f->ground_net);
  zc(f->ground_vertical_diff);
  zc(f->top_dn);
  zc(f->top_net);
  if (f->ground_dn_dir) {
    zc(f->ground_dn_dir);
    zc(f->top_dn_dir);
    if (f->ground_sunlit_frac) f->ground_sunlit_frac[icol] = 0.0;
  }
  if (ilay2 >= ilay1) {
    if (f->roof_in) {
      zl(f->roof_in);
      zl(f->roof_net);
      zl(f->wall_in);
      zl(f->wall_net);
    }
    if (f->roof_in_dir) {
      zl(f->roof_in_dir);
      zl(f->wall_in_dir);
      zl1(f->roof_sunlit_frac);
      zl1(f->wall_sunlit_frac);
    }
    zl(f->clear_air_abs);
    if (f->veg_abs) {
      zl(f->veg_abs);
      zl(f->veg_air_abs);
    }
    if (f->veg_abs_dir) {
      zl(f->veg_abs_dir);
      zl1(f->veg_sunlit_frac);
    }
    if (f->flux_dn_layer_top) {
      zl(f->flux_dn_layer_top);
      zl(f->flux_up_layer_top);
      zl(f->flux_dn_layer_base);
      zl(f->flux_up_layer_base);
    }
    if (f->flux_dn_dir_layer_top) {
      zl(f->flux_dn_dir_layer_top);
      zl(f->flux_dn_dir_layer_base);
    }
  }
}

static inline real vsum(const Vec &v, int i0, int n) {
  real s = 0.0;
  for (int i = i0; i < i0 + n; ++i) s += v[i];
  return s;
}

// Common region property set-up (urban_sw:342-369, urban_lw:303-352 and the
// forest equivalents).
struct Ctx {
  const ssb200_config *cfg;
  const ssb200_canopy_properties *cp;
};

// ---------------------------------------------------------------------------
// spartacus_urban_sw (radsurf_urban_sw.F90:35-1007) and
// spartacus_forest_sw (radsurf_forest_sw.F90:35-783); `urban` selects which.
// icol and ilay1 are 0-based.
// ---------------------------------------------------------------------------
static void spartacus_sw(bool urban, const ssb200_config &cfg, int nsw, int ns, int nreg,
                         int nlay, int icol, int ilay1, const LegendreGauss &lg, real cos_sza,
                         const ssb200_canopy_properties &cp,
                         const ssb200_sw_spectral_properties &sw, const double *ground_albedo_diff,
                         const double *ground_albedo_dir, double *top_albedo_diff,
                         double *top_albedo_dir, ssb200_canopy_flux *ndir,
                         ssb200_canopy_flux *ndiff) {
  const int n = nreg * ns;
  const int nrb = urban ? nreg + 1 : nreg; // regions just below an interface
  const int m = nrb * ns;
  const int ilay2 = ilay1 + nlay - 1;
  const double *dz = cp.dz + ilay1;
  const double *building_fraction = urban ? cp.building_fraction + ilay1 : nullptr;
  const double *veg_fraction = cp.veg_fraction ? cp.veg_fraction + ilay1 : nullptr;
  const double *veg_ext = cp.veg_ext ? cp.veg_ext + ilay1 : nullptr;
  const double *veg_fsd = cp.veg_fsd ? cp.veg_fsd + ilay1 : nullptr;
#define SP(arr, g, jlay) (sw.arr[(g) + (size_t)nsw * (ilay1 + (jlay))])
  // urban_sw:268 clamps; forest_sw uses the raw value except inside tan0 (App. B7)
  const real zcos_sza = urban ? rmax(cos_sza, 1.0e-6) : cos_sza;
  const bool do_vegetation = (nreg > 1);
  real sin0 = 0.0, tan0;
  if (urban) {
    sin0 = std::sqrt(1.0 - zcos_sza * zcos_sza);
    tan0 = sin0 / zcos_sza;
  } else {
    tan0 = std::sqrt(1.0 - cos_sza * cos_sza) / rmax(cos_sza, 1.0e-6);
  }

  // Region fractions (urban_sw:284-291, forest_sw:244-248)
  Mat frac(nreg, nlay + 1);
  if (urban) {
    for (int j = 0; j < nlay; ++j) frac(0, j) = 1.0 - building_fraction[j];
    frac(0, nlay) = 1.0;
    if (do_vegetation) {
      for (int j = 0; j < nlay; ++j) {
        frac(0, j) = rmax(0.0, frac(0, j) - veg_fraction[j]);
        const real fv = rmax(0.0, 1.0 - building_fraction[j] - frac(0, j)) / (real)(nreg - 1);
        for (int r = 1; r < nreg; ++r) frac(r, j) = fv;
      }
      for (int r = 1; r < nreg; ++r) frac(r, nlay) = 0.0;
    }
  } else {
    for (int j = 0; j < nlay; ++j) {
      frac(0, j) = 1.0 - veg_fraction[j];
      for (int r = 1; r < nreg; ++r) frac(r, j) = veg_fraction[j] / (real)(nreg - 1);
    }
    frac(0, nlay) = 1.0;
    for (int r = 1; r < nreg; ++r) frac(r, nlay) = 0.0;
  }
  Vec roof_fraction(nlay + 1, 0.0), non_building_fraction(nlay + 1, 1.0);
  if (urban) { // urban_sw:293-298
    roof_fraction[nlay] = 0.0;
    roof_fraction[nlay - 1] = building_fraction[nlay - 1];
    for (int j = 0; j < nlay - 1; ++j)
      roof_fraction[j] = rmax(0.0, building_fraction[j] - building_fraction[j + 1]);
    non_building_fraction[nlay] = 1.0;
    for (int j = 0; j < nlay; ++j) non_building_fraction[j] = 1.0 - building_fraction[j];
  }

  // Most transparent interval (urban_sw:310): first minimum of column OD
  int itransp = 0;
  {
    real best = 0.0;
    for (int g = 0; g < nsw; ++g) {
      real od = 0.0;
      for (int j = 0; j < nlay; ++j) od += SP(air_ext, g, j) * dz[j];
      if (g == 0 || od < best) {
        best = od;
        itransp = g;
      }
    }
  }

  std::vector<Mat> u_overlap, v_overlap;
  Mat norm_perim, norm_perim_wall;
  if (urban) {
    calc_overlap_matrices_urban(nlay, nreg, frac, u_overlap, v_overlap,
                                cfg.min_vegetation_fraction);
    calc_norm_perim_urban(cfg, nlay, nreg, cp.building_fraction + ilay1,
                          cp.building_scale + ilay1, cp.veg_fraction ? cp.veg_fraction + ilay1 : nullptr,
                          cp.veg_scale ? cp.veg_scale + ilay1 : nullptr,
                          cp.veg_contact_fraction ? cp.veg_contact_fraction + ilay1 : nullptr,
                          norm_perim, norm_perim_wall);
  } else {
    calc_overlap_matrices(nlay, nreg, frac, u_overlap, v_overlap, cfg.min_vegetation_fraction);
    calc_norm_perim_forest(cfg, nlay, nreg, cp.veg_fraction + ilay1, cp.veg_scale + ilay1,
                           norm_perim);
  }

  // Per-layer, per-interval matrices
  auto L = [&](int g, int j) { return (size_t)g + (size_t)nsw * j; };
  std::vector<Mat> trans_diff(nsw * nlay, Mat(n, n)), ref_diff(nsw * nlay, Mat(n, n));
  std::vector<Mat> ref_dir(nsw * nlay, Mat(n, nreg)), trans_dir_diff(nsw * nlay, Mat(n, nreg));
  std::vector<Mat> trans_dir_dir(nsw * nlay, Mat(nreg, nreg));
  std::vector<Mat> int_diff(nsw * nlay, Mat(n, n)), int_dir(nsw * nlay, Mat(nreg, nreg));
  std::vector<Mat> int_dir_diff(nsw * nlay, Mat(n, nreg));
  Mat f_wall(nreg, nlay), od_scaling(nreg, nlay);
  Vec f_wall_dir_clear(nlay, 0.0);

  for (int jlay = 0; jlay < nlay; ++jlay) {
    // Exchange rates are spectrally independent (urban_sw:373-410)
    Mat f_exchange(nreg, nreg);
    int jreg;
    for (jreg = 0; jreg < nreg - 1; ++jreg) {
      if (frac(jreg, jlay) <= cfg.min_vegetation_fraction ||
          frac(jreg + 1, jlay) <= cfg.min_vegetation_fraction) {
        f_exchange(jreg + 1, jreg) = 0.0;
        f_exchange(jreg, jreg + 1) = 0.0;
      } else {
        f_exchange(jreg + 1, jreg) = norm_perim(jreg, jlay) / (kPi * frac(jreg, jlay));
        f_exchange(jreg, jreg + 1) = norm_perim(jreg, jlay) / (kPi * frac(jreg + 1, jlay));
      }
    }
    // After the loop the Fortran jreg == nreg (1-based), i.e. index nreg-1
    // here (App. B3); for nreg < 2 the loop body never runs.
    const int jreg_exit = std::max(nreg - 1, 0);
    if (nreg > 2 && norm_perim(nreg - 1, jlay) > 0.0) {
      if (frac(2, jlay) <= cfg.min_vegetation_fraction ||
          frac(0, jlay) <= cfg.min_vegetation_fraction) {
        f_exchange(0, 2) = 0.0;
        f_exchange(2, 0) = 0.0;
      } else {
        f_exchange(0, 2) = norm_perim(jreg_exit, jlay) / (kPi * frac(2, jlay));
        f_exchange(2, 0) = norm_perim(jreg_exit, jlay) / (kPi * frac(0, jlay));
      }
    }
    if (urban) {
      for (jreg = 0; jreg < nreg; ++jreg) {
        if (frac(jreg, jlay) <= cfg.min_vegetation_fraction)
          f_wall(jreg, jlay) = 0.0;
        else
          f_wall(jreg, jlay) = norm_perim_wall(jreg, jlay) / (kPi * frac(jreg, jlay));
      }
      if (non_building_fraction[jlay] <= cfg.min_building_fraction) {
        f_wall_dir_clear[jlay] = 0.0;
      } else {
        real s = 0.0;
        for (jreg = 0; jreg < nreg; ++jreg) s += norm_perim_wall(jreg, jlay);
        f_wall_dir_clear[jlay] = s / (kPi * non_building_fraction[jlay]);
      }
    }
    if (nreg == 2) {
      od_scaling(1, jlay) = 1.0;
    } else if (nreg == 3) {
      od_scaling(1, jlay) =
          std::exp(-veg_fsd[jlay] * (1.0 + 0.5 * veg_fsd[jlay] * (1.0 + 0.5 * veg_fsd[jlay])));
      od_scaling(2, jlay) = 2.0 - od_scaling(1, jlay);
    }

    for (int g = 0; g < nsw; ++g) {
      // Section 3a (urban_sw:342-369)
      Vec ext_reg(nreg), ssa_reg(nreg);
      ext_reg[0] = SP(air_ext, g, jlay);
      ssa_reg[0] = SP(air_ssa, g, jlay);
      if (nreg == 2) {
        ext_reg[1] = SP(air_ext, g, jlay) + veg_ext[jlay];
        ssa_reg[1] = (ext_reg[0] * ssa_reg[0] + veg_ext[jlay] * SP(veg_ssa, g, jlay)) /
                     rmax(ext_reg[1], 1.0e-8);
      } else if (nreg == 3) {
        ext_reg[1] = SP(air_ext, g, jlay) + od_scaling(1, jlay) * veg_ext[jlay];
        ext_reg[2] = SP(air_ext, g, jlay) + od_scaling(2, jlay) * veg_ext[jlay];
        ssa_reg[1] = (ext_reg[0] * ssa_reg[0] +
                      od_scaling(1, jlay) * veg_ext[jlay] * SP(veg_ssa, g, jlay)) /
                     rmax(ext_reg[1], 1.0e-8);
        ssa_reg[2] = (ext_reg[0] * ssa_reg[0] +
                      od_scaling(2, jlay) * veg_ext[jlay] * SP(veg_ssa, g, jlay)) /
                     rmax(ext_reg[2], 1.0e-8);
      }
      real wall_ext = 0.0, wall_factor = 0.0;
      if (urban) { // urban_sw:414-418
        wall_ext = 1.0 - SP(wall_albedo, g, jlay) * SP(wall_specular_frac, g, jlay);
        wall_factor = SP(wall_albedo, g, jlay) * (1.0 - SP(wall_specular_frac, g, jlay));
      }

      // Section 3b (urban_sw:426-494; forest_sw:326-372)
      Mat gamma0(nreg, nreg), gamma1(n, n), gamma2(n, n), gamma3(n, nreg);
      for (int jreg_fr = 0; jreg_fr < nreg; ++jreg_fr)
        for (int jreg_to = 0; jreg_to < nreg; ++jreg_to)
          if (jreg_fr != jreg_to) {
            gamma0(jreg_fr, jreg_fr) = gamma0(jreg_fr, jreg_fr) - tan0 * f_exchange(jreg_to, jreg_fr);
            gamma0(jreg_to, jreg_fr) = +tan0 * f_exchange(jreg_to, jreg_fr);
            for (int js = 0; js < ns; ++js) {
              const int ifr = js + jreg_fr * ns, ito = js + jreg_to * ns;
              gamma1(ifr, ifr) = gamma1(ifr, ifr) - lg.tan_ang[js] * f_exchange(jreg_to, jreg_fr);
              gamma1(ito, ifr) = +lg.tan_ang[js] * f_exchange(jreg_to, jreg_fr);
            }
          }
      for (jreg = 0; jreg < nreg; ++jreg) {
        if (urban)
          gamma0(jreg, jreg) = gamma0(jreg, jreg) - ext_reg[jreg] / zcos_sza -
                               tan0 * f_wall(jreg, jlay) * wall_ext;
        else
          gamma0(jreg, jreg) = gamma0(jreg, jreg) - ext_reg[jreg] / cos_sza;
        for (int js = 0; js < ns; ++js) {
          const int ifr = js + jreg * ns;
          if (urban)
            gamma1(ifr, ifr) = gamma1(ifr, ifr) - ext_reg[jreg] / lg.mu[js] -
                               lg.tan_ang[js] * f_wall(jreg, jlay) * wall_ext;
          else
            gamma1(ifr, ifr) = gamma1(ifr, ifr) - ext_reg[jreg] / lg.mu[js];
        }
      }
      for (int js_fr = 0; js_fr < ns; ++js_fr)
        for (int js_to = 0; js_to < ns; ++js_to)
          for (jreg = 0; jreg < nreg; ++jreg) {
            const int ifr = js_fr + jreg * ns, ito = js_to + jreg * ns;
            if (urban)
              gamma2(ito, ifr) =
                  0.5 * (lg.weight[js_to] * ext_reg[jreg] * ssa_reg[jreg] / lg.mu[js_fr] +
                         lg.vweight[js_to] * lg.tan_ang[js_fr] * f_wall(jreg, jlay) * wall_factor);
            else
              gamma2(ito, ifr) =
                  0.5 * lg.weight[js_to] * ext_reg[jreg] * ssa_reg[jreg] / lg.mu[js_fr];
          }
      gamma1 = gamma1 + gamma2;
      for (jreg = 0; jreg < nreg; ++jreg)
        for (int js = 0; js < ns; ++js) {
          const int ito = js + jreg * ns;
          if (urban)
            gamma3(ito, jreg) = 0.5 * (lg.weight[js] * ext_reg[jreg] * ssa_reg[jreg] +
                                       lg.vweight[js] * sin0 * f_wall(jreg, jlay) * wall_factor);
          else
            gamma3(ito, jreg) = 0.5 * lg.weight[js] * ext_reg[jreg] * ssa_reg[jreg];
        }

      // Section 3c (urban_sw:512-583; forest_sw:382-431)
      const size_t k = L(g, jlay);
      int r0, nr; // first region and number of regions solved
      bool veg_branching = urban ? do_vegetation : true;
      if (veg_branching) {
        if (veg_fraction[jlay] <= cfg.min_vegetation_fraction) {
          r0 = 0;
          nr = 1;
        } else if (frac(0, jlay) <= cfg.min_vegetation_fraction) {
          r0 = 1;
          nr = nreg - 1;
        } else {
          r0 = 0;
          nr = nreg;
        }
      } else {
        r0 = 0;
        nr = 1;
      }
      const int i0 = r0 * ns, nn = nr * ns;
      Mat R, T, Su, Sd, E, Idir, Idiff, Idd;
      calc_matrices_sw_eig(nn, nr, dz[jlay], zcos_sza, sub(gamma0, r0, nr, r0, nr),
                           sub(gamma1, i0, nn, i0, nn), sub(gamma2, i0, nn, i0, nn),
                           sub(gamma3, i0, nn, r0, nr), R, T, Su, Sd, E, Idir, Idiff, Idd);
      paste(ref_diff[k], i0, i0, R);
      paste(trans_diff[k], i0, i0, T);
      paste(ref_dir[k], i0, r0, Su);
      paste(trans_dir_diff[k], i0, r0, Sd);
      paste(trans_dir_dir[k], r0, r0, E);
      paste(int_dir[k], r0, r0, Idir);
      paste(int_diff[k], i0, i0, Idiff);
      paste(int_dir_diff[k], i0, r0, Idd);
    }
  }

  // Section 4: albedo of scene at each interface (urban_sw:591-654)
  std::vector<Mat> a_above(nsw * (nlay + 1), Mat(n, n)), d_above(nsw * (nlay + 1), Mat(n, nreg));
  std::vector<Mat> a_below(nsw * (nlay + 1), Mat(m, m)), d_below(nsw * (nlay + 1), Mat(m, nrb));
  std::vector<Mat> denominator(nsw * nlay, Mat(n, n));
  for (int g = 0; g < nsw; ++g)
    for (int jreg = 0; jreg < nreg; ++jreg)
      for (int js_to = 0; js_to < ns; ++js_to) {
        d_above[L(g, 0)](js_to + jreg * ns, jreg) = zcos_sza * ground_albedo_dir[g] * lg.hweight[js_to];
        for (int js_fr = 0; js_fr < ns; ++js_fr)
          a_above[L(g, 0)](js_to + jreg * ns, js_fr + jreg * ns) =
              ground_albedo_diff[g] * lg.hweight[js_to];
      }
  for (int jlay = 0; jlay < nlay; ++jlay)
    for (int g = 0; g < nsw; ++g) {
      const size_t k = L(g, jlay), k1 = L(g, jlay + 1);
      denominator[k] = identity_minus_mat_x_mat(a_above[k], ref_diff[k]);
      Mat ab = ref_diff[k] + matmul(trans_diff[k], solve_mat(denominator[k], matmul(a_above[k], trans_diff[k])));
      Mat db = ref_dir[k] + matmul(trans_diff[k],
                                   solve_rect_mat(denominator[k], matmul(d_above[k], trans_dir_dir[k]) +
                                                                      matmul(a_above[k], trans_dir_diff[k])));
      paste(a_below[k1], 0, 0, ab);
      paste(d_below[k1], 0, 0, db);
      if (urban) {
        for (int js = 0; js < ns; ++js)
          for (int j2 = 0; j2 < ns; ++j2)
            a_below[k1](n + js, n + j2) = SP(roof_albedo, g, jlay) * lg.hweight[js];
        for (int js = 0; js < ns; ++js) {
          if (sw.roof_albedo_dir)
            d_below[k1](n + js, nreg) = zcos_sza * SP(roof_albedo_dir, g, jlay) * lg.hweight[js];
          else
            d_below[k1](n + js, nreg) = zcos_sza * SP(roof_albedo, g, jlay) * lg.hweight[js];
        }
      }
      a_above[k1] = expandedmat_x_mat(nreg, nrb, ns, u_overlap[jlay + 1],
                                      mat_x_expandedmat(nrb, nreg, ns, a_below[k1], v_overlap[jlay + 1]));
      d_above[k1] = expandedmat_x_mat(nreg, nrb, ns, u_overlap[jlay + 1],
                                      matmul(d_below[k1], v_overlap[jlay + 1]));
    }

  // Top-of-canopy boundary conditions (urban_sw:672-674)
  Vec talb_dir(nsw), talb_diff(nsw);
  for (int g = 0; g < nsw; ++g) {
    Vec y = matvec(sub(a_above[L(g, nlay)], 0, ns, 0, ns), lg.hweight);
    talb_diff[g] = vsum(y, 0, ns);
    real s = 0.0;
    for (int js = 0; js < ns; ++js) s += d_above[L(g, nlay)](js, 0);
    talb_dir[g] = s / zcos_sza;
    top_albedo_diff[g] = talb_diff[g];
    top_albedo_dir[g] = talb_dir[g];
  }

  // Section 5 (urban_sw:681-...)
  flux_zero(ndiff, icol, ilay1, ilay2);
  flux_zero(ndir, icol, ilay1, ilay2);
#define FC(f, member, g) (f->member[(g) + (size_t)nsw * icol])
#define FL(f, member, g, il) (f->member[(g) + (size_t)nsw * (il)])

  std::vector<Vec> dn_dir_above(nsw, Vec(nreg, 0.0)), dn_diff_above(nsw, Vec(n, 0.0)), up_above(nsw, Vec(n, 0.0));
  for (int g = 0; g < nsw; ++g) {
    dn_dir_above[g][0] = 1.0 / zcos_sza;
    FC(ndir, top_dn_dir, g) = 1.0;
    FC(ndir, top_dn, g) = FC(ndir, top_dn_dir, g);
    FC(ndir, top_net, g) = FC(ndir, top_dn_dir, g) * (1.0 - talb_dir[g]);
  }
  if (urban && ndir->roof_sunlit_frac) ndir->roof_sunlit_frac[ilay2] = 1.0;
  real flux_dn_dir_clear = 1.0 / zcos_sza;

  for (int jlay = nlay - 1; jlay >= 0; --jlay) {
    const int ilay = ilay1 + jlay;
    for (int g = 0; g < nsw; ++g) {
      const size_t k = L(g, jlay), k1 = L(g, jlay + 1);
      Vec dn_dir_below = matvec(v_overlap[jlay + 1], dn_dir_above[g]);
      Vec dn_diff_below = expandedmat_x_vec(nrb, nreg, ns, v_overlap[jlay + 1], dn_diff_above[g]);
      Vec up_below = matvec(a_below[k1], dn_diff_below) + matvec(d_below[k1], dn_dir_below);
      if (urban) {
        FL(ndir, roof_in_dir, g, ilay) = zcos_sza * dn_dir_below[nreg];
        FL(ndir, roof_in, g, ilay) = FL(ndir, roof_in_dir, g, ilay) + vsum(dn_diff_below, n, ns);
        FL(ndir, roof_net, g, ilay) = FL(ndir, roof_in, g, ilay) - vsum(up_below, n, ns);
      }
      Vec dir_b(dn_dir_below.begin(), dn_dir_below.begin() + nreg);
      Vec diff_b(dn_diff_below.begin(), dn_diff_below.begin() + n);
      Vec up_b(up_below.begin(), up_below.begin() + n);
      dn_dir_above[g] = matvec(trans_dir_dir[k], dir_b);
      Vec reflected = matvec(d_above[k], dn_dir_above[g]);
      dn_diff_above[g] = solve_vec(denominator[k], matvec(trans_diff[k], diff_b) +
                                                       matvec(ref_diff[k], reflected) +
                                                       matvec(trans_dir_diff[k], dir_b));
      up_above[g] = matvec(a_above[k], dn_diff_above[g]) + reflected;

      if (ndir->flux_dn_layer_top) {
        FL(ndir, flux_dn_dir_layer_top, g, ilay) = zcos_sza * vsum(dir_b, 0, nreg);
        FL(ndir, flux_dn_layer_top, g, ilay) = FL(ndir, flux_dn_dir_layer_top, g, ilay) + vsum(diff_b, 0, n);
        FL(ndir, flux_up_layer_top, g, ilay) = vsum(up_b, 0, n);
        FL(ndir, flux_dn_dir_layer_base, g, ilay) = zcos_sza * vsum(dn_dir_above[g], 0, nreg);
        FL(ndir, flux_dn_layer_base, g, ilay) =
            FL(ndir, flux_dn_dir_layer_base, g, ilay) + vsum(dn_diff_above[g], 0, n);
        FL(ndir, flux_up_layer_base, g, ilay) = vsum(up_above[g], 0, n);
      }

      Vec int_flux_dir = matvec(int_dir[k], dir_b - dn_dir_above[g]);
      Vec conv(n);
      for (int i = 0; i < n; ++i) conv[i] = diff_b[i] - dn_diff_above[g][i] - up_b[i] + up_above[g][i];
      Vec int_flux_diff = matvec(int_diff[k], conv) + matvec(int_dir_diff[k], dir_b - dn_dir_above[g]);

      auto sum_over_mu = [&](int jreg) {
        real s = 0.0;
        for (int js = 0; js < ns; ++js) s += int_flux_diff[jreg * ns + js] * (1.0 / lg.mu[js]);
        return s;
      };
      auto sum_tan = [&](int jreg) {
        real s = 0.0;
        for (int js = 0; js < ns; ++js) s += int_flux_diff[jreg * ns + js] * lg.tan_ang[js];
        return s;
      };
      const real air_abs = SP(air_ext, g, jlay) * (1.0 - SP(air_ssa, g, jlay));
      FL(ndir, clear_air_abs, g, ilay) =
          FL(ndir, clear_air_abs, g, ilay) + air_abs * (int_flux_dir[0] + sum_over_mu(0));
      if (do_vegetation) {
        for (int jreg = 1; jreg < nreg; ++jreg) {
          const real vabs = veg_ext[jlay] * (1.0 - SP(veg_ssa, g, jlay));
          FL(ndir, veg_air_abs, g, ilay) =
              FL(ndir, veg_air_abs, g, ilay) + air_abs * (int_flux_dir[jreg] + sum_over_mu(jreg));
          FL(ndir, veg_abs_dir, g, ilay) =
              FL(ndir, veg_abs_dir, g, ilay) + vabs * int_flux_dir[jreg] * od_scaling(jreg, jlay);
          FL(ndir, veg_abs, g, ilay) =
              FL(ndir, veg_abs, g, ilay) +
              vabs * (int_flux_dir[jreg] + sum_over_mu(jreg)) * od_scaling(jreg, jlay);
        }
      }
      if (urban) {
        for (int jreg = 0; jreg < nreg; ++jreg)
          FL(ndir, wall_in_dir, g, ilay) =
              FL(ndir, wall_in_dir, g, ilay) + f_wall(jreg, jlay) * sin0 * int_flux_dir[jreg];
        FL(ndir, wall_in, g, ilay) = FL(ndir, wall_in_dir, g, ilay);
        for (int jreg = 0; jreg < nreg; ++jreg)
          FL(ndir, wall_in, g, ilay) = FL(ndir, wall_in, g, ilay) + f_wall(jreg, jlay) * sum_tan(jreg);
        FL(ndir, wall_net, g, ilay) = FL(ndir, wall_in, g, ilay) * (1.0 - SP(wall_albedo, g, jlay));
      }
    }
    // Spectrally independent diagnostics (urban_sw:805-848; forest_sw:...)
    if (urban) {
      ndir->roof_sunlit_frac[ilay] =
          FL(ndir, roof_in_dir, itransp, ilay) * non_building_fraction[jlay + 1] /
          (zcos_sza * flux_dn_dir_clear * rmax(cfg.min_building_fraction, roof_fraction[jlay]));
      flux_dn_dir_clear =
          flux_dn_dir_clear * non_building_fraction[jlay] / non_building_fraction[jlay + 1];
    }
    const real trans_dir_clear = std::exp(-SP(air_ext, itransp, jlay) * dz[jlay] / zcos_sza);
    real int_flux_dir_clear;
    if (SP(air_ext, itransp, jlay) > 0.0)
      int_flux_dir_clear =
          flux_dn_dir_clear * (1.0 - trans_dir_clear) * zcos_sza / SP(air_ext, itransp, jlay);
    else
      int_flux_dir_clear = flux_dn_dir_clear * dz[jlay];
    if (urban ? do_vegetation : true) {
      // forest computes this unconditionally (forest_sw: veg_sunlit_frac)
      if (veg_ext && veg_fraction && sw.veg_ssa && ndir->veg_sunlit_frac) {
        const real veg_abs_dir_clear = int_flux_dir_clear * veg_ext[jlay] *
                                         (1.0 - SP(veg_ssa, itransp, jlay)) * veg_fraction[jlay];
        ndir->veg_sunlit_frac[ilay] =
            FL(ndir, veg_abs_dir, itransp, ilay) / rmax(kEps, veg_abs_dir_clear);
      }
    }
    if (urban)
      ndir->wall_sunlit_frac[ilay] =
          0.5 * FL(ndir, wall_in_dir, itransp, ilay) /
          rmax(kEps, (f_wall_dir_clear[jlay] * sin0 * int_flux_dir_clear));
    flux_dn_dir_clear = flux_dn_dir_clear * trans_dir_clear;
  }
  for (int g = 0; g < nsw; ++g) {
    FC(ndir, ground_dn_dir, g) = zcos_sza * vsum(dn_dir_above[g], 0, nreg);
    FC(ndir, ground_dn, g) = FC(ndir, ground_dn_dir, g) + vsum(dn_diff_above[g], 0, n);
    FC(ndir, ground_net, g) = FC(ndir, ground_dn, g) - vsum(up_above[g], 0, n);
    for (int jreg = 0; jreg < nreg; ++jreg)
      for (int js = 0; js < ns; ++js) {
        const int ifr = js + jreg * ns;
        FC(ndir, ground_vertical_diff, g) =
            FC(ndir, ground_vertical_diff, g) +
            (dn_diff_above[g][ifr] + up_above[g][ifr]) * lg.tan_ang[js] / kPi;
      }
  }
  ndir->ground_sunlit_frac[icol] = FC(ndir, ground_dn_dir, itransp) / (zcos_sza * flux_dn_dir_clear);

  // Diffuse source at canopy top (urban_sw:884-984)
  for (int g = 0; g < nsw; ++g) {
    dn_dir_above[g].assign(nreg, 0.0);
    dn_diff_above[g].assign(n, 0.0);
    for (int js = 0; js < ns; ++js) dn_diff_above[g][js] = lg.hweight[js];
    FC(ndiff, top_dn_dir, g) = 0.0;
    FC(ndiff, top_dn, g) = 1.0;
    FC(ndiff, top_net, g) = 1.0 - talb_diff[g];
  }
  for (int jlay = nlay - 1; jlay >= 0; --jlay) {
    const int ilay = ilay1 + jlay;
    for (int g = 0; g < nsw; ++g) {
      const size_t k = L(g, jlay), k1 = L(g, jlay + 1);
      Vec dn_diff_below = expandedmat_x_vec(nrb, nreg, ns, v_overlap[jlay + 1], dn_diff_above[g]);
      Vec up_below = matvec(a_below[k1], dn_diff_below);
      if (urban) {
        FL(ndiff, roof_in, g, ilay) = +vsum(dn_diff_below, n, ns);
        FL(ndiff, roof_net, g, ilay) = FL(ndiff, roof_in, g, ilay) - vsum(up_below, n, ns);
      }
      Vec diff_b(dn_diff_below.begin(), dn_diff_below.begin() + n);
      Vec up_b(up_below.begin(), up_below.begin() + n);
      dn_diff_above[g] = solve_vec(denominator[k], matvec(trans_diff[k], diff_b));
      up_above[g] = matvec(a_above[k], dn_diff_above[g]);
      if (ndiff->flux_dn_layer_top) {
        FL(ndiff, flux_dn_layer_top, g, ilay) = vsum(diff_b, 0, n);
        FL(ndiff, flux_up_layer_top, g, ilay) = vsum(up_b, 0, n);
        FL(ndiff, flux_dn_layer_base, g, ilay) = vsum(dn_diff_above[g], 0, n);
        FL(ndiff, flux_up_layer_base, g, ilay) = vsum(up_above[g], 0, n);
      }
      Vec conv(n);
      for (int i = 0; i < n; ++i) conv[i] = diff_b[i] - dn_diff_above[g][i] - up_b[i] + up_above[g][i];
      Vec int_flux_diff = matvec(int_diff[k], conv);
      auto sum_over_mu = [&](int jreg) {
        real s = 0.0;
        for (int js = 0; js < ns; ++js) s += int_flux_diff[jreg * ns + js] * (1.0 / lg.mu[js]);
        return s;
      };
      auto sum_tan = [&](int jreg) {
        real s = 0.0;
        for (int js = 0; js < ns; ++js) s += int_flux_diff[jreg * ns + js] * lg.tan_ang[js];
        return s;
      };
      const real air_abs = SP(air_ext, g, jlay) * (1.0 - SP(air_ssa, g, jlay));
      FL(ndiff, clear_air_abs, g, ilay) = FL(ndiff, clear_air_abs, g, ilay) + air_abs * sum_over_mu(0);
      if (do_vegetation) {
        for (int jreg = 1; jreg < nreg; ++jreg) {
          const real vabs = veg_ext[jlay] * (1.0 - SP(veg_ssa, g, jlay));
          FL(ndiff, veg_air_abs, g, ilay) = FL(ndiff, veg_air_abs, g, ilay) + air_abs * sum_over_mu(jreg);
          FL(ndiff, veg_abs, g, ilay) =
              FL(ndiff, veg_abs, g, ilay) + vabs * sum_over_mu(jreg) * od_scaling(jreg, jlay);
        }
      }
      if (urban) {
        for (int jreg = 0; jreg < nreg; ++jreg)
          FL(ndiff, wall_in, g, ilay) = FL(ndiff, wall_in, g, ilay) + f_wall(jreg, jlay) * (sum_tan(jreg));
        FL(ndiff, wall_net, g, ilay) = FL(ndiff, wall_in, g, ilay) * (1.0 - SP(wall_albedo, g, jlay));
      }
    }
  }
  for (int g = 0; g < nsw; ++g) {
    FC(ndiff, ground_dn_dir, g) = 0.0;
    FC(ndiff, ground_dn, g) = vsum(dn_diff_above[g], 0, n);
    FC(ndiff, ground_net, g) = FC(ndiff, ground_dn, g) - vsum(up_above[g], 0, n);
    for (int jreg = 0; jreg < nreg; ++jreg)
      for (int js = 0; js < ns; ++js) {
        const int ifr = js + jreg * ns;
        FC(ndiff, ground_vertical_diff, g) =
            FC(ndiff, ground_vertical_diff, g) +
            (dn_diff_above[g][ifr] + up_above[g][ifr]) * lg.tan_ang[js] / kPi;
      }
  }
#undef SP
}

// ---------------------------------------------------------------------------
// spartacus_urban_lw (radsurf_urban_lw.F90:35-883) and
// spartacus_forest_lw (radsurf_forest_lw.F90:35-715).
// ---------------------------------------------------------------------------
static void spartacus_lw(bool urban, const ssb200_config &cfg, int nlw, int ns, int nreg,
                         int nlay, int icol, int ilay1, const LegendreGauss &lg,
                         const ssb200_canopy_properties &cp,
                         const ssb200_lw_spectral_properties &lw, double *top_emissivity,
                         double *top_emission, ssb200_canopy_flux *lint,
                         ssb200_canopy_flux *lnorm) {
  const int n = nreg * ns;
  const int nrb = urban ? nreg + 1 : nreg;
  const int m = nrb * ns;
  const int ilay2 = ilay1 + nlay - 1;
  const int nsw = nlw; // for the FC/FL macros
  const double *dz = cp.dz + ilay1;
  const double *building_fraction = urban ? cp.building_fraction + ilay1 : nullptr;
  const double *veg_fraction = cp.veg_fraction ? cp.veg_fraction + ilay1 : nullptr;
  const double *veg_ext = cp.veg_ext ? cp.veg_ext + ilay1 : nullptr;
  const double *veg_fsd = cp.veg_fsd ? cp.veg_fsd + ilay1 : nullptr;
#define LP(arr, g, jlay) (lw.arr[(g) + (size_t)nlw * (ilay1 + (jlay))])
  const bool do_vegetation = (nreg > 1);

  Mat frac(nreg, nlay + 1);
  if (urban) {
    for (int j = 0; j < nlay; ++j) frac(0, j) = 1.0 - building_fraction[j];
    frac(0, nlay) = 1.0;
    if (do_vegetation) {
      for (int j = 0; j < nlay; ++j) {
        frac(0, j) = rmax(0.0, frac(0, j) - veg_fraction[j]);
        const real fv = rmax(0.0, 1.0 - building_fraction[j] - frac(0, j)) / (real)(nreg - 1);
        for (int r = 1; r < nreg; ++r) frac(r, j) = fv;
      }
      for (int r = 1; r < nreg; ++r) frac(r, nlay) = 0.0;
    }
  } else { // forest_lw:  frac(2:) = (1 - frac(1)) / (nreg-1)
    frac(0, nlay) = 1.0;
    for (int j = 0; j < nlay; ++j) {
      frac(0, j) = 1.0 - veg_fraction[j];
      for (int r = 1; r < nreg; ++r) frac(r, j) = (1.0 - frac(0, j)) / (real)(nreg - 1);
    }
    for (int r = 1; r < nreg; ++r) frac(r, nlay) = 0.0;
  }

  std::vector<Mat> u_overlap, v_overlap;
  Mat norm_perim, norm_perim_wall(nreg, nlay);
  if (urban) {
    calc_overlap_matrices_urban(nlay, nreg, frac, u_overlap, v_overlap,
                                cfg.min_vegetation_fraction);
    calc_norm_perim_urban(cfg, nlay, nreg, cp.building_fraction + ilay1,
                          cp.building_scale + ilay1, cp.veg_fraction ? cp.veg_fraction + ilay1 : nullptr,
                          cp.veg_scale ? cp.veg_scale + ilay1 : nullptr,
                          cp.veg_contact_fraction ? cp.veg_contact_fraction + ilay1 : nullptr,
                          norm_perim, norm_perim_wall);
  } else {
    calc_overlap_matrices(nlay, nreg, frac, u_overlap, v_overlap, cfg.min_vegetation_fraction);
    calc_norm_perim_forest(cfg, nlay, nreg, cp.veg_fraction + ilay1, cp.veg_scale + ilay1,
                           norm_perim);
  }

  auto L = [&](int g, int j) { return (size_t)g + (size_t)nlw * j; };
  std::vector<Mat> trans(nlw * nlay, Mat(n, n)), ref(nlw * nlay, Mat(n, n));
  std::vector<Mat> int_flux_mat(nlw * nlay, Mat(n, n));
  std::vector<Vec> source_lay(nlw * nlay, Vec(n, 0.0)), int_source(nlw * nlay, Vec(n, 0.0));
  std::vector<Vec> emiss_reg(nlw * nlay, Vec(nreg, 0.0)), emiss_air(nlw * nlay, Vec(nreg, 0.0)),
      emiss_veg(nlw * nlay, Vec(nreg, 0.0));
  Vec emiss_wall(nlw * nlay, 0.0);
  Mat f_wall(nreg, nlay), od_scaling(nreg, nlay);

  for (int jlay = 0; jlay < nlay; ++jlay) {
    Mat f_exchange(nreg, nreg);
    int jreg;
    for (jreg = 0; jreg < nreg - 1; ++jreg) {
      if (frac(jreg, jlay) <= cfg.min_vegetation_fraction ||
          frac(jreg + 1, jlay) <= cfg.min_vegetation_fraction) {
        f_exchange(jreg + 1, jreg) = 0.0;
        f_exchange(jreg, jreg + 1) = 0.0;
      } else {
        f_exchange(jreg + 1, jreg) = norm_perim(jreg, jlay) / (kPi * frac(jreg, jlay));
        f_exchange(jreg, jreg + 1) = norm_perim(jreg, jlay) / (kPi * frac(jreg + 1, jlay));
      }
    }
    const int jreg_exit = std::max(nreg - 1, 0);
    if (nreg > 2 && norm_perim(nreg - 1, jlay) > 0.0) {
      if (frac(2, jlay) <= cfg.min_vegetation_fraction ||
          frac(0, jlay) <= cfg.min_vegetation_fraction) {
        f_exchange(0, 2) = 0.0;
        f_exchange(2, 0) = 0.0;
      } else {
        f_exchange(0, 2) = norm_perim(jreg_exit, jlay) / (kPi * frac(2, jlay));
        f_exchange(2, 0) = norm_perim(jreg_exit, jlay) / (kPi * frac(0, jlay));
      }
    }
    if (urban) { // urban_lw:374-382
      for (jreg = 0; jreg < nreg; ++jreg) {
        if (frac(jreg, jlay) <= cfg.min_vegetation_fraction)
          f_wall(jreg, jlay) = 0.0;
        else
          f_wall(jreg, jlay) =
              norm_perim_wall(jreg, jlay) * lg.vadjustment2 / (kPi * frac(jreg, jlay));
      }
    }
    if (nreg == 2) {
      od_scaling(1, jlay) = 1.0;
    } else if (nreg == 3) {
      od_scaling(1, jlay) =
          std::exp(-veg_fsd[jlay] * (1.0 + 0.5 * veg_fsd[jlay] * (1.0 + 0.5 * veg_fsd[jlay])));
      od_scaling(2, jlay) = 2.0 - od_scaling(1, jlay);
    }
    real emiss_factor = 0.0; // urban_lw:447
    for (int js = 0; js < ns; ++js) emiss_factor += lg.hweight[js] / lg.mu[js];
    emiss_factor = 2.0 * emiss_factor;

    for (int g = 0; g < nlw; ++g) {
      Vec ext_reg(nreg), ssa_reg(nreg), planck_reg(nreg);
      ext_reg[0] = LP(air_ext, g, jlay);
      ssa_reg[0] = LP(air_ssa, g, jlay);
      planck_reg[0] = LP(clear_air_planck, g, jlay);
      if (nreg == 2) {
        ext_reg[1] = LP(air_ext, g, jlay) + veg_ext[jlay];
        ssa_reg[1] = (ext_reg[0] * ssa_reg[0] + veg_ext[jlay] * LP(veg_ssa, g, jlay)) /
                     rmax(ext_reg[1], 1.0e-8);
        planck_reg[1] = (ext_reg[0] * (1.0 - ssa_reg[0]) * LP(veg_air_planck, g, jlay) +
                         veg_ext[jlay] * (1.0 - LP(veg_ssa, g, jlay)) * LP(veg_planck, g, jlay)) /
                        rmax(ext_reg[1] * (1.0 - ssa_reg[1]), 1.0e-8);
      } else if (nreg == 3) {
        for (int r = 1; r < 3; ++r) {
          ext_reg[r] = LP(air_ext, g, jlay) + od_scaling(r, jlay) * veg_ext[jlay];
          ssa_reg[r] = (ext_reg[0] * ssa_reg[0] +
                        od_scaling(r, jlay) * veg_ext[jlay] * LP(veg_ssa, g, jlay)) /
                       rmax(ext_reg[r], 1.0e-8);
        }
        for (int r = 1; r < 3; ++r)
          planck_reg[r] = (ext_reg[0] * (1.0 - ssa_reg[0]) * LP(veg_air_planck, g, jlay) +
                           od_scaling(r, jlay) * veg_ext[jlay] * (1.0 - LP(veg_ssa, g, jlay)) *
                               LP(veg_planck, g, jlay)) /
                          rmax(ext_reg[r] * (1.0 - ssa_reg[r]), 1.0e-8);
      }
      const real wall_ext = 1.0;
      // sic: spectral index 1 for every interval (urban_lw:392, App. B2)
      const real wall_factor = urban ? 1.0 - LP(wall_emissivity, 0, jlay) : 0.0;

      Mat gamma1(n, n), gamma2(n, n);
      for (int jreg_fr = 0; jreg_fr < nreg; ++jreg_fr)
        for (int jreg_to = 0; jreg_to < nreg; ++jreg_to)
          if (jreg_fr != jreg_to)
            for (int js = 0; js < ns; ++js) {
              const int ifr = js + jreg_fr * ns, ito = js + jreg_to * ns;
              gamma1(ifr, ifr) = gamma1(ifr, ifr) - lg.tan_ang[js] * f_exchange(jreg_to, jreg_fr);
              gamma1(ito, ifr) = +lg.tan_ang[js] * f_exchange(jreg_to, jreg_fr);
            }
      for (jreg = 0; jreg < nreg; ++jreg)
        for (int js = 0; js < ns; ++js) {
          const int ifr = js + jreg * ns;
          if (urban)
            gamma1(ifr, ifr) = gamma1(ifr, ifr) - ext_reg[jreg] / lg.mu[js] -
                               lg.tan_ang[js] * f_wall(jreg, jlay) * wall_ext;
          else
            gamma1(ifr, ifr) = gamma1(ifr, ifr) - ext_reg[jreg] / lg.mu[js];
        }
      for (int js_fr = 0; js_fr < ns; ++js_fr)
        for (int js_to = 0; js_to < ns; ++js_to)
          for (jreg = 0; jreg < nreg; ++jreg) {
            const int ifr = js_fr + jreg * ns, ito = js_to + jreg * ns;
            if (urban)
              gamma2(ito, ifr) =
                  0.5 * (lg.weight[js_to] * ext_reg[jreg] * ssa_reg[jreg] / lg.mu[js_fr] +
                         lg.vweight[js_to] * lg.tan_ang[js_fr] * f_wall(jreg, jlay) * wall_factor);
            else // forest_lw:343-344 groups the constants first
              gamma2(ito, ifr) =
                  (0.5 * lg.weight[js_to] / lg.mu[js_fr]) * ext_reg[jreg] * ssa_reg[jreg];
          }
      gamma1 = gamma1 + gamma2;

      const size_t k = L(g, jlay);
      Vec emiss_rate(n, 0.0);
      for (jreg = 0; jreg < nreg; ++jreg) {
        const real volume_emiss =
            frac(jreg, jlay) * (ext_reg[jreg] * (1.0 - ssa_reg[jreg]) * planck_reg[jreg]);
        real wall_emiss = 0.0;
        if (urban) wall_emiss = norm_perim_wall(jreg, jlay) * lg.vadjustment * LP(wall_emission, g, jlay);
        for (int js = 0; js < ns; ++js) {
          const int ifr = js + jreg * ns;
          if (urban)
            emiss_rate[ifr] = (lg.hweight[js] / lg.mu[js]) * volume_emiss +
                              (0.5 * lg.vweight[js]) * wall_emiss;
          else
            emiss_rate[ifr] = (lg.hweight[js] / lg.mu[js]) * volume_emiss;
        }
        emiss_reg[k][jreg] = emiss_factor * volume_emiss;
        if (jreg > 0) {
          emiss_air[k][jreg] = emiss_factor * frac(jreg, jlay) * ext_reg[0] * (1.0 - ssa_reg[0]) *
                               LP(veg_air_planck, g, jlay);
          emiss_veg[k][jreg] = emiss_factor * frac(jreg, jlay) * veg_ext[jlay] *
                               (1.0 - LP(veg_ssa, g, jlay)) * LP(veg_planck, g, jlay) *
                               od_scaling(jreg, jlay);
        }
      }
      if (urban) {
        real s = 0.0;
        for (jreg = 0; jreg < nreg; ++jreg) s += norm_perim_wall(jreg, jlay);
        emiss_wall[k] = (s * lg.vadjustment) * LP(wall_emission, g, jlay);
      }

      int r0, nr;
      bool veg_branching = urban ? do_vegetation : true;
      if (veg_branching) {
        if (veg_fraction[jlay] <= cfg.min_vegetation_fraction) {
          r0 = 0;
          nr = 1;
        } else if (frac(0, jlay) <= cfg.min_vegetation_fraction) {
          r0 = 1;
          nr = nreg - 1;
        } else {
          r0 = 0;
          nr = nreg;
        }
      } else {
        r0 = 0;
        nr = 1;
      }
      const int i0 = r0 * ns, nn = nr * ns;
      Mat R, T, IF;
      Vec src, isrc;
      Vec er(emiss_rate.begin() + i0, emiss_rate.begin() + i0 + nn);
      calc_matrices_lw_eig(nn, dz[jlay], sub(gamma1, i0, nn, i0, nn), sub(gamma2, i0, nn, i0, nn),
                           er, R, T, src, IF, isrc);
      paste(ref[k], i0, i0, R);
      paste(trans[k], i0, i0, T);
      paste(int_flux_mat[k], i0, i0, IF);
      for (int i = 0; i < nn; ++i) {
        source_lay[k][i0 + i] = src[i];
        int_source[k][i0 + i] = isrc[i];
      }
    }
  }

  // Section 4 (urban_lw:552-614)
  std::vector<Mat> a_above(nlw * (nlay + 1), Mat(n, n)), a_below(nlw * (nlay + 1), Mat(m, m));
  std::vector<Vec> source_above(nlw * (nlay + 1), Vec(n, 0.0)), source_below(nlw * (nlay + 1), Vec(m, 0.0));
  std::vector<Mat> denominator(nlw * nlay, Mat(n, n));
  for (int g = 0; g < nlw; ++g) {
    const real ground_emissivity = lw.ground_emissivity[g + (size_t)nlw * icol];
    const real ground_emission = lw.ground_emission[g + (size_t)nlw * icol];
    for (int jreg = 0; jreg < nreg; ++jreg) {
      for (int js_to = 0; js_to < ns; ++js_to)
        for (int js_fr = 0; js_fr < ns; ++js_fr)
          a_above[L(g, 0)](js_to + jreg * ns, js_fr + jreg * ns) =
              (1.0 - ground_emissivity) * lg.hweight[js_to];
      for (int js = 0; js < ns; ++js)
        source_above[L(g, 0)][js + jreg * ns] = (lg.hweight[js] * frac(jreg, 0)) * ground_emission;
    }
  }
  for (int jlay = 0; jlay < nlay; ++jlay)
    for (int g = 0; g < nlw; ++g) {
      const size_t k = L(g, jlay), k1 = L(g, jlay + 1);
      denominator[k] = identity_minus_mat_x_mat(a_above[k], ref[k]);
      Mat ab = ref[k] + matmul(trans[k], solve_mat(denominator[k], matmul(a_above[k], trans[k])));
      Vec sb = source_lay[k] +
               matvec(trans[k], solve_vec(denominator[k], source_above[k] + matvec(a_above[k], source_lay[k])));
      paste(a_below[k1], 0, 0, ab);
      for (int i = 0; i < n; ++i) source_below[k1][i] = sb[i];
      if (urban) {
        real exposed_roof_frac;
        if (jlay < nlay - 1)
          exposed_roof_frac = rmax(0.0, building_fraction[jlay] - building_fraction[jlay + 1]);
        else
          exposed_roof_frac = building_fraction[jlay];
        for (int js = 0; js < ns; ++js) {
          for (int j2 = 0; j2 < ns; ++j2)
            a_below[k1](n + js, n + j2) = (1.0 - LP(roof_emissivity, g, jlay)) * lg.hweight[js];
          source_below[k1][n + js] = lg.hweight[js] * LP(roof_emission, g, jlay) * exposed_roof_frac;
        }
      }
      a_above[k1] = expandedmat_x_mat(nreg, nrb, ns, u_overlap[jlay + 1],
                                      mat_x_expandedmat(nrb, nreg, ns, a_below[k1], v_overlap[jlay + 1]));
      source_above[k1] = expandedmat_x_vec(nreg, nrb, ns, u_overlap[jlay + 1], source_below[k1]);
    }

  Vec temis(nlw), tsrc(nlw);
  for (int g = 0; g < nlw; ++g) {
    Vec y = matvec(sub(a_above[L(g, nlay)], 0, ns, 0, ns), lg.hweight);
    temis[g] = 1.0 - vsum(y, 0, ns);
    tsrc[g] = vsum(source_above[L(g, nlay)], 0, ns);
    top_emissivity[g] = temis[g];
    top_emission[g] = tsrc[g];
  }

  flux_zero(lint, icol, ilay1, ilay2);
  flux_zero(lnorm, icol, ilay1, ilay2);

  // Internal emission pass (urban_lw:650-748)
  std::vector<Vec> dn_above(nlw, Vec(n, 0.0)), up_above(nlw, Vec(n, 0.0));
  for (int g = 0; g < nlw; ++g) {
    FC(lint, top_dn, g) = 0.0;
    FC(lint, top_net, g) = -tsrc[g];
  }
  for (int jlay = nlay - 1; jlay >= 0; --jlay) {
    const int ilay = ilay1 + jlay;
    for (int g = 0; g < nlw; ++g) {
      const size_t k = L(g, jlay), k1 = L(g, jlay + 1);
      Vec dn_below = expandedmat_x_vec(nrb, nreg, ns, v_overlap[jlay + 1], dn_above[g]);
      Vec up_below = matvec(a_below[k1], dn_below) + source_below[k1];
      if (urban) {
        FL(lint, roof_in, g, ilay) = vsum(dn_below, n, ns);
        FL(lint, roof_net, g, ilay) = FL(lint, roof_in, g, ilay) - vsum(up_below, n, ns);
      }
      Vec dn_b(dn_below.begin(), dn_below.begin() + n);
      dn_above[g] = solve_vec(denominator[k], (matvec(trans[k], dn_b) + matvec(ref[k], source_above[k])) + source_lay[k]);
      up_above[g] = matvec(a_above[k], dn_above[g]) + source_above[k];
      if (lint->flux_dn_layer_top) {
        FL(lint, flux_dn_layer_top, g, ilay) = vsum(dn_below, 0, n);
        FL(lint, flux_up_layer_top, g, ilay) = vsum(up_below, 0, n);
        FL(lint, flux_dn_layer_base, g, ilay) = vsum(dn_above[g], 0, n);
        FL(lint, flux_up_layer_base, g, ilay) = vsum(up_above[g], 0, n);
      }
      Vec int_flux = matvec(int_flux_mat[k], dn_b + up_above[g]) + int_source[k];
      auto sum_over_mu = [&](int jreg) {
        real s = 0.0;
        for (int js = 0; js < ns; ++js) s += int_flux[jreg * ns + js] * (1.0 / lg.mu[js]);
        return s;
      };
      auto sum_tan = [&](int jreg) {
        real s = 0.0;
        for (int js = 0; js < ns; ++js) s += int_flux[jreg * ns + js] * lg.tan_ang[js];
        return s;
      };
      const real air_abs = LP(air_ext, g, jlay) * (1.0 - LP(air_ssa, g, jlay));
      FL(lint, clear_air_abs, g, ilay) =
          FL(lint, clear_air_abs, g, ilay) + air_abs * sum_over_mu(0) - emiss_reg[k][0] * dz[jlay];
      if (do_vegetation) {
        for (int jreg = 1; jreg < nreg; ++jreg) {
          const real vabs = veg_ext[jlay] * (1.0 - LP(veg_ssa, g, jlay));
          FL(lint, veg_air_abs, g, ilay) =
              FL(lint, veg_air_abs, g, ilay) + air_abs * sum_over_mu(jreg) - emiss_air[k][jreg] * dz[jlay];
          FL(lint, veg_abs, g, ilay) = FL(lint, veg_abs, g, ilay) +
                                       vabs * sum_over_mu(jreg) * od_scaling(jreg, jlay) -
                                       emiss_veg[k][jreg] * dz[jlay];
        }
      }
      if (urban) {
        for (int jreg = 0; jreg < nreg; ++jreg)
          FL(lint, wall_in, g, ilay) = FL(lint, wall_in, g, ilay) + f_wall(jreg, jlay) * sum_tan(jreg);
        FL(lint, wall_net, g, ilay) =
            FL(lint, wall_in, g, ilay) * LP(wall_emissivity, g, jlay) - emiss_wall[k] * dz[jlay];
      }
    }
  }
  for (int g = 0; g < nlw; ++g) {
    FC(lint, ground_dn, g) = vsum(dn_above[g], 0, n);
    FC(lint, ground_net, g) = FC(lint, ground_dn, g) - vsum(up_above[g], 0, n);
    for (int jreg = 0; jreg < nreg; ++jreg)
      for (int js = 0; js < ns; ++js) {
        const int ifr = js + jreg * ns;
        FC(lint, ground_vertical_diff, g) =
            FC(lint, ground_vertical_diff, g) + (dn_above[g][ifr] + up_above[g][ifr]) * lg.tan_ang[js] / kPi;
      }
  }

  // Normalised incoming pass (urban_lw:754-858)
  for (int g = 0; g < nlw; ++g) {
    dn_above[g].assign(n, 0.0);
    for (int js = 0; js < ns; ++js) dn_above[g][js] = lg.hweight[js];
    FC(lnorm, top_dn, g) = 1.0;
    FC(lnorm, top_net, g) = temis[g];
  }
  for (int jlay = nlay - 1; jlay >= 0; --jlay) {
    const int ilay = ilay1 + jlay;
    for (int g = 0; g < nlw; ++g) {
      const size_t k = L(g, jlay), k1 = L(g, jlay + 1);
      Vec dn_below = expandedmat_x_vec(nrb, nreg, ns, v_overlap[jlay + 1], dn_above[g]);
      Vec up_below = matvec(a_below[k1], dn_below);
      if (urban) {
        FL(lnorm, roof_in, g, ilay) = vsum(dn_below, n, ns);
        FL(lnorm, roof_net, g, ilay) = FL(lnorm, roof_in, g, ilay) - vsum(up_below, n, ns);
      }
      Vec dn_b(dn_below.begin(), dn_below.begin() + n);
      dn_above[g] = solve_vec(denominator[k], matvec(trans[k], dn_b));
      up_above[g] = matvec(a_above[k], dn_above[g]);
      if (lnorm->flux_dn_layer_top) {
        FL(lnorm, flux_dn_layer_top, g, ilay) = vsum(dn_below, 0, n);
        FL(lnorm, flux_up_layer_top, g, ilay) = vsum(up_below, 0, n);
        FL(lnorm, flux_dn_layer_base, g, ilay) = vsum(dn_above[g], 0, n);
        FL(lnorm, flux_up_layer_base, g, ilay) = vsum(up_above[g], 0, n);
      }
      Vec int_flux = matvec(int_flux_mat[k], dn_b + up_above[g]);
      auto sum_over_mu = [&](int jreg) {
        real s = 0.0;
        for (int js = 0; js < ns; ++js) s += int_flux[jreg * ns + js] * (1.0 / lg.mu[js]);
        return s;
      };
      auto sum_tan = [&](int jreg) {
        real s = 0.0;
        for (int js = 0; js < ns; ++js) s += int_flux[jreg * ns + js] * lg.tan_ang[js];
        return s;
      };
      const real air_abs = LP(air_ext, g, jlay) * (1.0 - LP(air_ssa, g, jlay));
      FL(lnorm, clear_air_abs, g, ilay) = FL(lnorm, clear_air_abs, g, ilay) + air_abs * sum_over_mu(0);
      if (do_vegetation) {
        for (int jreg = 1; jreg < nreg; ++jreg) {
          const real vabs = veg_ext[jlay] * (1.0 - LP(veg_ssa, g, jlay));
          FL(lnorm, veg_air_abs, g, ilay) = FL(lnorm, veg_air_abs, g, ilay) + air_abs * sum_over_mu(jreg);
          FL(lnorm, veg_abs, g, ilay) =
              FL(lnorm, veg_abs, g, ilay) + vabs * sum_over_mu(jreg) * od_scaling(jreg, jlay);
        }
      }
      if (urban) {
        for (int jreg = 0; jreg < nreg; ++jreg)
          FL(lnorm, wall_in, g, ilay) = FL(lnorm, wall_in, g, ilay) + f_wall(jreg, jlay) * (sum_tan(jreg));
        FL(lnorm, wall_net, g, ilay) = FL(lnorm, wall_in, g, ilay) * LP(wall_emissivity, g, jlay);
      }
    }
  }
  for (int g = 0; g < nlw; ++g) {
    FC(lnorm, ground_dn, g) = vsum(dn_above[g], 0, n);
    FC(lnorm, ground_net, g) = FC(lnorm, ground_dn, g) - vsum(up_above[g], 0, n);
    // forest_lw:687-694 accumulates this pass into lw_internal (App. B5)
    ssb200_canopy_flux *tgt = urban ? lnorm : lint;
    for (int jreg = 0; jreg < nreg; ++jreg)
      for (int js = 0; js < ns; ++js) {
        const int ifr = js + jreg * ns;
        FC(tgt, ground_vertical_diff, g) =
            FC(tgt, ground_vertical_diff, g) + (dn_above[g][ifr] + up_above[g][ifr]) * lg.tan_ang[js] / kPi;
      }
  }
#undef LP
}

// ---------------------------------------------------------------------------
// simple_urban_sw: radsurf_simple_urban_sw.F90:28-294.  icol, ilay 0-based.
// The reference indexes several (nspec,ncol) fields with ilay (:193,219,253-256; App. B6), which
// is only defined when istartlay(icol) = icol (as in test/single_layer).  They are indexed with
// icol here: identical whenever the reference is well defined, in bounds otherwise.
// ---------------------------------------------------------------------------
static int simple_urban_sw(const ssb200_config &cfg, bool is_infinite_street, int nsw, int icol,
                           int ilay, real cos_sza, const ssb200_canopy_properties &cp,
                           const ssb200_sw_spectral_properties &sw,
                           const double *ground_albedo_diff, const double *ground_albedo_dir,
                           ssb200_canopy_flux *ndir, ssb200_canopy_flux *ndiff) {
  const real dz = cp.dz[ilay];
  const double building_fraction = cp.building_fraction[ilay];
  const double building_scale = cp.building_scale[ilay];
  double veg_fraction = 0.0, veg_scale = 1.0, veg_contact_fraction = 0.0;
  Mat norm_perim, norm_perim_wall;
  calc_norm_perim_urban(cfg, 1, 1, &building_fraction, &building_scale, &veg_fraction, &veg_scale,
                        &veg_contact_fraction, norm_perim, norm_perim_wall);
  const real npw = norm_perim_wall(0, 0);
  real view_ground_sky, view_wall_wall, view_dir_ground;
  if (is_infinite_street) {
    const real street_width = 2.0 * (1.0 - building_fraction) / npw;
    calc_view_factors_inf(dz / street_width, view_ground_sky, view_wall_wall, &cos_sza, &view_dir_ground);
  } else {
    const real building_separation_scale = kPi * (1.0 - building_fraction) / npw;
    calc_view_factors_exp(dz / building_separation_scale, view_ground_sky, view_wall_wall, &cos_sza,
                          &view_dir_ground);
  }
  const real view_dir_wall = 1.0 - view_dir_ground;
  const real view_wall_ground = 0.5 * (1.0 - view_wall_wall);
  const real view_ground_wall = 1.0 - view_ground_sky;
  flux_zero(ndiff, icol, ilay, ilay);
  flux_zero(ndir, icol, ilay, ilay);
#define XC(f, member, g, c) (f->member[(g) + (size_t)nsw * (c)])
  for (int g = 0; g < nsw; ++g) {
    const real roof_albedo = sw.roof_albedo[g + (size_t)nsw * ilay];
    const real wall_albedo = sw.wall_albedo[g + (size_t)nsw * ilay];
    Mat im(2, 2);
    im(0, 0) = 1.0;
    im(0, 1) = -view_wall_ground * wall_albedo;
    im(1, 0) = -view_ground_wall * ground_albedo_diff[g];
    im(1, 1) = 1.0 - view_wall_wall * wall_albedo;
    Vec src(2), sol;
    src[0] = 0.0;
    src[1] = (view_dir_wall + ground_albedo_dir[g] * view_dir_ground * view_ground_wall) *
             (1.0 - building_fraction);
    sol = solve_vec(im, src);
    XC(ndir, ground_dn_dir, g, icol) = view_dir_ground * (1.0 - building_fraction);
    XC(ndir, ground_dn, g, icol) = XC(ndir, ground_dn_dir, g, icol) + sol[0];
    XC(ndir, ground_net, g, icol) = XC(ndir, ground_dn_dir, g, icol) * (1.0 - ground_albedo_dir[g]) +
                                    sol[0] * (1.0 - ground_albedo_diff[g]);
    ndir->ground_sunlit_frac[icol] = view_dir_ground;
    XC(ndir, roof_in_dir, g, ilay) = building_fraction;
    XC(ndir, roof_in, g, ilay) = building_fraction;
    XC(ndir, roof_net, g, ilay) = building_fraction * (1.0 - roof_albedo);
    ndir->roof_sunlit_frac[ilay] = 1.0;
    XC(ndir, wall_in_dir, g, ilay) = view_dir_wall * (1.0 - building_fraction);
    XC(ndir, wall_in, g, ilay) = sol[1];
    XC(ndir, wall_net, g, ilay) = XC(ndir, wall_in, g, ilay) * (1.0 - wall_albedo);
    const real tan_sza = std::sqrt(1.0 / (cos_sza * cos_sza) - 1.0);
    ndir->wall_sunlit_frac[ilay] =
        0.5 * view_dir_wall /
        (rmax(tan_sza, 1.0e-6) * npw * dz / (kPi * (1.0 - building_fraction)));
    XC(ndir, top_dn_dir, g, icol) = 1.0;
    XC(ndir, top_dn, g, icol) = 1.0;
    XC(ndir, top_net, g, icol) =
        1.0 - building_fraction * roof_albedo -
        (XC(ndir, ground_dn, g, icol) - XC(ndir, ground_net, g, icol)) * view_ground_sky -
        (XC(ndir, wall_in, g, ilay) - XC(ndir, wall_net, g, ilay)) * view_wall_ground;
    if (ndir->flux_dn_layer_top) {
      XC(ndir, flux_dn_dir_layer_top, g, ilay) = (1.0 - building_fraction);
      XC(ndir, flux_dn_layer_top, g, ilay) = (1.0 - building_fraction);
      XC(ndir, flux_up_layer_top, g, ilay) =
          (XC(ndir, ground_dn, g, icol) - XC(ndir, ground_net, g, icol)) * view_ground_sky +
          (XC(ndir, wall_in, g, ilay) - XC(ndir, wall_net, g, ilay)) * view_wall_ground;
      XC(ndir, flux_dn_dir_layer_base, g, ilay) = XC(ndir, ground_dn_dir, g, icol);
      XC(ndir, flux_dn_layer_base, g, ilay) = XC(ndir, ground_dn, g, icol);
      XC(ndir, flux_up_layer_base, g, ilay) = XC(ndir, ground_dn, g, icol) - XC(ndir, ground_net, g, icol);
    }
    src[0] = view_ground_sky * (1.0 - building_fraction);
    src[1] = view_ground_wall * (1.0 - building_fraction);
    sol = solve_vec(im, src);
    XC(ndiff, ground_dn_dir, g, icol) = 0.0;
    XC(ndiff, ground_dn, g, icol) = sol[0];
    XC(ndiff, ground_net, g, icol) = XC(ndiff, ground_dn, g, icol) * (1.0 - ground_albedo_diff[g]);
    XC(ndiff, roof_in, g, ilay) = building_fraction;
    XC(ndiff, roof_net, g, ilay) = building_fraction * (1.0 - roof_albedo);
    XC(ndiff, wall_in, g, ilay) = sol[1];
    XC(ndiff, wall_net, g, ilay) = XC(ndiff, wall_in, g, ilay) * (1.0 - wall_albedo);
    XC(ndiff, top_dn_dir, g, icol) = 0.0;
    XC(ndiff, top_dn, g, icol) = 1.0;
    XC(ndiff, top_net, g, icol) =
        1.0 - building_fraction * roof_albedo -
        (XC(ndiff, ground_dn, g, icol) - XC(ndiff, ground_net, g, icol)) * view_ground_sky -
        (XC(ndiff, wall_in, g, ilay) - XC(ndiff, wall_net, g, ilay)) * view_wall_ground;
    if (ndiff->flux_dn_layer_top) {
      XC(ndiff, flux_dn_layer_top, g, ilay) = (1.0 - building_fraction);
      XC(ndiff, flux_up_layer_top, g, ilay) =
          (XC(ndiff, ground_dn, g, icol) - XC(ndiff, ground_net, g, icol)) * view_ground_sky +
          (XC(ndiff, wall_in, g, ilay) - XC(ndiff, wall_net, g, ilay)) * view_wall_ground;
      XC(ndiff, flux_dn_layer_base, g, ilay) = XC(ndiff, ground_dn, g, icol);
      XC(ndiff, flux_up_layer_base, g, ilay) = XC(ndiff, ground_dn, g, icol) - XC(ndiff, ground_net, g, icol);
    }
  }
  return 0;
}

// simple_urban_lw: radsurf_simple_urban_lw.F90:28-257.
static int simple_urban_lw(const ssb200_config &cfg, bool is_infinite_street, int nlw, int icol,
                           int ilay, const ssb200_canopy_properties &cp,
                           const ssb200_lw_spectral_properties &lw, ssb200_canopy_flux *lint,
                           ssb200_canopy_flux *lnorm) {
  const int nsw = nlw;
  const real dz = cp.dz[ilay];
  const double building_fraction = cp.building_fraction[ilay];
  const double building_scale = cp.building_scale[ilay];
  double veg_fraction = 0.0, veg_scale = 1.0, veg_contact_fraction = 0.0;
  Mat norm_perim, norm_perim_wall;
  calc_norm_perim_urban(cfg, 1, 1, &building_fraction, &building_scale, &veg_fraction, &veg_scale,
                        &veg_contact_fraction, norm_perim, norm_perim_wall);
  const real npw = norm_perim_wall(0, 0);
  real view_ground_sky, view_wall_wall;
  if (is_infinite_street) {
    const real street_width = 2.0 * (1.0 - building_fraction) / npw;
    calc_view_factors_inf(dz / street_width, view_ground_sky, view_wall_wall, nullptr, nullptr);
  } else {
    const real building_separation_scale = kPi * (1.0 - building_fraction) / npw;
    calc_view_factors_exp(dz / building_separation_scale, view_ground_sky, view_wall_wall, nullptr, nullptr);
  }
  const real view_wall_ground = 0.5 * (1.0 - view_wall_wall);
  const real view_ground_wall = 1.0 - view_ground_sky;
  flux_zero(lnorm, icol, ilay, ilay);
  flux_zero(lint, icol, ilay, ilay);
  for (int g = 0; g < nlw; ++g) {
    const real ground_emissivity = lw.ground_emissivity[g + (size_t)nlw * icol];
    const real ground_emission = lw.ground_emission[g + (size_t)nlw * icol];
    const real roof_emissivity = lw.roof_emissivity[g + (size_t)nlw * ilay];
    const real roof_emission = lw.roof_emission[g + (size_t)nlw * ilay];
    const real wall_emissivity = lw.wall_emissivity[g + (size_t)nlw * ilay];
    const real wall_emission = lw.wall_emission[g + (size_t)nlw * ilay];
    Mat im(2, 2);
    im(0, 0) = 1.0;
    im(0, 1) = -view_wall_ground * (1.0 - wall_emissivity);
    im(1, 0) = -view_ground_wall * (1.0 - ground_emissivity);
    im(1, 1) = 1.0 - view_wall_wall * (1.0 - ground_emissivity); // sic (App. B6)
    Vec src(2), sol;
    src[0] = view_wall_ground * wall_emission * npw * dz;
    src[1] = view_ground_wall * ground_emission * (1.0 - building_fraction) +
             view_wall_wall * wall_emission * npw * dz;
    sol = solve_vec(im, src);
    XC(lint, ground_dn, g, icol) = sol[0];
    XC(lint, ground_net, g, icol) = sol[0] * ground_emissivity - ground_emission * (1.0 - building_fraction);
    XC(lint, roof_in, g, ilay) = 0.0;
    XC(lint, roof_net, g, ilay) = -building_fraction * roof_emission;
    XC(lint, wall_in, g, ilay) = sol[1];
    XC(lint, wall_net, g, ilay) = sol[1] * wall_emissivity - wall_emission * npw * dz;
    XC(lint, top_dn, g, icol) = 0.0;
    XC(lint, top_net, g, icol) =
        -building_fraction * roof_emission -
        (XC(lint, ground_dn, g, icol) - XC(lint, ground_net, g, icol)) * view_ground_sky -
        (XC(lint, wall_in, g, ilay) - XC(lint, wall_net, g, ilay)) * view_wall_ground;
    if (lint->flux_dn_layer_top) {
      XC(lint, flux_dn_layer_top, g, ilay) = 0.0;
      XC(lint, flux_up_layer_top, g, ilay) =
          (XC(lint, ground_dn, g, icol) - XC(lint, ground_net, g, icol)) * view_ground_sky +
          (XC(lint, wall_in, g, ilay) - XC(lint, wall_net, g, ilay)) * view_wall_ground;
      XC(lint, flux_dn_layer_base, g, ilay) = XC(lint, ground_dn, g, icol);
      XC(lint, flux_up_layer_base, g, ilay) = XC(lint, ground_dn, g, icol) - XC(lint, ground_net, g, icol);
    }
    src[0] = view_ground_sky * (1.0 - building_fraction);
    src[1] = view_ground_wall * (1.0 - building_fraction);
    sol = solve_vec(im, src);
    XC(lnorm, ground_dn, g, icol) = sol[0];
    XC(lnorm, ground_net, g, icol) = XC(lnorm, ground_dn, g, icol) * ground_emissivity;
    XC(lnorm, roof_in, g, ilay) = building_fraction;
    XC(lnorm, roof_net, g, ilay) = building_fraction * roof_emissivity;
    XC(lnorm, wall_in, g, ilay) = sol[1];
    XC(lnorm, wall_net, g, ilay) = XC(lnorm, wall_in, g, ilay) * wall_emissivity;
    XC(lnorm, top_dn, g, icol) = 1.0;
    XC(lnorm, top_net, g, icol) =
        1.0 - building_fraction * (1.0 - roof_emissivity) -
        (XC(lnorm, ground_dn, g, icol) - XC(lnorm, ground_net, g, icol)) * view_ground_sky -
        (XC(lnorm, wall_in, g, ilay) - XC(lnorm, wall_net, g, ilay)) * view_wall_ground;
    if (lnorm->flux_dn_layer_top) {
      XC(lnorm, flux_dn_layer_top, g, ilay) = 1.0 - building_fraction;
      XC(lnorm, flux_up_layer_top, g, ilay) =
          (XC(lnorm, ground_dn, g, icol) - XC(lnorm, ground_net, g, icol)) * view_ground_sky +
          (XC(lnorm, wall_in, g, ilay) - XC(lnorm, wall_net, g, ilay)) * view_wall_ground;
      XC(lnorm, flux_dn_layer_base, g, ilay) = XC(lnorm, ground_dn, g, icol);
      XC(lnorm, flux_up_layer_base, g, ilay) = XC(lnorm, ground_dn, g, icol) - XC(lnorm, ground_net, g, icol);
    }
  }
  return 0;
}
#undef XC

// ---------------------------------------------------------------------------
// radsurf: radsurf/radsurf_interface.F90:20-317.  One column.
// ---------------------------------------------------------------------------
static int radsurf_column(const ssb200_config &cfg, const ssb200_canopy_properties &cp,
                          const ssb200_sw_spectral_properties *sw,
                          const ssb200_lw_spectral_properties *lw, ssb200_boundary_conds_out *bc,
                          int jcol, ssb200_canopy_flux *ndir, ssb200_canopy_flux *ndiff,
                          ssb200_canopy_flux *lint, ssb200_canopy_flux *lnorm,
                          const LegendreGauss &lg_sw_f, const LegendreGauss &lg_sw_u,
                          const LegendreGauss &lg_lw_f, const LegendreGauss &lg_lw_u) {
  const int irep = cp.i_representation[jcol];
  const int nsw = cfg.nsw, nlw = cfg.nlw;
  int ilay1 = 0, ilay2 = -1, nlay = 0;
  if (irep != SSB200_TILE_FLAT) {
    ilay1 = cp.istartlay[jcol] - 1;
    nlay = cp.nlay[jcol];
    ilay2 = ilay1 + nlay - 1;
  }
  const double *galb = sw ? sw->ground_albedo + (size_t)nsw * jcol : nullptr;
  const double *galb_dir = nullptr;
  if (sw) galb_dir = (cfg.use_sw_direct_albedo ? sw->ground_albedo_dir : sw->ground_albedo) + (size_t)nsw * jcol;
  const int nsw_ = nsw;
#define CC(f, member, g) (f->member[(g) + (size_t)nspec_ * jcol])
  switch (irep) {
  case SSB200_TILE_FLAT: {
    if (cfg.do_sw) {
      const int nspec_ = nsw_;
      for (int g = 0; g < nsw; ++g) {
        bc->sw_albedo[g + (size_t)nsw * jcol] = galb[g];
        bc->sw_albedo_dir[g + (size_t)nsw * jcol] = galb_dir[g];
        CC(ndir, ground_dn_dir, g) = 1.0;
        CC(ndir, ground_dn, g) = 1.0;
        CC(ndir, ground_net, g) = 1.0 - galb_dir[g];
        CC(ndir, ground_vertical_diff, g) = 0.5 * galb_dir[g];
        CC(ndir, top_dn_dir, g) = 1.0;
        CC(ndir, top_dn, g) = 1.0;
        CC(ndir, top_net, g) = 1.0 - galb_dir[g];
        CC(ndiff, ground_dn_dir, g) = 0.0;
        CC(ndiff, ground_dn, g) = 1.0;
        CC(ndiff, ground_net, g) = 1.0 - galb[g];
        CC(ndiff, ground_vertical_diff, g) = 0.5 * (1.0 + galb[g]);
        CC(ndiff, top_dn_dir, g) = 0.0;
        CC(ndiff, top_dn, g) = 1.0;
        CC(ndiff, top_net, g) = 1.0 - galb[g];
      }
    }
    if (cfg.do_lw) {
      const int nspec_ = nlw;
      for (int g = 0; g < nlw; ++g) {
        const real em = lw->ground_emissivity[g + (size_t)nlw * jcol];
        const real es = lw->ground_emission[g + (size_t)nlw * jcol];
        bc->lw_emissivity[g + (size_t)nlw * jcol] = em;
        bc->lw_emission[g + (size_t)nlw * jcol] = es;
        CC(lint, ground_dn, g) = 0.0;
        CC(lint, ground_net, g) = -es;
        CC(lint, ground_vertical_diff, g) = 0.5 * es;
        CC(lint, top_dn, g) = 0.0;
        CC(lint, top_net, g) = -es;
        CC(lnorm, ground_dn, g) = 1.0;
        CC(lnorm, ground_net, g) = em;
        CC(lnorm, ground_vertical_diff, g) = 0.5 * (2.0 - em);
        CC(lnorm, top_dn, g) = 1.0;
        CC(lnorm, top_net, g) = em;
      }
    }
    break;
  }
  case SSB200_TILE_FOREST: {
    if (cfg.do_sw) {
      if (cp.cos_sza[jcol] > 0.0) {
        spartacus_sw(false, cfg, nsw, lg_sw_f.nstream, cfg.n_vegetation_region_forest + 1, nlay, jcol,
                     ilay1, lg_sw_f, cp.cos_sza[jcol], cp, *sw, galb, galb_dir,
                     bc->sw_albedo + (size_t)nsw * jcol, bc->sw_albedo_dir + (size_t)nsw * jcol, ndir, ndiff);
      } else {
        flux_zero(ndir, jcol, ilay1, ilay2);
        flux_zero(ndiff, jcol, ilay1, ilay2);
      }
    }
    if (cfg.do_lw)
      spartacus_lw(false, cfg, nlw, lg_lw_f.nstream, cfg.n_vegetation_region_forest + 1, nlay, jcol, ilay1,
                   lg_lw_f, cp, *lw, bc->lw_emissivity + (size_t)nlw * jcol,
                   bc->lw_emission + (size_t)nlw * jcol, lint, lnorm);
    break;
  }
  case SSB200_TILE_URBAN: {
    if (cfg.do_sw) {
      flux_zero(ndir, jcol, ilay1, ilay2);
      flux_zero(ndiff, jcol, ilay1, ilay2);
      if (cp.cos_sza[jcol] > 0.0)
        spartacus_sw(true, cfg, nsw, lg_sw_u.nstream, 1, nlay, jcol, ilay1, lg_sw_u, cp.cos_sza[jcol], cp,
                     *sw, galb, galb_dir, bc->sw_albedo + (size_t)nsw * jcol,
                     bc->sw_albedo_dir + (size_t)nsw * jcol, ndir, ndiff);
    }
    if (cfg.do_lw) {
      flux_zero(lint, jcol, ilay1, ilay2);
      flux_zero(lnorm, jcol, ilay1, ilay2);
      spartacus_lw(true, cfg, nlw, lg_lw_u.nstream, 1, nlay, jcol, ilay1, lg_lw_u, cp, *lw,
                   bc->lw_emissivity + (size_t)nlw * jcol, bc->lw_emission + (size_t)nlw * jcol, lint, lnorm);
    }
    break;
  }
  case SSB200_TILE_VEGETATED_URBAN: {
    if (cfg.do_sw) {
      if (cp.cos_sza[jcol] > 0.0) {
        spartacus_sw(true, cfg, nsw, lg_sw_u.nstream, cfg.n_vegetation_region_urban + 1, nlay, jcol, ilay1,
                     lg_sw_u, cp.cos_sza[jcol], cp, *sw, galb, galb_dir,
                     bc->sw_albedo + (size_t)nsw * jcol, bc->sw_albedo_dir + (size_t)nsw * jcol, ndir, ndiff);
      } else {
        flux_zero(ndir, jcol, ilay1, ilay2);
        flux_zero(ndiff, jcol, ilay1, ilay2);
      }
    }
    if (cfg.do_lw)
      spartacus_lw(true, cfg, nlw, lg_lw_u.nstream, cfg.n_vegetation_region_urban + 1, nlay, jcol, ilay1,
                   lg_lw_u, cp, *lw, bc->lw_emissivity + (size_t)nlw * jcol,
                   bc->lw_emission + (size_t)nlw * jcol, lint, lnorm);
    break;
  }
  case SSB200_TILE_SIMPLE_URBAN:
  case SSB200_TILE_INFINITE_STREET: {
    const bool is_inf = (irep == SSB200_TILE_INFINITE_STREET);
    if (nlay > 1) return SSB200_ERR_SIMPLE_URBAN_LAYERS;
    if (cfg.do_sw) {
      if (cp.cos_sza[jcol] > 0.0) {
        int rc = simple_urban_sw(cfg, is_inf, nsw, jcol, ilay1, cp.cos_sza[jcol], cp, *sw, galb, galb_dir,
                                 ndir, ndiff);
        if (rc) return rc;
      } else {
        flux_zero(ndir, jcol, ilay1, ilay2);
        flux_zero(ndiff, jcol, ilay1, ilay2);
      }
    }
    if (cfg.do_lw) {
      int rc = simple_urban_lw(cfg, is_inf, nlw, jcol, ilay1, cp, *lw, lint, lnorm);
      if (rc) return rc;
    }
    break;
  }
  default:
    break;
  }
#undef CC
  return 0;
}

// FP64 <-> real at the C boundary (identity copies in the double builds)
static void load_reals(real *dst, const double *src, size_t n) {
  for (size_t i = 0; i < n; ++i) dst[i] = src[i];
}
static void store_reals(double *dst, const real *src, size_t n) {
  for (size_t i = 0; i < n; ++i) dst[i] = (double)src[i];
}

} // namespace orc

// ---------------------------------------------------------------------------
// C entry points of liboracle.so (loaded by tests / bench cpu_baseline only)
// ---------------------------------------------------------------------------
namespace orc {
static bool g_count_flops = false;
static double g_flops_total = 0.0;
}

extern "C" {

// Instrumented flop counter of the radtool layer (see oracle_radtool.hpp): enable, run
// oracle_radsurf, read.  Elementwise work outside radtool (Gamma assembly, flux partition)
// is not counted - it is a few hundred flops per layer.
void oracle_flops_enable(int on) {
  orc::g_count_flops = on != 0;
  orc::g_flops_total = 0.0;
}
double oracle_flops_read(void) { return orc::g_flops_total; }

int oracle_legendre_gauss_init(int32_t nstream, ssb200_legendre_gauss *out) {
  if (!out || nstream < 1 || nstream > SSB200_MAX_NSTREAM) return SSB200_ERR_ARG;
  orc::LegendreGauss lg;
  orc::legendre_gauss_initialize(lg, nstream);
  std::memset(out, 0, sizeof(*out));
  out->nstream = nstream;
  for (int i = 0; i < nstream; ++i) {
    out->mu[i] = lg.mu[i];
    out->sin_ang[i] = lg.sin_ang[i];
    out->tan_ang[i] = lg.tan_ang[i];
    out->weight[i] = lg.weight[i];
    out->hweight[i] = lg.hweight[i];
    out->vweight[i] = lg.vweight[i];
  }
  out->vadjustment = lg.vadjustment;
  out->vadjustment2 = lg.vadjustment2;
  return 0;
}

// nthreads <= 0: all OpenMP threads.  Columns are distributed in blocks of
// `nblocksize` like driver/spartacus_surface_driver.F90:203-234 (default 16).
int oracle_radsurf(const ssb200_config *config, const ssb200_canopy_properties *cp,
                   const ssb200_sw_spectral_properties *sw,
                   const ssb200_lw_spectral_properties *lw, ssb200_boundary_conds_out *bc,
                   int32_t istartcol, int32_t iendcol, ssb200_canopy_flux *sw_norm_dir,
                   ssb200_canopy_flux *sw_norm_diff, ssb200_canopy_flux *lw_internal,
                   ssb200_canopy_flux *lw_norm, int32_t nthreads, int32_t nblocksize) {
  if (!config || !cp || !bc) return SSB200_ERR_ARG;
  int icol1 = istartcol > 0 ? istartcol : 1;
  int icol2 = iendcol > 0 ? iendcol : cp->ncol;
  if (icol2 > cp->ncol) icol2 = cp->ncol;
  const orc::LegendreGauss lg_sw_f = orc::lg_from_c(config->lg_sw_forest);
  const orc::LegendreGauss lg_sw_u = orc::lg_from_c(config->lg_sw_urban);
  const orc::LegendreGauss lg_lw_f = orc::lg_from_c(config->lg_lw_forest);
  const orc::LegendreGauss lg_lw_u = orc::lg_from_c(config->lg_lw_urban);
  if (nblocksize <= 0) nblocksize = 16;
  const int ncols = icol2 - icol1 + 1;
  const int nblock = (ncols + nblocksize - 1) / nblocksize;
  int rc_all = 0;
  bool simple_present = false;
  for (int j = icol1 - 1; j < icol2; ++j)
    if (cp->i_representation[j] >= SSB200_TILE_SIMPLE_URBAN) simple_present = true;
#ifdef _OPENMP
  int nt = nthreads > 0 ? nthreads : omp_get_max_threads();
#pragma omp parallel for schedule(dynamic) num_threads(nt)
#endif
  for (int jb = 0; jb < nblock; ++jb) {
    const int c1 = icol1 - 1 + jb * nblocksize;
    const int c2 = std::min(c1 + nblocksize, icol2);
    orc::g_flops = 0.0;
    for (int jcol = c1; jcol < c2; ++jcol) {
      int rc = orc::radsurf_column(*config, *cp, sw, lw, bc, jcol, sw_norm_dir, sw_norm_diff, lw_internal,
                                   lw_norm, lg_sw_f, lg_sw_u, lg_lw_f, lg_lw_u);
      if (rc) {
#ifdef _OPENMP
#pragma omp critical
#endif
        rc_all = rc;
      }
    }
    if (orc::g_count_flops) {
#ifdef _OPENMP
#pragma omp atomic
#endif
      orc::g_flops_total += orc::g_flops;
    }
  }
  (void)simple_present;
  return rc_all;
}

// Unit entry points used by tests (matrices are column-major, one interval).
int oracle_eigen_decomposition_real(int32_t n, const double *amat, double *eigenvalue, double *eigenvector) {
  orc::Mat A(n, n), V;
  orc::Vec w;
  orc::load_reals(A.a.data(), amat, n * n);
  int nerr = orc::eigen_decomposition_real(n, A, w, V);
  orc::store_reals(eigenvalue, w.data(), n);
  orc::store_reals(eigenvector, V.a.data(), n * n);
  return nerr;
}

int oracle_calc_matrices_sw_eig(int32_t ndiff, int32_t ndir, double dz, double mu0, const double *gamma0,
                                const double *gamma1, const double *gamma2, const double *gamma3,
                                double *reflectance, double *transmittance, double *s_up, double *s_dn,
                                double *trans_dir, double *int_dir, double *int_diff, double *int_dir_diff) {
  orc::Mat g0(ndir, ndir), g1(ndiff, ndiff), g2(ndiff, ndiff), g3(ndiff, ndir);
  orc::load_reals(g0.a.data(), gamma0, ndir * ndir);
  orc::load_reals(g1.a.data(), gamma1, ndiff * ndiff);
  orc::load_reals(g2.a.data(), gamma2, ndiff * ndiff);
  orc::load_reals(g3.a.data(), gamma3, ndiff * ndir);
  orc::Mat R, T, Su, Sd, E, Id, Idf, Idd;
  orc::calc_matrices_sw_eig(ndiff, ndir, dz, mu0, g0, g1, g2, g3, R, T, Su, Sd, E, Id, Idf, Idd);
  orc::store_reals(reflectance, R.a.data(), ndiff * ndiff);
  orc::store_reals(transmittance, T.a.data(), ndiff * ndiff);
  orc::store_reals(s_up, Su.a.data(), ndiff * ndir);
  orc::store_reals(s_dn, Sd.a.data(), ndiff * ndir);
  orc::store_reals(trans_dir, E.a.data(), ndir * ndir);
  orc::store_reals(int_dir, Id.a.data(), ndir * ndir);
  orc::store_reals(int_diff, Idf.a.data(), ndiff * ndiff);
  orc::store_reals(int_dir_diff, Idd.a.data(), ndiff * ndir);
  return 0;
}

int oracle_calc_matrices_lw_eig(int32_t n, double dz, const double *gamma1, const double *gamma2,
                                const double *emiss_rate, double *reflectance, double *transmittance,
                                double *source, double *int_flux, double *int_flux_source) {
  orc::Mat g1(n, n), g2(n, n);
  orc::load_reals(g1.a.data(), gamma1, n * n);
  orc::load_reals(g2.a.data(), gamma2, n * n);
  orc::Vec b(emiss_rate, emiss_rate + n), src, isrc;
  orc::Mat R, T, IF;
  orc::calc_matrices_lw_eig(n, dz, g1, g2, b, R, T, src, IF, isrc);
  orc::store_reals(reflectance, R.a.data(), n * n);
  orc::store_reals(transmittance, T.a.data(), n * n);
  orc::store_reals(source, src.data(), n);
  orc::store_reals(int_flux, IF.a.data(), n * n);
  orc::store_reals(int_flux_source, isrc.data(), n);
  return 0;
}

int oracle_schur_invert_sw(int32_t n0, int32_t n1, const double *g0, const double *g1, const double *g2,
                           const double *g3, double *g0i, double *g1i, double *g2i, double *g3i) {
  orc::Mat G0(n0, n0), G1(n1, n1), G2(n1, n1), G3(n1, n0), a, b, c, d;
  orc::load_reals(G0.a.data(), g0, n0 * n0);
  orc::load_reals(G1.a.data(), g1, n1 * n1);
  orc::load_reals(G2.a.data(), g2, n1 * n1);
  orc::load_reals(G3.a.data(), g3, n1 * n0);
  orc::schur_invert_sw(G0, G1, G2, G3, a, b, c, d);
  orc::store_reals(g0i, a.a.data(), n0 * n0);
  orc::store_reals(g1i, b.a.data(), n1 * n1);
  orc::store_reals(g2i, c.a.data(), n1 * n1);
  orc::store_reals(g3i, d.a.data(), n1 * n0);
  return 0;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
}
