/* harness.c - TEST INFRASTRUCTURE: a plain C caller of libspartacus_b200.so, compiled by gcc
 * against include/spartacus_b200.h.  It does what the Fortran shim
 * (spartacus_surface_b200/fortran/radsurf_interface_b200.F90) does for the reference's
 * `radsurf(config, canopy_props, sw, lw, bc_out, istartcol, iendcol, sw_norm_dir, sw_norm_diff,
 * lw_internal, lw_norm)` (radsurf/radsurf_interface.F90:20-25): fills the C structs with the
 * addresses of caller-owned column-major arrays (NULL = member not allocated) and makes ONE call.
 * No Fortran compiler exists in this image, so this is the closest compiled check of the header and
 * of the calling convention the shim relies on.
 *
 * Input: a small closed-form canopy (3 columns: flat, vegetated urban with 3 layers, forest with 2)
 * that tests/test_c_harness.py rebuilds with numpy for the oracle.  Output: "name index value" lines.
 * Exit code: the return value of ssb200_radsurf mapped to 0 (ok), 3 (no GPU), 1 (anything else). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "spartacus_b200.h"

#define NCOL 3
#define NTOT 5

static double *lay(double base, double step) {
  double *p = (double *)malloc(sizeof(double) * NTOT);
  for (int i = 0; i < NTOT; ++i) p[i] = base + step * i;
  return p;
}
static double *col(double base, double step) {
  double *p = (double *)malloc(sizeof(double) * NCOL);
  for (int i = 0; i < NCOL; ++i) p[i] = base + step * i;
  return p;
}
static double *zl(void) { return (double *)calloc(NTOT, sizeof(double)); }
static double *zc(void) { return (double *)calloc(NCOL, sizeof(double)); }

static void alloc_flux(ssb200_canopy_flux *f, int direct) {
  memset(f, 0, sizeof(*f));
  f->nspec = 1;
  f->ncol = NCOL;
  f->ntotlay = NTOT;
  f->ground_dn = zc(); f->ground_net = zc(); f->ground_vertical_diff = zc(); f->top_dn = zc(); f->top_net = zc();
  f->roof_in = zl(); f->roof_net = zl(); f->wall_in = zl(); f->wall_net = zl();
  f->clear_air_abs = zl(); f->veg_abs = zl(); f->veg_air_abs = zl();
  if (direct) {
    f->ground_dn_dir = zc(); f->top_dn_dir = zc(); f->ground_sunlit_frac = zc();
    f->roof_in_dir = zl(); f->wall_in_dir = zl(); f->roof_sunlit_frac = zl(); f->wall_sunlit_frac = zl();
    f->veg_abs_dir = zl(); f->veg_sunlit_frac = zl();
  }
}
static void dump(const char *obj, const ssb200_canopy_flux *f) {
#define DC(m) if (f->m) for (int i = 0; i < NCOL; ++i) printf("%s.%s %d %.17g\n", obj, #m, i, f->m[i]);
#define DL(m) if (f->m) for (int i = 0; i < NTOT; ++i) printf("%s.%s %d %.17g\n", obj, #m, i, f->m[i]);
  DC(ground_dn) DC(ground_net) DC(ground_vertical_diff) DC(top_dn) DC(top_net) DC(ground_dn_dir) DC(top_dn_dir)
  DC(ground_sunlit_frac)
  DL(roof_in) DL(roof_net) DL(wall_in) DL(wall_net) DL(roof_in_dir) DL(wall_in_dir) DL(roof_sunlit_frac)
  DL(wall_sunlit_frac) DL(clear_air_abs) DL(veg_abs) DL(veg_air_abs) DL(veg_abs_dir) DL(veg_sunlit_frac)
}

int main(void) {
  int64_t sizes[7];
  if (ssb200_abi_sizes(sizes) != 0 || sizes[1] != (int64_t)sizeof(ssb200_config) ||
      sizes[5] != (int64_t)sizeof(ssb200_canopy_flux)) {
    fprintf(stderr, "harness: ABI size mismatch\n");
    return 1;
  }
  ssb200_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.do_sw = cfg.do_lw = 1;
  cfg.do_vegetation = cfg.do_urban = 1;
  cfg.n_vegetation_region_forest = cfg.n_vegetation_region_urban = 2;
  cfg.nsw = cfg.nlw = 1;
  cfg.use_symmetric_vegetation_scale_forest = cfg.use_symmetric_vegetation_scale_urban = 1;
  cfg.min_vegetation_fraction = 1.0e-6;
  cfg.min_building_fraction = 1.0e-6;
  ssb200_legendre_gauss_init(2, &cfg.lg_sw_forest);
  cfg.lg_sw_urban = cfg.lg_lw_forest = cfg.lg_lw_urban = cfg.lg_sw_forest;

  static const int32_t nlay[NCOL] = {0, 3, 2}, istartlay[NCOL] = {1, 1, 4}, irep[NCOL] = {0, 3, 1};
  ssb200_canopy_properties cp;
  memset(&cp, 0, sizeof(cp));
  cp.ncol = NCOL;
  cp.ntotlay = NTOT;
  cp.nlay = nlay;
  cp.istartlay = istartlay;
  cp.i_representation = irep;
  cp.cos_sza = col(0.5, 0.15);
  cp.dz = lay(2.0, 0.5);
  cp.building_fraction = lay(0.4, -0.05);
  cp.building_scale = lay(20.0, 1.0);
  cp.veg_fraction = lay(0.1, 0.02);
  cp.veg_scale = lay(5.0, 1.0);
  cp.veg_ext = lay(0.2, 0.05);
  cp.veg_fsd = lay(0.6, 0.05);
  cp.veg_contact_fraction = lay(0.2, 0.03);

  ssb200_sw_spectral_properties sw;
  memset(&sw, 0, sizeof(sw));
  sw.nspec = 1;
  sw.air_ext = lay(1.0e-5, 0.0);
  sw.air_ssa = lay(0.999, 0.0);
  sw.veg_ssa = lay(0.4, 0.05);
  sw.ground_albedo = col(0.2, 0.05);
  sw.roof_albedo = lay(0.15, 0.02);
  sw.wall_albedo = lay(0.3, 0.02);
  sw.wall_specular_frac = lay(0.0, 0.0);

  ssb200_lw_spectral_properties lw;
  memset(&lw, 0, sizeof(lw));
  lw.nspec = 1;
  lw.air_ext = lay(1.0e-5, 0.0);
  lw.air_ssa = lay(0.0, 0.0);
  lw.clear_air_planck = lay(340.0, 2.0);
  lw.veg_ssa = lay(0.03, 0.002);
  lw.veg_planck = lay(345.0, 2.0);
  lw.veg_air_planck = lay(342.0, 2.0);
  lw.ground_emissivity = col(0.95, 0.01);
  lw.ground_emission = col(360.0, 5.0);
  lw.roof_emissivity = lay(0.9, 0.01);
  lw.wall_emissivity = lay(0.92, 0.01);
  lw.roof_emission = lay(350.0, 3.0);
  lw.wall_emission = lay(355.0, 3.0);

  ssb200_boundary_conds_out bc;
  bc.sw_albedo = zc(); bc.sw_albedo_dir = zc(); bc.lw_emissivity = zc(); bc.lw_emission = zc();
  ssb200_canopy_flux f[4];
  alloc_flux(&f[0], 1);
  alloc_flux(&f[1], 1);
  alloc_flux(&f[2], 0);
  alloc_flux(&f[3], 0);

  const int rc = ssb200_radsurf(&cfg, &cp, &sw, &lw, &bc, 0, 0, &f[0], &f[1], &f[2], &f[3]);
  if (rc != 0) {
    fprintf(stderr, "harness: ssb200_radsurf returned %d: %s\n", rc, ssb200_last_error());
    return rc == SSB200_ERR_NOGPU ? 3 : 1;
  }
  for (int i = 0; i < NCOL; ++i)
    printf("bc.sw_albedo %d %.17g\nbc.sw_albedo_dir %d %.17g\nbc.lw_emissivity %d %.17g\nbc.lw_emission %d %.17g\n", i,
           bc.sw_albedo[i], i, bc.sw_albedo_dir[i], i, bc.lw_emissivity[i], i, bc.lw_emission[i]);
  dump("sw_norm_dir", &f[0]);
  dump("sw_norm_diff", &f[1]);
  dump("lw_internal", &f[2]);
  dump("lw_norm", &f[3]);
  return 0;
}
