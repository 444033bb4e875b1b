"""Degenerate and near-degenerate layer problems (SURVEY section 7 "hard parts"):
what the symmetrised Jacobi eigen-solver and the guarded particular solutions of
csrc/ssb_layer_math.cuh must survive.  Shared by the host-check (CPU) and the GPU parity tests.

Groups of columns cut out of the regular synthetic vegetated-urban canopy (8 layers):
  0  untouched (control)
  1  veg_ext = 0 with a vegetation fraction of ~half the open area: the vegetated regions are
     optically identical to clear air (test/simple/test_surfaces_in.nc column 2 layer 1; TODO:13)
  2  veg_fsd = 0: the two vegetated regions are identical (od_scaling = 1 for both)
  3  cos_sza = 0.01: grazing sun
  4  optically thick: dz x 25 and veg_ext x 8 (exp(-lambda dz) underflows, like test/rami5)
  5  veg_fraction = 2e-6 of the layer (just above min_vegetation_fraction): regions differing by
     1e6 in area
  6  forest tiles with the settings of groups 1 and 2
"""
import numpy as np

from spartacus_surface_b200 import config_type, canopy_flux_type, boundary_conds_out_type
from spartacus_surface_b200.radsurf_canopy_flux import ALL_FIELDS
from spartacus_surface_b200.radsurf_canopy_properties import ITileForest
from spartacus_surface_b200.synthetic import make_synthetic

NLAY = 8
NGROUP = 7
PER_GROUP = 12
BC_FIELDS = ("sw_albedo", "sw_albedo_dir", "lw_emissivity", "lw_emission")


def make_degenerate(streams, legendre_gauss_init=None):
    cfg = config_type(do_sw=True, do_lw=True, nsw=1, nlw=1, n_vegetation_region_urban=2,
                      n_vegetation_region_forest=2, n_stream_sw_urban=streams, n_stream_lw_urban=streams,
                      n_stream_sw_forest=streams, n_stream_lw_forest=streams)
    cfg = cfg.consolidate(legendre_gauss_init) if legendre_gauss_init else cfg.consolidate()
    ncol = NGROUP * PER_GROUP
    cp, sw, lw = make_synthetic(cfg, ncol, NLAY, seed=424242)
    grp = np.repeat(np.arange(ncol) // PER_GROUP, NLAY)  # group of every packed layer
    colgrp = np.arange(ncol) // PER_GROUP
    for obj in (cp, sw, lw):
        for k, v in list(vars(obj).items()):
            if isinstance(v, np.ndarray) and v.dtype == np.float64:
                setattr(obj, k, v.copy())
    open_frac = 1.0 - cp.building_fraction
    m = (grp == 1) | (grp == 6)
    cp.veg_fraction[m] = 0.5 * open_frac[m]
    cp.veg_ext[grp == 1] = 0.0
    cp.veg_ext[(grp == 6) & (np.arange(grp.size) % 2 == 0)] = 0.0
    cp.veg_fsd[(grp == 2) | (grp == 6)] = 0.0
    cp.veg_fraction[grp == 2] = np.maximum(cp.veg_fraction[grp == 2], 0.2 * open_frac[grp == 2])
    cp.cos_sza[colgrp == 3] = 0.01
    cp.dz[grp == 4] *= 25.0
    cp.veg_ext[grp == 4] *= 8.0
    cp.veg_fraction[grp == 4] = np.maximum(cp.veg_fraction[grp == 4], 0.2 * open_frac[grp == 4])
    cp.veg_fraction[grp == 5] = 2.0e-6
    cp.veg_contact_fraction = np.minimum(1.0, cp.veg_fraction / np.maximum(1.0e-6, open_frac))
    cp.i_representation = cp.i_representation.copy()
    cp.i_representation[colgrp == 6] = ITileForest
    return cfg, cp, sw, lw


def run_all(cfg, cp, sw, lw, solver):
    bc = boundary_conds_out_type().allocate(cp.ncol, cfg.nsw, cfg.nlw)
    fl = [canopy_flux_type().allocate(cfg, cp.ncol, cp.ntotlay, 1, use_direct=d) for d in (True, True, False, False)]
    status = solver(cfg, cp, sw, lw, bc, None, None, *fl)
    out = {n: {k: getattr(f, k) for k in ALL_FIELDS if getattr(f, k) is not None}
           for n, f in zip(("sw_norm_dir", "sw_norm_diff", "lw_internal", "lw_norm"), fl)}
    out["bc"] = {k: getattr(bc, k) for k in BC_FIELDS}
    return out, status
