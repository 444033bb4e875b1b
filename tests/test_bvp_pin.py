"""Independent pin of the oracle for the urban / vegetated-urban solvers
(SURVEY.md section 8c asked for a second implementation).

tests/bvp_reference.py shares nothing with oracle/src: geometry, Gamma matrices
and flux partition are written from the Fortran in Python, every layer is the
matrix exponential of its full rate matrix in 150-digit arithmetic (mpmath) and
integrated fluxes come from the dense inverse of that matrix.

 * mode "sic": those layer operators through the reference's sweep recurrences.
   Must agree with the _Float128 oracle (the parity truth) to 1e-11 on EVERY
   field - it does to ~1e-15.
 * mode "bvp": one global banded solve of all layers at once (no adding method
   at all).  It agrees with the truth on everything the upward sweep alone
   determines (top-of-canopy albedo / emissivity / emission, top_net) and, when
   the diffuse order is 1, on every field.  For order > 1 the reference's
   downward pass departs from the boundary-value solution by ~1e-4: it solves
       (I - a_above R) x = T v + ...            radsurf_urban_sw.F90:609,725-731
   where continuity of the fluxes at the layer base (x = T v + R a_above x)
   requires (I - R a_above); the two differ unless a_above and R commute.
   Parity is with the reference, so the CUDA path mirrors the reference here
   (SURVEY App. B "sic" list); this test pins the size of the effect (it is the
   source of the non-zero LW budget residual the reference's TODO:14 mentions).
"""
import numpy as np
import pytest

import bvp_reference as bvp
import golden_io
import oracle_lib
from spartacus_surface_b200.driver.spartacus_surface_driver import run_radsurf

LG = oracle_lib.legendre_gauss_init
TOL = 1e-11
UPWARD_ONLY = {"top_net", "top_dn", "top_dn_dir", "sw_albedo", "sw_albedo_dir", "lw_emissivity", "lw_emission"}


def _errors(r, truth, icol, band, mode):
    """{(object, field): error relative to max(|truth|, 1e-3 flux scale)} for one column."""
    cp, cfg = r.canopy_props, r.config
    il1, L = int(cp.istartlay[icol]) - 1, int(cp.nlay[icol])
    spec, nspec = (r.sw_spectral_props, cfg.nsw) if band == "sw" else (r.lw_spectral_props, cfg.nlw)
    errs = {}
    for g in range(nspec):
        out = bvp.solve_column(cfg, cp, spec, icol, g, band, mode=mode)
        for obj, fields in out.items():
            scale = 1.0 if obj == "bc" else max(abs(float(truth[obj][k][icol, g]))
                                                for k in ("top_dn", "top_net", "ground_dn", "ground_net"))
            for k, v in fields.items():
                t = truth[obj][k]
                if isinstance(v, list):
                    tv, bv = t[il1:il1 + L, g], np.array([float(x) for x in v])
                else:
                    tv, bv = np.array([t[icol, g]]), np.array([float(v)])
                e = float(np.abs(bv - tv).max() / max(np.abs(tv).max(), 1e-3 * scale))
                errs[(obj, k)] = max(errs.get((obj, k), 0.0), e)
    return errs


def _golden(case):
    r, _ = golden_io.load_case(case, legendre_gauss_init=LG)
    return r, golden_io.load_truth(case)


def _synthetic(ncol):
    from bench import make_config
    from spartacus_surface_b200.synthetic import make_synthetic
    from spartacus_surface_b200 import canopy_flux_type, boundary_conds_out_type
    from spartacus_surface_b200.driver.spartacus_surface_driver import DriverResult, allocate_outputs
    cfg = make_config(2)
    cfg.do_save_flux_profile = True
    cfg.consolidate(LG)
    r = DriverResult()
    r.config = cfg
    r.canopy_props, r.sw_spectral_props, r.lw_spectral_props = make_synthetic(cfg, ncol)
    r.top_flux_dn_sw = r.top_flux_dn_direct_sw = r.top_flux_dn_lw = None
    allocate_outputs(r)
    run_radsurf(r, oracle_lib.make_solver(quad=True))
    return r, golden_io.outputs_of(r)


@pytest.mark.parametrize("case,cols,bands", [
    ("simple_surfaces.npz", (2, 3), ("sw", "lw")),       # urban and vegetated urban, nreg 3, 2 streams
    ("simple_surfaces_1stream.npz", (2, 3), ("sw", "lw")),
    ("simple_noscat.npz", (2, 3), ("sw", "lw")),          # no scattering: the case the FP64 reference loses
    ("simple_empty_layers.npz", (1, 2, 3), ("sw", "lw")),  # clear-only sub-block aloft
    ("urban_2stream.npz", (0,), ("sw",)),                 # russell_square, 8 layers, nreg 2
])
def test_reference_recurrences_with_independent_operators(case, cols, bands):
    r, truth = _golden(case)
    for icol in cols:
        for band in bands:
            errs = _errors(r, truth, icol, band, "sic")
            worst = max(errs, key=lambda k: errs[k])
            assert errs[worst] <= TOL, (case, icol, band, worst, errs[worst])


def test_synthetic_baseline_columns():
    """Two columns of the BASELINE workload (16 layers, nreg 3, 2 streams, SW + LW): the truth the
    bench line is checked against is itself pinned."""
    r, truth = _synthetic(2)
    for icol in range(2):
        for band in ("sw", "lw"):
            errs = _errors(r, truth, icol, band, "sic")
            worst = max(errs, key=lambda k: errs[k])
            assert errs[worst] <= TOL, (icol, band, worst, errs[worst])


@pytest.mark.parametrize("case,icol,band", [("simple_surfaces.npz", 3, "sw"), ("simple_surfaces.npz", 3, "lw"),
                                            ("urban_2stream.npz", 0, "sw")])
def test_global_boundary_value_solve(case, icol, band):
    """No adding method anywhere: upward-sweep quantities agree to 1e-11; the interior fluxes of the
    reference differ from the boundary-value solution by 1e-6..1e-2 (see module docstring)."""
    r, truth = _golden(case)
    errs = _errors(r, truth, icol, band, "bvp")
    for (obj, k), e in errs.items():
        if k in UPWARD_ONLY:
            assert e <= TOL, (obj, k, e)
    interior = max(e for (obj, k), e in errs.items() if k not in UPWARD_ONLY)
    assert 1e-7 < interior < 5e-2, interior


def test_global_solve_equals_reference_when_matrices_commute():
    """Diffuse order 1 (one region, one stream): a_above and R are scalars and the reference's
    downward pass IS the boundary-value solution - every field to 1e-11."""
    r, truth = _golden("simple_surfaces_1stream.npz")
    for band in ("sw", "lw"):
        errs = _errors(r, truth, 2, band, "bvp")  # column 3 of test_surfaces_in.nc: unvegetated urban
        worst = max(errs, key=lambda k: errs[k])
        assert errs[worst] <= TOL, (band, worst, errs[worst])
