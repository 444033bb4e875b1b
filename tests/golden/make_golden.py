#!/usr/bin/env python
"""Generate tests/golden/*.npz from the reference's own test fixtures.

Runs in the build container only (needs /root/reference/test/*).  For each
case of the reference's `make test` suites it builds
the effective inputs exactly like the reference driver (namelist + netCDF +
overrides) and stores them with the CPU oracle's outputs.  The reference
ships no expected outputs, so the stored outputs are the oracle's (pinned to
the documentation's printed budget table by tests/test_oracle.py).

Usage: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib  # noqa: E402
import golden_io  # noqa: E402
from spartacus_surface_b200.driver.spartacus_surface_driver import (setup_case, allocate_outputs,  # noqa: E402
                                                                   run_radsurf)
from spartacus_surface_b200.radsurf_simple_spectrum import calc_simple_spectrum_lw  # noqa: E402

REF_TEST = "/root/reference/test"

# test/rami4pilps/duplicate_profiles.sh:19 (stored into a float32 variable by ncap2)
COS_SZA_46 = np.array([float(x) for x in (
    "1.0,0.999391,0.997564,0.994522,0.990268,0.984808,0.978148,0.970296,0.961262,0.951057,0.939693,"
    "0.927184,0.913545,0.898794,0.882948,0.866025,0.848048,0.829038,0.809017,0.788011,0.766044,0.743145,"
    "0.71934,0.694658,0.669131,0.642788,0.615661,0.587785,0.559193,0.529919,0.5,0.469472,0.438371,0.406737,"
    "0.374607,0.34202,0.309017,0.275637,0.241922,0.207912,0.173648,0.139173,0.104528,0.0697565,0.0348995,0.01"
).split(",")], dtype=np.float32).astype(np.float64)


def replicate_columns(r, cos_sza):
    """duplicate_profiles.sh: the single column repeated once per solar zenith angle."""
    n = len(cos_sza)
    cp = r.canopy_props
    assert cp.ncol == 1
    nl = int(cp.nlay[0])
    for obj in (cp, r.sw_spectral_props, r.lw_spectral_props):
        for k, v in list(vars(obj).items()):
            if isinstance(v, np.ndarray) and k not in ("nlay", "istartlay", "i_representation"):
                reps = (n,) + (1,) * (v.ndim - 1)
                setattr(obj, k, np.ascontiguousarray(np.tile(v, reps)))
    cp.i_representation = np.tile(cp.i_representation, n).astype(np.int32)
    cp.set_layers(np.full(n, nl, dtype=np.int32))
    if r.driver_config.cos_sza_override < 0.0:
        cp.cos_sza = np.array(cos_sza, dtype=np.float64)
    for name in ("top_flux_dn_sw", "top_flux_dn_direct_sw", "top_flux_dn_lw"):
        v = getattr(r, name)
        if v is not None:
            setattr(r, name, np.ascontiguousarray(np.tile(v, (n, 1))))
    allocate_outputs(r)


def rami4pilps_args(band, surf, frac, vregs=None, streams=None):
    """test/rami4pilps/expand_args.sh:19-53."""
    ssa = {"vis": 0.1301, "nir": 0.8058}[band]
    albedo = {("vis", "med"): 0.1217, ("vis", "snw"): 0.9640, ("nir", "med"): 0.2142, ("nir", "snw"): 0.5568}[(band, surf)]
    drv = dict(vegetation_fraction=frac, ground_sw_albedo=albedo, vegetation_sw_ssa=ssa)
    rad = {}
    if vregs:
        rad["n_vegetation_region_forest"] = vregs
    if streams:
        rad["n_stream_sw_forest"] = streams
    return rad, drv


def cases():
    S = REF_TEST + "/simple/"
    for name in ("surfaces", "consistency", "empty_layers", "nearly_empty_layers", "noscat", "overhang", "closed"):
        yield f"simple_{name}", S + "config.nam", S + f"test_{name}_in.nc", {}, {}, None
    yield "simple_surfaces_1stream", S + "config_1stream.nam", S + "test_surfaces_in.nc", {}, {}, None
    yield ("simple_surfaces_doc", S + "config.nam", S + "test_surfaces_in.nc", {},
           dict(vegetation_extinction=0.25), None)
    R = REF_TEST + "/rami4pilps/"
    for vregs in (1, 2):
        for streams in (1, 2, 4, 8):
            rad, drv = rami4pilps_args("vis", "snw", 0.3, vregs, streams)
            yield f"rami4pilps_vis-snw-0.3-{vregs}-{streams}", R + "config.nam", R + "rami4pilps_base_profile.nc", rad, drv, COS_SZA_46
    for band, surf, frac in (("vis", "med", 0.1), ("nir", "med", 0.5), ("nir", "snw", 0.3)):
        rad, drv = rami4pilps_args(band, surf, frac)
        yield f"rami4pilps_{band}-{surf}-{frac}", R + "config.nam", R + "rami4pilps_base_profile.nc", rad, drv, COS_SZA_46
    U = REF_TEST + "/urban/"
    yield "urban_single", U + "config.nam", U + "russell_square.nc", {}, dict(cos_solar_zenith_angle=0.5), None
    for ns in (1, 2, 4):
        yield (f"urban_{ns}stream", U + "config.nam", U + "russell_square.nc",
               dict(n_stream_sw_urban=ns, n_stream_lw_urban=ns), {}, COS_SZA_46)
    M = REF_TEST + "/rami5/"
    scenes = {"HET09_JBS_SUM": (56, 41), "HET15_JBS_WIN": (76,), "HET07_JPS_SUM": (0,), "HET08_OPS_WIN": (47,),
              "HET14_WCO_UND": (42, 67)}
    for scene, szas in scenes.items():
        nc = M + f"scene_nc/rami5_{scene}_scene.nc"
        yield f"rami5_{scene}-diffuse", M + "config.nam", nc, {}, {}, None
        for sza in szas:
            yield (f"rami5_{scene}-{sza:02d}-direct", M + "config.nam", nc, {},
                   dict(top_flux_dn_direct_sw=1.0, solar_zenith_angle=sza), None)
    yield ("rami5_HET09_JBS_SUM-56-direct-blacksoil", M + "config.nam", M + "scene_nc/rami5_HET09_JBS_SUM_scene.nc", {},
           dict(top_flux_dn_direct_sw=1.0, solar_zenith_angle=56, ground_sw_albedo=0.0), None)
    L = REF_TEST + "/single_layer/"
    for tag, typ in (("sp", 2), ("exp", 4), ("inf", 5)):
        yield f"single_layer_{tag}", L + "config.nam", L + "test_single_layer.nc", {}, dict(isurfacetype=typ), None


def main():
    solver = oracle_lib.make_solver()
    for old in golden_io.list_cases():
        os.remove(os.path.join(golden_io.GOLDEN_DIR, old))
    for name, nam, nc, rad, drv, dup in cases():
        r = setup_case(nam, nc, rad, drv, legendre_gauss_init=oracle_lib.legendre_gauss_init)
        if dup is not None:
            replicate_columns(r, dup)
            if r.config.do_lw:
                calc_simple_spectrum_lw(r.config, r.canopy_props, r.lw_spectral_props)
        rc = run_radsurf(r, solver)
        golden_io.save_case(os.path.join(golden_io.GOLDEN_DIR, name + ".npz"), r,
                            meta=dict(namelist=os.path.relpath(nam, "/root/reference"),
                                      input=os.path.relpath(nc, "/root/reference"),
                                      radsurf_overrides=rad, driver_overrides=drv,
                                      sza_duplicated=dup is not None, oracle_status=rc))
        bad = [k for f in golden_io.outputs_of(r).values() for k, v in f.items() if not np.all(np.isfinite(v))]
        print(f"{name:45s} ncol={r.canopy_props.ncol:3d} ntotlay={r.canopy_props.ntotlay:4d} rc={rc} nonfinite={bad}")


if __name__ == "__main__":
    main()
