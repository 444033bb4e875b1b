"""A plain C program (gcc, tests/c_harness/harness.c) that includes include/spartacus_b200.h and
calls ssb200_radsurf the way the Fortran shim does.  CPU: it compiles, links and reports the
no-GPU code (there is no CPU fallback).  GPU: its output equals the oracle's on the same inputs."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_harness", "harness.c")
EXE = os.path.join(ROOT, "tests", "_build", "c_harness")
LIBDIR = os.path.join(ROOT, "spartacus_surface_b200", "csrc")


def _build():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"), SRC,
                           "-o", EXE, "-L", LIBDIR, "-lspartacus_b200", f"-Wl,-rpath,{LIBDIR}"])


def _run():
    return subprocess.run([EXE], capture_output=True, text=True)


def _have_gpu():
    from spartacus_surface_b200._lib import load
    return load().ssb200_device_count() > 0


def test_compiles_links_and_has_no_cpu_fallback():
    _build()
    if _have_gpu():
        pytest.skip("GPU present: covered by test_matches_oracle")
    p = _run()
    assert p.returncode == 3, (p.returncode, p.stderr)  # SSB200_ERR_NOGPU
    assert "returned -5" in p.stderr


@pytest.mark.gpu
def test_matches_oracle():
    import oracle_lib
    from spartacus_surface_b200 import (config_type, canopy_properties_type, sw_spectral_properties_type,
                                        lw_spectral_properties_type, canopy_flux_type, boundary_conds_out_type)
    from spartacus_surface_b200.radsurf_canopy_flux import ALL_FIELDS
    _build()
    p = _run()
    assert p.returncode == 0, p.stderr
    got = {}
    for line in p.stdout.splitlines():
        name, idx, val = line.split()
        got.setdefault(name, {})[int(idx)] = float(val)
    # the same closed-form inputs as harness.c
    ncol, ntot = 3, 5
    lay = lambda b, s: (b + s * np.arange(ntot))
    col = lambda b, s: (b + s * np.arange(ncol))
    cfg = config_type(do_sw=True, do_lw=True, nsw=1, nlw=1, n_vegetation_region_urban=2, n_vegetation_region_forest=2,
                      n_stream_sw_urban=2, n_stream_lw_urban=2, n_stream_sw_forest=2, n_stream_lw_forest=2)
    cfg = cfg.consolidate(oracle_lib.legendre_gauss_init)
    cp = canopy_properties_type()
    cp.set_layers(np.array([0, 3, 2], dtype=np.int32))
    cp.i_representation = np.array([0, 3, 1], dtype=np.int32)
    cp.cos_sza = col(0.5, 0.15)
    cp.dz, cp.building_fraction, cp.building_scale = lay(2.0, 0.5), lay(0.4, -0.05), lay(20.0, 1.0)
    cp.veg_fraction, cp.veg_scale, cp.veg_ext = lay(0.1, 0.02), lay(5.0, 1.0), lay(0.2, 0.05)
    cp.veg_fsd, cp.veg_contact_fraction = lay(0.6, 0.05), lay(0.2, 0.03)
    two = lambda a: np.ascontiguousarray(a.reshape(-1, 1))
    sw = sw_spectral_properties_type(1)
    sw.air_ext, sw.air_ssa, sw.veg_ssa = two(lay(1e-5, 0.0)), two(lay(0.999, 0.0)), two(lay(0.4, 0.05))
    sw.ground_albedo, sw.roof_albedo, sw.wall_albedo = two(col(0.2, 0.05)), two(lay(0.15, 0.02)), two(lay(0.3, 0.02))
    sw.wall_specular_frac = two(lay(0.0, 0.0))
    lw = lw_spectral_properties_type(1)
    lw.air_ext, lw.air_ssa, lw.clear_air_planck = two(lay(1e-5, 0.0)), two(lay(0.0, 0.0)), two(lay(340.0, 2.0))
    lw.veg_ssa, lw.veg_planck, lw.veg_air_planck = two(lay(0.03, 0.002)), two(lay(345.0, 2.0)), two(lay(342.0, 2.0))
    lw.ground_emissivity, lw.ground_emission = two(col(0.95, 0.01)), two(col(360.0, 5.0))
    lw.roof_emissivity, lw.wall_emissivity = two(lay(0.9, 0.01)), two(lay(0.92, 0.01))
    lw.roof_emission, lw.wall_emission = two(lay(350.0, 3.0)), two(lay(355.0, 3.0))
    bc = boundary_conds_out_type().allocate(ncol, 1, 1)
    fl = [canopy_flux_type().allocate(cfg, ncol, ntot, 1, use_direct=d, do_save_flux_profile=False)
          for d in (True, True, False, False)]
    # (the _Float128 build: the reference's own FP64 arithmetic is only good to ~1e-9 on these columns)
    assert oracle_lib.make_solver(quad=True)(cfg, cp, sw, lw, bc, None, None, *fl) == 0
    checked = 0
    for name, f in zip(("sw_norm_dir", "sw_norm_diff", "lw_internal", "lw_norm"), fl):
        for k in ALL_FIELDS:
            a = getattr(f, k)
            if a is None or f"{name}.{k}" not in got:
                continue
            g = np.array([got[f"{name}.{k}"][i] for i in range(a.shape[0])])
            scale = max(float(np.abs(a).max()), 1e-3)  # absorptions are ~1e-7 of the unit incoming flux (tests/parity.py)
            assert np.abs(g - a.reshape(-1)).max() <= 1e-9 * scale, (name, k)
            checked += 1
    for k in ("sw_albedo", "sw_albedo_dir", "lw_emissivity", "lw_emission"):
        g = np.array([got[f"bc.{k}"][i] for i in range(ncol)])
        assert np.allclose(g, getattr(bc, k).reshape(-1), rtol=1e-9), k
    assert checked >= 40
