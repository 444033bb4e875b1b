"""Output stage of the offline driver (spartacus_surface_b200/radsurf_save.py against
radsurf/radsurf_save.F90): file layout, ragged unpacking with the -9999 fill value, broadband sums.
CPU only: the fluxes come from the oracle through the driver mirror."""
import os

import numpy as np
import pytest
from scipy.io import netcdf_file

import golden_io
import oracle_lib
from spartacus_surface_b200.driver.spartacus_surface_driver import run_radsurf, scale_and_sum
from spartacus_surface_b200.radsurf_save import save_canopy_fluxes, FillValueFlux


def _case(name, tmp_path, **cfg_overrides):
    r, _ = golden_io.load_case(name, legendre_gauss_init=oracle_lib.legendre_gauss_init)
    for k, v in cfg_overrides.items():
        setattr(r.config, k, v)
    assert run_radsurf(r, oracle_lib.make_solver()) == 0
    scale_and_sum(r)
    path = os.path.join(tmp_path, name[:-4] + "_out.nc")
    save_canopy_fluxes(path, r.config, r.canopy_props, r.sw_flux, r.lw_flux)
    return r, path


def test_layout_and_values_ragged(tmp_path):
    r, path = _case("simple_surfaces.npz", str(tmp_path), do_save_spectral_flux=True)
    cp = r.canopy_props
    nmax = int(cp.nlay.max())
    with netcdf_file(path, "r", mmap=False) as f:
        assert f.dimensions["column"] == cp.ncol and f.dimensions["layer"] == nmax
        assert f.dimensions["layer_interface"] == nmax + 1 and f.dimensions["band_sw"] == r.config.nsw
        assert f.title.decode().startswith("Radiative fluxes from the SPARTACUS-Surface")
        assert np.array_equal(f.variables["surface_type"][:], cp.i_representation)
        assert np.array_equal(f.variables["nlayer"][:], cp.nlay)
        assert f.variables["nlayer"].data.dtype == np.dtype(">i2")
        h = f.variables["height"][:]
        for j in range(cp.ncol):
            l0, n = int(cp.istartlay[j]) - 1, int(cp.nlay[j])
            assert np.allclose(h[j, :n + 1], np.concatenate([[0.0], np.cumsum(cp.dz[l0:l0 + n])]), rtol=1e-6)
            assert np.all(h[j, n + 1:] == -1.0)
            # ragged unpack with the fill value (radsurf_save.F90:629-649)
            for var, member in (("clear_air_absorption_sw", "clear_air_abs"), ("wall_flux_net_sw", "wall_net"),
                                ("veg_absorption_direct_sw", "veg_abs_dir")):
                v = f.variables[var]
                assert v.data.dtype == np.dtype(">f4") and v._FillValue == np.float32(FillValueFlux)
                exp = getattr(r.sw_flux, member)[l0:l0 + n].sum(axis=1)
                assert np.allclose(v[j, :n], exp, rtol=2e-7, atol=1e-30)
                assert np.all(v[j, n:] == np.float32(FillValueFlux))
            s = f.variables["roof_spectral_flux_in_sw"]
            assert s.shape == (cp.ncol, nmax, r.config.nsw)
            assert np.allclose(s[j, :n], r.sw_flux.roof_in[l0:l0 + n], rtol=2e-7, atol=1e-30)
        assert np.allclose(f.variables["top_flux_net_sw"][:], r.sw_flux.top_net.sum(axis=1), rtol=2e-7)
        assert np.allclose(f.variables["ground_sunlit_fraction"][:], r.sw_flux.ground_sunlit_frac, rtol=2e-7)
        assert np.allclose(f.variables["ground_spectral_flux_vertical_lw"][:], r.lw_flux.ground_vertical_diff, rtol=2e-7)
        # the reference never sets do_broadband_lw (radsurf_save.F90:67-75): no broadband longwave variables
        assert "top_flux_net_lw" not in f.variables and "top_spectral_flux_net_lw" in f.variables
        assert "ground_flux_vertical_diffuse_sw" in f.variables and "ground_flux_vertical_sw" not in f.variables


def test_default_flags_broadband_only(tmp_path):
    r, path = _case("urban_2stream.npz", str(tmp_path))
    with netcdf_file(path, "r", mmap=False) as f:
        assert "band_sw" not in f.dimensions
        names = set(f.variables)
        assert {"height", "surface_type", "nlayer", "ground_flux_dn_sw", "roof_flux_in_direct_sw",
                "wall_sunlit_fraction", "veg_sunlit_fraction", "veg_air_absorption_sw"} <= names
        assert not any(n.endswith("_lw") for n in names)  # (App. B12)
        assert ("flux_dn_layer_top_sw" in names) == bool(r.config.do_save_flux_profile)


def test_single_layer_tiles(tmp_path):
    """Simple-urban / infinite-street columns (one layer each)."""
    r, path = _case("single_layer_exp.npz", str(tmp_path), do_save_flux_profile=False)
    with netcdf_file(path, "r", mmap=False) as f:
        assert f.variables["surface_type"][:].max() >= 4
        assert f.dimensions["layer"] == 1
