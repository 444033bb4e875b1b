"""read_input (driver/spartacus_surface_read_input.F90:20-384) on top of
scipy.io.netcdf_file (all reference fixtures are netCDF-3 classic).

Reproduces the reader's rules: float32 file data promoted to float64, ragged
packing with `nlayer` (:417-496), namelist overrides (a value >= 0 replaces
the whole array), the defaults for absent variables (veg_contact_fraction
:155-166, air temperatures :228-257, roof_sw_albedo_direct :311-324,
wall_sw_specular_fraction :335-344), the hard-coded air optics (:258-269,
:362-365) and the `top_flux_lw_sw` typo that makes the LW top flux always
come from sky_temperature (:273-280).

File dimension order (column, layer[, wavelength]) is the reverse of the
Fortran order, so a packed Fortran (nspec, ntotlay) array is the C array
(ntotlay, nspec) built here.
"""
import numpy as np
from scipy.io import netcdf_file

from ..radsurf_canopy_properties import canopy_properties_type
from ..radsurf_sw_spectral_properties import sw_spectral_properties_type
from ..radsurf_lw_spectral_properties import lw_spectral_properties_type, StefanBoltzmann


class _File:
    def __init__(self, path):
        self.nc = netcdf_file(path, "r", mmap=False)

    def exists(self, name):
        return name in self.nc.variables

    def get(self, name):
        if name not in self.nc.variables:
            raise KeyError(f"variable '{name}' not found in input file")
        return np.array(self.nc.variables[name].data, dtype=np.float64)

    def rank(self, name):
        return len(self.nc.variables[name].shape)

    def close(self):
        self.nc.close()


def _pack_1d(f, name, nlay):
    v = f.get(name)  # (column, layer)
    return np.ascontiguousarray(np.concatenate([v[j, :n] for j, n in enumerate(nlay)]) if len(nlay) else np.zeros(0))


def _pack_2d(f, name, nlay):
    v = f.get(name)
    if f.rank(name) == 2:  # (column, layer) -> (ntotlay, 1)
        return _pack_1d(f, name, nlay)[:, None].copy()
    return np.ascontiguousarray(np.concatenate([v[j, :n, :] for j, n in enumerate(nlay)], axis=0))


def _read_2d(f, name):
    v = f.get(name)
    if v.ndim == 1:  # (column) -> (ncol, 1)
        return v[:, None].copy()
    return np.ascontiguousarray(v)


def read_input(path, config, driver_config):
    """Return (canopy_props, sw_spectral_props, lw_spectral_props,
    top_flux_dn_sw, top_flux_dn_direct_sw, top_flux_dn_lw); the top fluxes
    are (ncol, nspec) arrays or None."""
    f = _File(path)
    dc = driver_config
    cp = canopy_properties_type()
    sw = sw_spectral_properties_type(config.nsw)
    lw = lw_spectral_properties_type(config.nlw)
    top_sw = top_sw_dir = top_lw = None
    try:
        nlay = f.get("nlayer").astype(np.int32)
        cp.set_layers(nlay)
        ncol, ntotlay = cp.ncol, cp.ntotlay
        full_lay = lambda val: np.full(ntotlay, float(val))
        if config.do_sw:
            if dc.cos_sza_override >= 0.0:
                cp.cos_sza = np.full(ncol, float(dc.cos_sza_override))
            else:
                cp.cos_sza = f.get("cos_solar_zenith_angle")
        height = f.get("height")
        cp.dz = np.ascontiguousarray(np.concatenate(
            [height[j, 1:n + 1] - height[j, 0:n] for j, n in enumerate(nlay)]))
        if dc.isurfacetype >= 0:
            cp.i_representation = np.full(ncol, dc.isurfacetype, dtype=np.int32)
        else:
            cp.i_representation = f.get("surface_type").astype(np.int32)
        if config.do_urban:
            cp.building_fraction = _pack_1d(f, "building_fraction", nlay)
            cp.building_scale = _pack_1d(f, "building_scale", nlay)
        if config.do_vegetation:
            if dc.vegetation_fraction >= 0.0:
                cp.veg_fraction = full_lay(dc.vegetation_fraction)
            else:
                cp.veg_fraction = _pack_1d(f, "veg_fraction", nlay)
            cp.veg_ext = _pack_1d(f, "veg_extinction", nlay)
            if dc.vegetation_extinction >= 0.0:
                cp.veg_ext = full_lay(dc.vegetation_extinction)
            elif dc.vegetation_extinction_scaling >= 0.0:
                cp.veg_ext = cp.veg_ext * dc.vegetation_extinction_scaling
            cp.veg_scale = _pack_1d(f, "veg_scale", nlay)
            if dc.vegetation_fsd >= 0.0:
                cp.veg_fsd = full_lay(dc.vegetation_fsd)
            else:
                cp.veg_fsd = _pack_1d(f, "veg_fsd", nlay)
            if config.do_urban:
                if f.exists("veg_contact_fraction"):
                    cp.veg_contact_fraction = _pack_1d(f, "veg_contact_fraction", nlay)
                else:
                    cp.veg_contact_fraction = np.minimum(
                        1.0, cp.veg_fraction / np.maximum(config.min_vegetation_fraction,
                                                          1.0 - cp.building_fraction))
        if config.do_lw:
            cp.ground_temperature = f.get("ground_temperature")
            if config.do_urban:
                cp.roof_temperature = _pack_1d(f, "roof_temperature", nlay)
                cp.wall_temperature = _pack_1d(f, "wall_temperature", nlay)
            lw.ground_emissivity = _read_2d(f, "ground_lw_emissivity")
            if dc.ground_lw_emissivity >= 0.0:
                lw.ground_emissivity[...] = dc.ground_lw_emissivity
            if config.do_urban:
                lw.roof_emissivity = _pack_2d(f, "roof_lw_emissivity", nlay)
                if dc.roof_lw_emissivity >= 0.0:
                    lw.roof_emissivity[...] = dc.roof_lw_emissivity
                lw.wall_emissivity = _pack_2d(f, "wall_lw_emissivity", nlay)
                if dc.wall_lw_emissivity >= 0.0:
                    lw.wall_emissivity[...] = dc.wall_lw_emissivity
            if config.do_vegetation:
                lw.veg_ssa = _pack_2d(f, "veg_lw_ssa", nlay)
                if dc.vegetation_lw_ssa >= 0.0:
                    lw.veg_ssa[...] = dc.vegetation_lw_ssa
            if config.do_vegetation or config.do_urban:
                if f.exists("clear_air_temperature"):
                    cp.clear_air_temperature = _pack_1d(f, "clear_air_temperature", nlay)
                    if config.do_vegetation:
                        cp.veg_air_temperature = _pack_1d(f, "veg_air_temperature", nlay)
                else:
                    cp.clear_air_temperature = _pack_1d(f, "air_temperature", nlay)
                    if config.do_vegetation:
                        cp.veg_air_temperature = cp.clear_air_temperature.copy()
                if config.do_vegetation:
                    if f.exists("veg_temperature"):
                        cp.veg_temperature = _pack_1d(f, "veg_temperature", nlay)
                    else:
                        cp.veg_temperature = cp.clear_air_temperature.copy()
                lw.air_ext = np.full((ntotlay, config.nlw), 1.0e-5)
                lw.air_ssa = np.zeros((ntotlay, config.nlw))
                lw.clear_air_planck = np.zeros((ntotlay, config.nlw))
                if config.do_vegetation:
                    lw.veg_planck = np.zeros((ntotlay, config.nlw))
                    lw.veg_air_planck = np.zeros((ntotlay, config.nlw))
            if f.exists("top_flux_lw_sw"):  # sic (:273)
                top_lw = _read_2d(f, "top_flux_dn_lw")
            else:
                top_lw = StefanBoltzmann * _read_2d(f, "sky_temperature") ** 4
        if config.do_sw:
            sw.ground_albedo = _read_2d(f, "ground_sw_albedo")
            if dc.ground_sw_albedo >= 0.0:
                sw.ground_albedo[...] = dc.ground_sw_albedo
            if f.exists("ground_sw_albedo_direct"):
                sw.ground_albedo_dir = _read_2d(f, "ground_sw_albedo_direct")
            if config.do_urban:
                sw.roof_albedo = _pack_2d(f, "roof_sw_albedo", nlay)
                if dc.roof_sw_albedo >= 0.0:
                    sw.roof_albedo[...] = dc.roof_sw_albedo
                if f.exists("roof_sw_albedo_direct"):
                    sw.roof_albedo_dir = _pack_2d(f, "roof_sw_albedo_direct", nlay)
                else:
                    sw.roof_albedo_dir = sw.roof_albedo.copy()
                    if dc.roof_sw_albedo >= 0.0:
                        sw.roof_albedo_dir[...] = dc.roof_sw_albedo
                sw.wall_albedo = _pack_2d(f, "wall_sw_albedo", nlay)
                if dc.wall_sw_albedo >= 0.0:
                    sw.wall_albedo[...] = dc.wall_sw_albedo
                if f.exists("wall_sw_specular_fraction"):
                    sw.wall_specular_frac = _pack_2d(f, "wall_sw_specular_fraction", nlay)
                else:
                    sw.wall_specular_frac = np.zeros((ntotlay, sw.roof_albedo.shape[1]))
            if config.do_vegetation:
                sw.veg_ssa = _pack_2d(f, "veg_sw_ssa", nlay)
                if dc.vegetation_sw_ssa >= 0.0:
                    sw.veg_ssa[...] = dc.vegetation_sw_ssa
            if config.do_vegetation or config.do_urban:
                sw.air_ext = np.full((ntotlay, config.nsw), 1.0e-5)
                sw.air_ssa = np.full((ntotlay, config.nsw), 0.999)
            if dc.top_flux_dn_sw >= 0.0:
                top_sw = np.full((ncol, config.nsw), float(dc.top_flux_dn_sw))
            else:
                top_sw = _read_2d(f, "top_flux_dn_sw")
            if dc.top_flux_dn_direct_sw >= 0.0:
                top_sw_dir = np.full((ncol, config.nsw), float(dc.top_flux_dn_direct_sw))
            else:
                top_sw_dir = _read_2d(f, "top_flux_dn_direct_sw")
    finally:
        f.close()
    return cp, sw, lw, top_sw, top_sw_dir, top_lw
