"""save_canopy_fluxes: the output stage of the offline driver (radsurf/radsurf_save.F90:26-166,
define_canopy_flux_variables :168-432, write_canopy_flux_variables :435-627, the three
unpack_variable* helpers :629-693).

Same file layout as the reference writes (readable by its Matlab / NCO post-processing):
dimensions `column`, `layer` (= max number of layers), `layer_interface` and, for spectral
output, `band_sw` / `band_lw`; `height`, `surface_type`, `nlayer`; per band the broadband
(sum over spectral intervals) and / or spectral variables of every ALLOCATED member of the
summed flux object, ragged per-layer members unpacked to (column, layer) with the fill value
-9999 in unused slots; single-precision reals like easy_netcdf's default (NF90_FLOAT),
short integers for surface_type / nlayer.  netCDF-3 classic through scipy.io.netcdf_file
(the reference's `is_hdf5_file` option is not offered: no HDF5 library in this environment).

Reference behaviours kept: `do_broadband_lw` is never set by the reference (it assigns
do_broadband_sw twice, radsurf_save.F90:67-75), so broadband LONGWAVE
variables are not written unless `fix_broadband_lw=True`.
"""
import numpy as np

FillValueFlux = -9999.0

TITLE = "Radiative fluxes from the SPARTACUS-Surface radiation model"
REFERENCES = ("Hogan, R. J., T. Quaife and R. Braghiere, 2018: Fast matrix treatment of 3-D radiative"
              " transfer in vegetation canopies: SPARTACUS-Vegetation 1.1. Geosci. Model Dev., 11, 339-350.\n"
              "Hogan, R. J., 2019: Flexible treatment of radiative transfer in complex urban"
              " canopies for use in weather and climate models. Boundary-Layer Meteorol., 173, 53-78.")
SOURCE = "SPARTACUS-Surface offline radiation model"
COMMENT = ("All fluxes and absorption rates are in terms of power per unit horizontal area of the domain. "
           "Net fluxes are downwelling (or incoming) minus upwelling (or outgoing).")
SURFACE_TYPES = ("0: Flat\n1: Forest\n2: Unvegetated urban\n3: Vegetated urban\n4: Simple urban\n"
                 "5: Infinite street")


def _host(a):
    if a is None:
        return None
    if type(a).__module__.startswith("torch"):
        a = a.cpu().numpy()
    return np.asarray(a)


def _column_table(flux, band_long):
    """(member, broadband stem, spectral stem, long name) of the per-column members, in the
    reference's order (radsurf_save.F90:198-225)."""
    direct = flux.ground_dn_dir is not None
    t = [("ground_dn", "ground_flux_dn_", "ground_spectral_flux_dn_", f"Downwelling {band_long} {{}}flux at ground"),
         ("ground_net", "ground_flux_net_", "ground_spectral_flux_net_", f"Net {band_long} {{}}flux at ground")]
    if direct:
        t += [("ground_dn_dir", "ground_flux_dn_direct_", "ground_spectral_flux_dn_direct_",
               f"Downwelling direct {band_long} {{}}flux at ground"),
              ("ground_vertical_diff", "ground_flux_vertical_diffuse_", "ground_spectral_flux_vertical_diffuse_",
               f"Diffuse {band_long} {{}}flux into a vertical surface at ground level")]
    else:
        t += [("ground_vertical_diff", "ground_flux_vertical_", "ground_spectral_flux_vertical_",
               f"Flux in {band_long} into a vertical surface at ground level")]
    t += [("top_dn", "top_flux_dn_", "top_spectral_flux_dn_", f"Downwelling {band_long} {{}}flux at top of canopy"),
          ("top_net", "top_flux_net_", "top_spectral_flux_net_", f"Net {band_long} {{}}flux at top of canopy")]
    if flux.top_dn_dir is not None:
        t += [("top_dn_dir", "top_flux_dn_direct_", "top_spectral_flux_dn_direct_",
               f"Downwelling direct {band_long} {{}}flux at top of canopy")]
    return t


def _layer_table(flux, band_long):
    """The per-layer members (radsurf_save.F90:226-293)."""
    t = []
    if flux.roof_in is not None:
        t.append(("roof_in", "roof_flux_in_", "roof_spectral_flux_in_", f"Incoming {band_long} {{}}flux at roofs"))
        if flux.roof_in_dir is not None:
            t.append(("roof_in_dir", "roof_flux_in_direct_", "roof_spectral_flux_in_direct_",
                      f"Direct incoming {band_long} {{}}flux at roofs"))
        t.append(("roof_net", "roof_flux_net_", "roof_spectral_flux_net_", f"Net {band_long} {{}}flux at roofs"))
        t.append(("wall_in", "wall_flux_in_", "wall_spectral_flux_in_", f"Incoming {band_long} {{}}flux at walls"))
        if flux.wall_in_dir is not None:
            t.append(("wall_in_dir", "wall_flux_in_direct_", "wall_spectral_flux_in_direct_",
                      f"Direct incoming {band_long} {{}}flux at walls"))
        t.append(("wall_net", "wall_flux_net_", "wall_spectral_flux_net_", f"Net {band_long} {{}}flux at walls"))
    if flux.clear_air_abs is not None:
        t.append(("clear_air_abs", "clear_air_absorption_", "clear_air_spectral_absorption_",
                  f"Absorbed {band_long} in clear air"))
    if flux.veg_abs is not None:
        t.append(("veg_abs", "veg_absorption_", "veg_spectral_absorption_", f"Absorbed {band_long} by vegetation"))
        t.append(("veg_air_abs", "veg_air_absorption_", "veg_air_spectral_absorption_",
                  f"Absorbed {band_long} by air in vegetated regions"))
    if flux.veg_abs_dir is not None:
        t.append(("veg_abs_dir", "veg_absorption_direct_", "veg_spectral_absorption_direct_",
                  f"Absorbed direct {band_long} by vegetation"))
    if flux.flux_dn_layer_top is not None:
        t.append(("flux_dn_layer_top", "flux_dn_layer_top_", "spectral_flux_dn_layer_top_",
                  f"Downwelling {band_long} {{}}flux at top of layer"))
        if flux.flux_dn_dir_layer_top is not None:
            t.append(("flux_dn_dir_layer_top", "flux_dn_direct_layer_top_", "spectral_flux_dn_direct_layer_top_",
                      f"Downwelling direct {band_long} {{}}flux at top of layer"))
        t.append(("flux_up_layer_top", "flux_up_layer_top_", "spectral_flux_up_layer_top_",
                  f"Upwelling {band_long} {{}}flux at top of layer"))
        t.append(("flux_dn_layer_base", "flux_dn_layer_base_", "spectral_flux_dn_layer_base_",
                  f"Downwelling {band_long} {{}}flux at base of layer"))
        if flux.flux_dn_dir_layer_base is not None:
            t.append(("flux_dn_dir_layer_base", "flux_dn_direct_layer_base_", "spectral_flux_dn_direct_layer_base_",
                      f"Downwelling direct {band_long} {{}}flux at base of layer"))
        t.append(("flux_up_layer_base", "flux_up_layer_base_", "spectral_flux_up_layer_base_",
                  f"Upwelling {band_long} {{}}flux at base of layer"))
    return t


class _Unpacker:
    """unpack_variable / _broadband / _spectral (radsurf_save.F90:629-693): packed ragged layers ->
    (column, layer) padded with the fill value.  The reference walks the packed array from its
    first element in column order (it does not use istartlay), which is how read_input packs it."""

    def __init__(self, nlay, nmaxlay):
        nlay = np.asarray(nlay, dtype=np.int64)
        self.ncol, self.nmaxlay = nlay.size, nmaxlay
        self.mask = np.arange(nmaxlay)[None, :] < nlay[:, None]  # (column, layer)
        self.ntot = int(nlay.sum())

    def layer(self, packed):  # (ntotlay,) -> (column, layer)
        out = np.full((self.ncol, self.nmaxlay), FillValueFlux, dtype=np.float64)
        out[self.mask] = np.asarray(packed)[: self.ntot]
        return out

    def broadband(self, packed):  # (ntotlay, nspec) -> (column, layer), summed over the spectral index
        return self.layer(np.asarray(packed)[: self.ntot].sum(axis=1))

    def spectral(self, packed):  # (ntotlay, nspec) -> (column, layer, band)
        packed = np.asarray(packed)
        out = np.full((self.ncol, self.nmaxlay, packed.shape[1]), FillValueFlux, dtype=np.float64)
        out[self.mask] = packed[: self.ntot]
        return out


def _define(f, name, dims, long_name, units="W m-2", fill=None, dtype="f4", **attrs):
    v = f.createVariable(name, dtype, dims)
    v.long_name = long_name
    if units is not None:
        v.units = units
    if fill is not None:
        v._FillValue = np.float32(fill)
    for k, a in attrs.items():
        setattr(v, k, a)
    return v


def _write_band(f, band, band_long, flux, unpack, do_broadband, do_spectral):
    nspec = flux.nspec
    # wavelength-independent members (radsurf_save.F90:179-197, 452-470)
    if flux.ground_dn_dir is not None:
        _define(f, "ground_sunlit_fraction", ("column",), "Fraction of ground in direct sunlight", "1")[:] = \
            _host(flux.ground_sunlit_frac)
    if flux.roof_in_dir is not None:
        _define(f, "roof_sunlit_fraction", ("column", "layer"), "Fraction of roof in direct sunlight", "1",
                FillValueFlux)[:] = unpack.layer(_host(flux.roof_sunlit_frac))
        _define(f, "wall_sunlit_fraction", ("column", "layer"), "Fraction of wall in direct sunlight", "1",
                FillValueFlux)[:] = unpack.layer(_host(flux.wall_sunlit_frac))
    if flux.veg_abs_dir is not None:
        _define(f, "veg_sunlit_fraction", ("column", "layer"), "Fraction of vegetation in direct sunlight", "1",
                FillValueFlux)[:] = unpack.layer(_host(flux.veg_sunlit_frac))
    cols, lays = _column_table(flux, band_long), _layer_table(flux, band_long)
    if do_broadband:
        for member, stem, _, long_name in cols:
            _define(f, stem + band, ("column",), long_name.format(""))[:] = _host(getattr(flux, member)).sum(axis=1)
        for member, stem, _, long_name in lays:
            _define(f, stem + band, ("column", "layer"), long_name.format(""), fill=FillValueFlux)[:] = \
                unpack.broadband(_host(getattr(flux, member)))
    if do_spectral:
        bdim = "band_" + band
        for member, _, stem, long_name in cols:
            _define(f, stem + band, ("column", bdim), long_name.format("spectral "))[:] = \
                _host(getattr(flux, member)).reshape(-1, nspec)
        for member, _, stem, long_name in lays:
            _define(f, stem + band, ("column", "layer", bdim), long_name.format("spectral "),
                    fill=FillValueFlux)[:] = unpack.spectral(_host(getattr(flux, member)))


def save_canopy_fluxes(file_name, config, canopy_props, flux_sw, flux_lw, iverbose=None, fix_broadband_lw=False):
    """Write the summed flux objects of a run (driver/spartacus_surface_driver.F90:295-296)."""
    from scipy.io import netcdf_file
    do_spectral_sw = bool(config.do_sw and config.do_save_spectral_flux)
    do_broadband_sw = bool(config.do_sw and config.do_save_broadband_flux)
    do_spectral_lw = bool(config.do_lw and config.do_save_spectral_flux)
    do_broadband_lw = False
    if config.do_lw:
        # radsurf_save.F90:71 assigns do_broadband_sw here: with do_sw too this changes nothing, with
        # do_lw alone it switches the broadband SHORTWAVE flag on; do_broadband_lw stays false
        do_broadband_sw = bool(config.do_save_broadband_flux)
        if fix_broadband_lw:
            do_broadband_lw = bool(config.do_save_broadband_flux)
    do_broadband_sw = do_broadband_sw and bool(config.do_sw)  # (the flag is only consulted under do_sw)
    nlay = np.asarray(canopy_props.nlay, dtype=np.int64)
    ncol = int(canopy_props.ncol)
    nmaxlay = int(nlay.max()) if ncol else 0
    unpack = _Unpacker(nlay, nmaxlay)
    f = netcdf_file(file_name, "w", version=1)
    try:
        f.createDimension("column", ncol)
        f.createDimension("layer", nmaxlay)
        f.createDimension("layer_interface", nmaxlay + 1)
        if do_spectral_sw:
            f.createDimension("band_sw", flux_sw.nspec)
        if do_spectral_lw:
            f.createDimension("band_lw", flux_lw.nspec)
        f.title, f.references, f.source, f.comment = TITLE, REFERENCES, SOURCE, COMMENT
        height = np.full((ncol, nmaxlay + 1), -1.0)
        height[:, 0] = 0.0
        dz = _host(canopy_props.dz)
        if nmaxlay > 0:
            padded = np.zeros((ncol, nmaxlay))
            padded[unpack.mask] = dz[: unpack.ntot]
            cum = np.cumsum(padded, axis=1)
            height[:, 1:][unpack.mask] = cum[unpack.mask]
        _define(f, "height", ("column", "layer_interface"), "Height of layer interfaces above ground", "m", -1.0,
                standard_name="height")[:] = height
        _define(f, "surface_type", ("column",), "Surface type", None, dtype="i2", definition=SURFACE_TYPES)[:] = \
            np.asarray(canopy_props.i_representation, dtype=np.int16)
        _define(f, "nlayer", ("column",), "Number of active layers", None, dtype="i2")[:] = nlay.astype(np.int16)
        if config.do_sw:
            _write_band(f, "sw", "shortwave", flux_sw, unpack, do_broadband_sw, do_spectral_sw)
        if config.do_lw:
            _write_band(f, "lw", "longwave", flux_lw, unpack, do_broadband_lw, do_spectral_lw)
    finally:
        f.close()
